// The whole unbinding loop of a plan as ONE persistent, cooperative kernel (sm_100a): the default driver
// on a single GPU.  north_star: "all three run in a device-resident loop until the bound set stops
// changing, with no host round-trip per iteration".
//
// The grid is one wave of co-resident blocks (SMs x the occupancy of the potential code, 128 threads each);
// blocks move through the phases of a pass together, separated by grid barriers:
//
//   [prologue]  masses of the sorted copies restored, (optional) pack, per-halo initial decision,
//               barrier, commit + first ticket table, barrier
//   pass:       potential tickets (potential_device.cuh; + the predicated re-evaluation of flagged haloes)
//               barrier
//               energy step per chunk; the block that finishes a halo's last chunk decides the halo
//               compaction per chunk, waiting for ITS halo's decision only (per-halo stamp, no barrier)
//               barrier
//               commit per halo (all blocks) next to the ticket table of the next pass (block 0)
//               barrier
//   [epilogue]  member lists
//
// Three grid barriers per pass replace six launches, a device-to-host flag copy and a host wake-up; a run is
// one kernel launch however many passes it takes.  Data written by one phase and read by TMA bulk copies in
// the next (the compacted working set, the sorted copies' masses, the removed members) crosses from the
// generic to the async proxy: writers fence (fence.proxy.async) before the barrier.
// The same phase functions run as stand-alone kernels in the multi-launch drivers (loop_kernels.cu), so the
// results are bit-identical across drivers.
#include <cooperative_groups.h>

#include "fused.h"
#include "loop_device.cuh"
#include "potential_device.cuh"

namespace cg = cooperative_groups;

namespace halma {

namespace {

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// T = 0: EXACT arithmetic; otherwise the FAST shape with T targets per lane.
template <int T, int MINB, bool NP, int SYM>
__global__ void __launch_bounds__(kPotentialBlock, MINB) k_unbind_loop(const PotParams pp, const LoopParams lp,
                                                                        const int do_pack)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    static_assert(kLT == kPotentialBlock, "the loop phases run on the potential kernel's blocks");
    static_assert(kSchedSmemBytes <= kWarpsPerBlock * kStages * kStageFloats * 4, "the ticket table is built in the ring buffers");
    cg::grid_group grid = cg::this_grid();
    LoopState *st = lp.st;
    const bool clock = blockIdx.x == 0 && threadIdx.x == 0;
    const unsigned long long t_begin = clock ? globaltimer_ns() : 0ull;
    unsigned long long t_ph[5] = {0ull, 0ull, 0ull, 0ull, 0ull};
    unsigned long long t_last = t_begin;
    int pass = 0, par = 0;
    auto lap = [&](int k) {          // block 0's clock; the barriers make its phase times the grid's
        if (clock) {
            const unsigned long long t = globaltimer_ns();
            t_ph[k] += t - t_last;
            if (k >= 1 && k <= 3 && pass < 16) st->pass_ns[pass][k - 1] = static_cast<unsigned int>(t - t_last);
            t_last = t;
        }
    };

    Ring rg;
    float *col;
    warp_ring_setup(smem_raw, rg, col);

    // ---- prologue ----
    if (lp.np_enabled) {
        // the loop zeroes the masses of removed members in the sorted copies: restore them
        const int64_t n = 3 * lp.n_spad, stride = static_cast<int64_t>(gridDim.x) * kLT;
        for (int64_t i = blockIdx.x * static_cast<int64_t>(kLT) + threadIdx.x; i < n; i += stride) {
            const int a = static_cast<int>(i / lp.n_spad);
            const int64_t k = i - a * lp.n_spad;
            lp.ax[a].m[k] = lp.ax_m0[a][k];
        }
    }
    if (do_pack) {
        pack_phase(lp);
        fence_proxy_async_global();
        grid.sync();
    }
    decide_init_phase(lp);
    fence_proxy_async_global();
    grid.sync();
    commit_phase(lp, 1);
    if (blockIdx.x == 0) schedule_block(lp, smem_raw, 1);      // the TMA ring is idle between passes
    grid.sync();
    lap(0);

    // ---- passes ----
    while (*reinterpret_cast<volatile int32_t *>(&st->any_active)) {
        const unsigned long long t_pot = (threadIdx.x & 31) == 0 ? globaltimer_ns() : 0ull;
        if (clock && pass < 16) st->dbg_items[pass] = static_cast<unsigned int>(st->n_items);
        if constexpr (T == 0)
            potential_pass_exact(pp, rg);
        else
            potential_pass_fast<T, NP, SYM, true>(pp, rg, col, false);
        if ((threadIdx.x & 31) == 0 && pass < 16) atomicAdd(&st->dbg_pot_busy[pass], globaltimer_ns() - t_pot);
        grid.sync();
        if constexpr (T != 0) {
            if (*reinterpret_cast<volatile int32_t *>(&st->redo_any)) {
                // haloes whose predicate-free sums came out non-finite, or whose incremental pass removed too
                // much of some member's potential, are recomputed in full with the predicate
                potential_pass_fast<T, false, 0, false>(pp, rg, col, true);
                grid.sync();
            }
        }
        lap(1);
        if (clock && pass < 16) st->dbg_phase_start[pass] = t_last;
        energy_phase(lp, par, pass);
        if ((threadIdx.x & 31) == 0 && pass < 16) atomicMax(&st->dbg_e_end[pass], globaltimer_ns());
        compact_phase(lp, par, pass, true);
        fence_proxy_async_global();
        grid.sync();
        lap(2);
        commit_phase(lp, 0);
        if (blockIdx.x == 0) schedule_block(lp, smem_raw, 0);
        grid.sync();
        lap(3);
        ++pass;
        par ^= 1;
    }

    // ---- epilogue ----
    finalize_phase(lp);
    {
        // run totals of the per-halo counters (exact integer atomics)
        unsigned long long pr = 0ull, ev = 0ull;
        for (int h = blockIdx.x * kLT + threadIdx.x; h < lp.n_halo; h += gridDim.x * kLT) {
            pr += lp.pairs[h];
            ev += lp.evals[h];
        }
        for (int o = 16; o > 0; o >>= 1) {
            pr += __shfl_down_sync(0xffffffffu, pr, o);
            ev += __shfl_down_sync(0xffffffffu, ev, o);
        }
        if ((threadIdx.x & 31) == 0 && (pr | ev)) {
            atomicAdd(&st->pairs_total, pr);
            atomicAdd(&st->evals_total, ev);
        }
    }
    lap(4);
    if (clock) {
        st->pot_ns = t_ph[1];
        st->loop_ns = t_last - t_begin;
        for (int k = 0; k < 5; ++k) st->phase_ns[k] = t_ph[k];
    }
}

typedef void (*LoopKernel)(const PotParams, const LoopParams, const int);

struct FusedEntry {
    LoopKernel fn;
    int blocks_per_sm;      // filled by fused_configure
};

// index: see fused_kernel_index()
FusedEntry g_fused[] = {
    {k_unbind_loop<0, 4, false, 0>, 0},      // EXACT
    {k_unbind_loop<4, 6, false, 0>, 0},      // FAST predicated, throughput shape (shape 0)
    {k_unbind_loop<1, 8, false, 0>, 0},      // FAST predicated, small-halo shape (shape 1)
    {k_unbind_loop<4, 6, true, 0>, 0},       // FAST predicate-free, shape 0
    {k_unbind_loop<1, 8, true, 0>, 0},       // FAST predicate-free, shape 1
    {k_unbind_loop<4, 6, true, 4>, 0},       // FAST predicate-free + symmetric self-term (4 row members per lane), shape 0
    {k_unbind_loop<4, 5, true, 8>, 0},       // ... with 8 row members per lane (pairs of row tiles)
};
constexpr int kNumFused = sizeof(g_fused) / sizeof(g_fused[0]);

}  // namespace

int fused_kernel_index(int mode, int variant, bool np, int sym_rows)
{
    const bool sym = sym_rows != 0;
    if (mode == HALMA_MODE_EXACT) return 0;
    if (variant != 0 && variant != 1) return -1;      // tuning shapes exist as stand-alone kernels only
    if (!np) return 1 + variant;
    if (sym) return variant == 0 ? (sym_rows == 8 ? 6 : 5) : -1;
    return 3 + variant;
}

cudaError_t fused_configure()
{
    for (int k = 0; k < kNumFused; ++k) {
        cudaError_t e = cudaFuncSetAttribute(g_fused[k].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_fused[k].blocks_per_sm, g_fused[k].fn, kPotentialBlock,
                                                          kSmemBytes);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int fused_blocks_per_sm(int index) { return index >= 0 && index < kNumFused ? g_fused[index].blocks_per_sm : 0; }

cudaError_t fused_launch(int index, const PotParams &pp, const LoopParams &lp, int do_pack, int sm_count,
                         cudaStream_t stream)
{
    if (index < 0 || index >= kNumFused || g_fused[index].blocks_per_sm < 1) return cudaErrorInvalidValue;
    void *args[] = {const_cast<PotParams *>(&pp), const_cast<LoopParams *>(&lp), &do_pack};
    return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(g_fused[index].fn),
                                       dim3(sm_count * g_fused[index].blocks_per_sm), dim3(kPotentialBlock), args,
                                       kSmemBytes, stream);
}

}  // namespace halma
