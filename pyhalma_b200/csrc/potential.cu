// All-pairs potential kernels (kernel 1 of the unbinding path), sm_100a.
//
// Replaces the double loop at fortran_modules/particle_subroutines.f90:497-510:
//   be(i) = sum_j [x_j != x_i && y_j != y_i && z_j != z_i]  m_j / sqrt(dx^2 + dy^2 + dz^2)
//
// Work decomposition: a persistent grid; every WARP takes tickets from a device counter.
// A main ticket is (halo, target group, j-split): 32*T targets held in registers against
// 1/S of every source segment of that halo.  Sources are streamed through a warp-private
// ring of shared-memory tiles filled by 1-D TMA bulk copies (cp.async.bulk, UBLKCP) that
// complete on a per-stage mbarrier, so no block-wide barrier is ever taken and warps of
// one block can work on different haloes.
//
// Four arithmetic paths share that machinery:
//
// EXACT  one target per thread, sources strictly in ascending order, IEEE sqrt and divide
//        (their branch-free fast paths inside a vetted exponent window, the library routines
//        otherwise), the reference's != predicate, float32 accumulator: bit-identical to the
//        reference (see oracle/halma_oracle.c for the contraction of r^2).
//
// FAST, predicated ("PRED")  per pair of sources and one target: 3 FADD2 + FMUL2 + 2 FFMA2,
//        2 FMNMX3 + 2 FSETP (min(|dx|,|dy|,|dz|) > 0 <=> all three coordinates differ),
//        2 MUFU.RSQ, 2 predicated FFMA.  Packed f32x2 instructions hold the dispatch port
//        for two cycles, so this body is dispatch-bound at ~21 cycles per source pair
//        against MUFU's 16 (profiles/ncu_fast_r01.md).
//
// FAST, predicate-free ("NP", default for plans)  the two predicate instructions per
//        interaction leave the hot loop: main tickets sum over ALL sources with a packed
//        accumulate (18 cycles per pair); only the tile that contains the group's own
//        members keeps an r^2 > 0 guard (self pairs).  The pairs the reference excludes are
//        then subtracted exactly by CORRECTION tickets: for each axis the halo's sources are
//        kept sorted by that coordinate's bit pattern (sortprep.cu), so the sources that can
//        share a coordinate with 128 consecutive members of the sorted order form one
//        contiguous range; the correction body counts a pair iff it shares this axis and no
//        lower axis (each excluded pair exactly once).  A sum that comes out non-finite
//        (zero separation outside the own tile: exact duplicates) flags the halo, which is
//        then recomputed by the predicated kernel, so correctness never rests on the fast
//        path.  float32 partial sums over <= 32 sources are flushed into float64.
//
// FAST, symmetric self-term ("SYM", on top of NP)  a pair of MEMBERS in different 128-member tiles
//        is evaluated once and feeds both particles (sym_ticket below): half the MUFU.RSQ work of
//        the member x member term, exact and therefore order-independent float64 accumulation.
//
// REUSE (the <..., REUSE = true> instantiations; halma_unbind_config.cache_external / .incremental)  the same
//        main tickets with a choice of WHICH sources they stream: the members only once the externals' sum
//        is cached (first pass: members, then the externals into their own planes), or only the members
//        the previous pass removed (incremental pass; loop_kernels.cu keeps the per-member sums).
#include <cstdlib>

#include "potential_device.cuh"

namespace halma {

// ---------------------------------------------------------------------------------------
// Stand-alone kernels: one launch = one potential pass (the multi-launch drivers and the f2py-level call).
// fused.cu runs the same pass functions inside the persistent loop kernel.
// ---------------------------------------------------------------------------------------
template <int T, int MINB, bool NP, int SYM = 0, bool REUSE = false>
__global__ void __launch_bounds__(kPotentialBlock, MINB) k_potential_fast(const PotParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LoopState *st = p.st;
    if (!st->any_active) return;
    if (p.redo_only && !st->redo_any) return;
    Ring rg;
    float *col;
    warp_ring_setup(smem_raw, rg, col);
    potential_pass_fast<T, NP, SYM, REUSE>(p, rg, col, p.redo_only != 0);
}

__global__ void __launch_bounds__(kPotentialBlock, 4) k_potential_exact(const PotParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (!p.st->any_active) return;
    Ring rg;
    float *col;
    warp_ring_setup(smem_raw, rg, col);
    potential_pass_exact(p, rg);
}

// Device self-test of exact_term_fast against __fdiv_rn(m, __fsqrt_rn(r2)).
// part 0: every float32 mantissa x both exponent parities of r2 (the square root only depends on
//         those), m = 1;  part 1: pseudo-random (m, r2) over the whole safe window.
__global__ void k_selftest_exact(int64_t n_random, uint64_t seed, unsigned long long *mismatch)
{
    const int64_t n_sqrt = int64_t(1) << 24;
    const int64_t total = n_sqrt + n_random;
    unsigned long long bad = 0;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float m, r2;
        if (i < n_sqrt) {
            const uint32_t mant = static_cast<uint32_t>(i) & 0x7fffffu, par = static_cast<uint32_t>(i >> 23);
            r2 = __uint_as_float(((110u + par) << 23) | mant);
            m = 1.0f;
        } else {
            uint64_t z = seed + 0x9e3779b97f4a7c15ull * static_cast<uint64_t>(i);       // splitmix64
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            z ^= z >> 31;
            uint64_t w = (z ^ 0xd1b54a32d192ed03ull) * 0x2545f4914f6cdd1dull;
            w ^= w >> 29;
            // exponent of r2 in [26, 253] (2^-101 .. 2^126), of m in [67, 186] (2^-60 .. 2^59)
            const uint32_t e_r = 26u + static_cast<uint32_t>((z >> 40) % 228u), e_m = 67u + static_cast<uint32_t>((w >> 40) % 120u);
            uint32_t mant_r = static_cast<uint32_t>(z) & 0x7fffffu, mant_m = static_cast<uint32_t>(w) & 0x7fffffu;
            const uint32_t kind = static_cast<uint32_t>(z >> 60);
            if (kind == 0) mant_r = 0x7fffffu - (mant_r & 0xffu);          // just below a power of two
            if (kind == 1) mant_r &= 0xffu;                                // just above
            if (kind == 2) mant_m = 0x7fffffu - (mant_m & 0xffu);
            if (kind == 3) mant_m &= 0xffu;
            r2 = __uint_as_float((e_r << 23) | mant_r);
            m = __uint_as_float((e_m << 23) | mant_m | ((static_cast<uint32_t>(w >> 62) & 1u) << 31));
            if (kind == 4) m = 0.0f;
        }
        if (!exact_r2_safe(r2) || !exact_mass_safe(m)) {
            ++bad;          // the generator must stay inside the window
            continue;
        }
        const float want = __fdiv_rn(m, __fsqrt_rn(r2));
        const float got = exact_term_fast(m, r2);
        if (__float_as_uint(want) != __float_as_uint(got)) ++bad;
    }
    if (bad) atomicAdd(mismatch, bad);
}

cudaError_t potential_selftest_exact(int64_t n_random, uint64_t seed, unsigned long long *d_mismatch, int sm_count,
                                     cudaStream_t stream)
{
    k_selftest_exact<<<sm_count * 8, 256, 0, stream>>>(n_random, seed, d_mismatch);
    return cudaGetLastError();
}



// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
namespace {

// Shapes of the FAST kernels: T targets per lane (a ticket = 32 T targets) and the minimum
// resident blocks per SM the register allocation is held to.  Shape 0 is the throughput
// shape; shape 1 (32 targets per ticket, 4x as many tickets) is picked when shape 0 could not
// give every resident warp a ticket, i.e. for small haloes that are latency-bound.
// HALMA_FAST_VARIANT=<index> forces one shape (tuning sweeps).
struct FastVariant {
    int targets, min_blocks;
    void (*pred)(const PotParams);
    void (*np)(const PotParams);
    void (*np_reuse)(const PotParams);      // external-sum cache / incremental passes
    void (*pred_reuse)(const PotParams);    // ... on the predicated path (plans too small for the sorted copies)
};

#define HALMA_VARIANT(T, B) \
    {T, B, k_potential_fast<T, B, false>, k_potential_fast<T, B, true>, k_potential_fast<T, B, true, 0, true>, \
     k_potential_fast<T, B, false, 0, true>}
const FastVariant kVariants[] = {
    HALMA_VARIANT(4, 6), HALMA_VARIANT(1, 8), HALMA_VARIANT(2, 8), HALMA_VARIANT(4, 4),
    HALMA_VARIANT(3, 6), HALMA_VARIANT(6, 3), HALMA_VARIANT(8, 3),
};
#undef HALMA_VARIANT
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

// The kernels that also run the symmetric tickets.  4 row members per lane: 6 resident blocks per SM (80 registers, a
// few spill slots outside the hot loop) beat 5 and 4 by 2-5 % (profiles/np_variants_r01.txt).  8 row members per
// lane (pairs of row tiles): 5 blocks per SM at 96 registers (6 would spill inside the rotation loop).
typedef void (*SymKernel)(const PotParams);
const SymKernel kSymKernels[4] = {k_potential_fast<4, 4, true, 4>, k_potential_fast<4, 5, true, 4>,
                                  k_potential_fast<4, 6, true, 4>, k_potential_fast<4, 5, true, 8>};
const SymKernel kSymReuseKernels[4] = {k_potential_fast<4, 4, true, 4, true>, k_potential_fast<4, 5, true, 4, true>,
                                       k_potential_fast<4, 6, true, 4, true>, k_potential_fast<4, 5, true, 8, true>};
int sym_choice()
{
    static int c = [] {
        const char *e = getenv("HALMA_SYM_MINB");       // tuning: 4, 5 or 6 resident blocks per SM
        const int v = e ? atoi(e) : 6;
        return v >= 4 && v <= 6 ? v - 4 : 0;
    }();
    return c;
}
int g_sym_bps[4] = {0, 0, 0, 0}, g_sym_reuse_bps[4] = {0, 0, 0, 0};
int g_bps[16] = {0};

int forced_variant()
{
    static int idx = [] {
        const char *e = getenv("HALMA_FAST_VARIANT");
        const int v = e ? atoi(e) : -1;
        return (v >= 0 && v < kNumVariants) ? v : -1;
    }();
    return idx;
}

}  // namespace

int potential_num_variants() { return kNumVariants; }

int potential_pick_variant(int64_t groups_of_128, int64_t max_sources, int resident_warps, int64_t max_members,
                           bool symmetric, int max_split, int min_split_sources)
{
    if (forced_variant() >= 0) return forced_variant();
    // symmetric tickets need the 128-member tiles of the throughput shape and bring their own
    // parallelism (tiles^2 / 2 tile pairs in chunks of 2..32): worth it from 64 tiles on
    // (scripts/probes/midsize_probe.py)
    if (symmetric && max_members >= 64 * 128) return 0;
    // tickets the throughput shape would have after the j-split (at most max_split ways, pieces of
    // >= min_split_sources sources): if that cannot occupy about half the resident warps, use small tickets.
    // (Shape 1 has one target per lane: four times the tickets, but one shared-memory load per evaluation, which
    // makes it LSU-bound at about a third of the throughput shape's rate -- scripts/cfg1_passes.py.)
    int64_t split = max_sources / min_split_sources;
    split = split < 1 ? 1 : (split > max_split ? max_split : split);
    return 2 * groups_of_128 * split < resident_warps ? 1 : 0;
}

int potential_group_size(int mode, int variant)
{
    return mode == HALMA_MODE_EXACT ? 32 : 32 * kVariants[variant].targets;
}

cudaError_t potential_configure(int mode, int variant, int *blocks_per_sm)
{
    cudaError_t e;
    if (mode == HALMA_MODE_EXACT) {
        e = cudaFuncSetAttribute(k_potential_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_potential_exact, kPotentialBlock,
                                                             kSmemBytes);
    }
    const FastVariant &v = kVariants[variant];
    e = cudaFuncSetAttribute(v.pred, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(v.np, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    int a = 0, b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, v.pred, kPotentialBlock, kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, v.np, kPotentialBlock, kSmemBytes);
    if (e != cudaSuccess) return e;
    // the reuse kernels are launched with the same grid: they are held to the same minimum residency
    e = cudaFuncSetAttribute(v.np_reuse, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    int c = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, v.np_reuse, kPotentialBlock, kSmemBytes);
    b = b < c ? b : c;
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(v.pred_reuse, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, v.pred_reuse, kPotentialBlock, kSmemBytes);
    b = b < c ? b : c;
    *blocks_per_sm = a < b ? a : b;
    if (variant < 16) g_bps[variant] = *blocks_per_sm;
    if (e == cudaSuccess && variant == 0) {
        for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
            e = cudaFuncSetAttribute(kSymKernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
            if (e == cudaSuccess)
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_sym_bps[k], kSymKernels[k], kPotentialBlock, kSmemBytes);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(kSymReuseKernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
            if (e == cudaSuccess)
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_sym_reuse_bps[k], kSymReuseKernels[k], kPotentialBlock,
                                                                  kSmemBytes);
        }
    }
    return e;
}

cudaError_t potential_launch(const PotParams &p, int mode, int variant, int grid_blocks, cudaStream_t stream)
{
    if (mode == HALMA_MODE_EXACT)
        k_potential_exact<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    else if (p.np_enabled && !p.redo_only && p.sym_enabled) {
        const int per_sm = variant < 16 && g_bps[variant] > 0 ? g_bps[variant] : 1;
        const int k = p.sym_rows == 8 ? 3 : sym_choice();
        if (p.cache_ext || p.incr_enabled)
            kSymReuseKernels[k]<<<grid_blocks / per_sm * (g_sym_reuse_bps[k] > 0 ? g_sym_reuse_bps[k] : 1), kPotentialBlock,
                                  kSmemBytes, stream>>>(p);
        else
            kSymKernels[k]<<<grid_blocks / per_sm * (g_sym_bps[k] > 0 ? g_sym_bps[k] : 1), kPotentialBlock, kSmemBytes,
                             stream>>>(p);
    } else if (p.np_enabled && !p.redo_only && (p.cache_ext || p.incr_enabled))
        kVariants[variant].np_reuse<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    else if (p.np_enabled && !p.redo_only)
        kVariants[variant].np<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    else if (!p.redo_only && (p.cache_ext || p.incr_enabled))
        kVariants[variant].pred_reuse<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    else
        kVariants[variant].pred<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace halma
