// All-pairs potential kernels (kernel 1 of the unbinding path), sm_100a.
//
// Replaces the double loop at fortran_modules/particle_subroutines.f90:497-510:
//   be(i) = sum_j [x_j != x_i && y_j != y_i && z_j != z_i]  m_j / sqrt(dx^2 + dy^2 + dz^2)
//
// Work decomposition: a persistent grid; every WARP takes tickets from a device counter.
// A ticket is (halo, target group, j-split): kTargets*32 targets held in registers against
// 1/S of every source segment of that halo.  Sources are streamed through a warp-private
// ring of shared-memory tiles filled by 1-D TMA bulk copies (cp.async.bulk, UBLKCP) that
// complete on a per-stage mbarrier, so no block-wide barrier is ever taken and warps of
// one block can work on different haloes.
//
// FAST kernel, per pair of sources (j, j+1) and one target (7 issue slots per interaction):
//   3 FADD2 (dx,dy,dz) + FMUL2 + 2 FFMA2 (r^2) + 2 FMNMX3 + 2 FSETP (exclusion predicate:
//   min(|dx|,|dy|,|dz|) > 0  <=>  all three coordinates differ) + 2 MUFU.RSQ + 2 predicated
//   FFMA (acc += m * rsqrt).  float32 partial sums over <= 32 sources are flushed into a
//   float64 accumulator.  Bound: MUFU, 16 interactions / clk / SM.
//
// EXACT kernel: one target per thread, sources strictly in ascending order, IEEE sqrt and
// divide, the reference's != predicate, float32 accumulator: bit-identical to the reference
// arithmetic (see oracle/halma_oracle.c for the contraction of r^2).
#include <cstdlib>

#include "halma_common.cuh"
#include "potential.h"

namespace halma {

constexpr int kTileJ = 128;                 // sources per shared-memory tile
constexpr int kStages = 3;                  // ring depth per warp
constexpr int kStageFloats = 4 * kTileJ;    // x | y | z | m
constexpr int kFlushQuads = 8;              // flush float32 partials every 8 quads = 32 sources
constexpr int kWarpsPerBlock = kPotentialBlock / 32;
constexpr int kSmemBytes = kWarpsPerBlock * kStages * (kStageFloats * 4 + 8);

// ---------------------------------------------------------------------------------------
// Cursor over the source tiles of one ticket: piece `s` of `S` of every segment, in
// segment order.  Uniform across the warp.
// ---------------------------------------------------------------------------------------
struct TileCursor {
    const HaloDesc *hd;
    int nseg, k, S, s, n_members, parity;
    int64_t base;
    int set, pos, end, flags;

    __device__ __forceinline__ void seek()
    {
        while (k < nseg) {
            const SegDesc sd = hd->seg[k];
            const int c = (sd.flags & kSegMembers) ? n_members : sd.count;
            const int per = (((c + S - 1) / S) + kTileJ - 1) / kTileJ * kTileJ;
            const int a = s * per;
            const int b = min(a + per, c);
            if (a < b) {
                base = sd.begin;
                set = (sd.flags & kSegMembers) ? parity : 2;
                pos = a;
                end = b;
                flags = sd.flags;
                return;
            }
            ++k;
        }
    }
    __device__ __forceinline__ void init(const HaloDesc *h, int S_, int s_, int n_members_, int parity_)
    {
        hd = h;
        nseg = h->nseg;
        k = 0;
        S = S_;
        s = s_;
        n_members = n_members_;
        parity = parity_;
        seek();
    }
    __device__ __forceinline__ bool valid() const { return k < nseg; }
    __device__ __forceinline__ int len() const { return min(kTileJ, end - pos); }
    __device__ __forceinline__ void next()
    {
        pos += kTileJ;
        if (pos >= end) {
            ++k;
            seek();
        }
    }
};

// Fill one stage with the cursor's tile.  Whole quads come by TMA; the last 1..3 sources of
// a segment are fetched with plain loads (never reading past the array) and the quad is
// padded with (inf, inf, inf, m = 0), which contributes exactly zero in both kernels.
__device__ __forceinline__ void issue_tile(const PotParams &p, const TileCursor &c, float *stage, uint64_t *bar,
                                           int lane)
{
    const F32Set sp = p.src[c.set];
    const int64_t g0 = c.base + c.pos;
    const int len = c.len();
    const int nfull = len & ~3;
    if (lane == 0) {
        const uint32_t bytes = static_cast<uint32_t>(nfull) * 4u;
        mbar_expect_tx(bar, 4u * bytes);
        if (bytes) {
            tma_load_1d(stage, sp.x + g0, bytes, bar);
            tma_load_1d(stage + kTileJ, sp.y + g0, bytes, bar);
            tma_load_1d(stage + 2 * kTileJ, sp.z + g0, bytes, bar);
            tma_load_1d(stage + 3 * kTileJ, sp.m + g0, bytes, bar);
        }
    }
    if (nfull != len && lane < 4) {
        const int e = nfull + lane;
        float x = __int_as_float(0x7f800000), y = x, z = x, m = 0.f;
        if (e < len) {
            x = sp.x[g0 + e];
            y = sp.y[g0 + e];
            z = sp.z[g0 + e];
            m = sp.m[g0 + e];
        }
        stage[e] = x;
        stage[kTileJ + e] = y;
        stage[2 * kTileJ + e] = z;
        stage[3 * kTileJ + e] = m;
        fence_proxy_async_smem();
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// Ticket decoding shared by both kernels.
// ---------------------------------------------------------------------------------------
struct Ticket {
    int h, group, s, S, n_tgt;
};

template <int kGroup>
__device__ __forceinline__ bool decode_ticket(const PotParams &p, int item, Ticket &t)
{
    // largest k with item_base[k] <= item
    int lo = 0, hi = p.n_halo;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.item_base[mid] <= item)
            lo = mid;
        else
            hi = mid;
    }
    const int local = item - p.item_base[lo];
    t.h = p.order[lo];
    t.n_tgt = p.cnt ? p.cnt[t.h] : p.halo[t.h].n0;
    t.S = p.nsplit[t.h];
    const int groups = (t.n_tgt + kGroup - 1) / kGroup;
    const int mine = (groups - p.rank + p.n_ranks - 1) / p.n_ranks;
    if (mine <= 0) return false;
    t.s = local / mine;
    t.group = (local % mine) * p.n_ranks + p.rank;
    return true;
}

// ---------------------------------------------------------------------------------------
// FAST kernel
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_fast(uint64_t X, uint64_t Y, uint64_t Z, float m0, float m1, float xi, float yi,
                                          float zi, float &acc)
{
    const uint64_t dx = sub2(X, pack2(xi, xi));
    const uint64_t dy = sub2(Y, pack2(yi, yi));
    const uint64_t dz = sub2(Z, pack2(zi, zi));
    const uint64_t r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    float dx0, dx1, dy0, dy1, dz0, dz1, r20, r21;
    unpack2(dx, dx0, dx1);
    unpack2(dy, dy0, dy1);
    unpack2(dz, dz0, dz1);
    unpack2(r2, r20, r21);
    const float t0 = min3abs(dx0, dy0, dz0);
    const float t1 = min3abs(dx1, dy1, dz1);
    const float i0 = rsqrt_ftz(r20);
    const float i1 = rsqrt_ftz(r21);
    if (t0 > 0.f) acc = fmaf(m0, i0, acc);
    if (t1 > 0.f) acc = fmaf(m1, i1, acc);
}

template <int T>
__device__ __forceinline__ void tile_fast(const float *__restrict__ stage, int len, const float (&xi)[T],
                                          const float (&yi)[T], const float (&zi)[T], double (&acc64)[T])
{
    const float4 *X = reinterpret_cast<const float4 *>(stage);
    const float4 *Y = reinterpret_cast<const float4 *>(stage + kTileJ);
    const float4 *Z = reinterpret_cast<const float4 *>(stage + 2 * kTileJ);
    const float4 *M = reinterpret_cast<const float4 *>(stage + 3 * kTileJ);
    const int nq = (len + 3) >> 2;
    for (int qb = 0; qb < nq; qb += kFlushQuads) {
        const int qe = min(qb + kFlushQuads, nq);
        float acc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.f;
#pragma unroll 2
        for (int q = qb; q < qe; ++q) {
            const float4 x4 = X[q], y4 = Y[q], z4 = Z[q], m4 = M[q];
            const uint64_t x01 = pack2(x4.x, x4.y), x23 = pack2(x4.z, x4.w);
            const uint64_t y01 = pack2(y4.x, y4.y), y23 = pack2(y4.z, y4.w);
            const uint64_t z01 = pack2(z4.x, z4.y), z23 = pack2(z4.z, z4.w);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                pair_fast(x01, y01, z01, m4.x, m4.y, xi[t], yi[t], zi[t], acc[t]);
                pair_fast(x23, y23, z23, m4.z, m4.w, xi[t], yi[t], zi[t], acc[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) acc64[t] += static_cast<double>(acc[t]);
    }
}

template <int T, int MINB>
__global__ void __launch_bounds__(kPotentialBlock, MINB) k_potential_fast(const PotParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LoopState *st = p.st;
    if (!st->any_active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ring = reinterpret_cast<float *>(smem_raw) + warp * (kStages * kStageFloats);
    uint64_t *bars =
        reinterpret_cast<uint64_t *>(smem_raw + kWarpsPerBlock * kStages * kStageFloats * 4) + warp * kStages;
    if (lane == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int n_items = st->n_items;
    const int parity = st->parity;
    uint32_t par = 0;         // per-stage wait parity bits
    int fill = 0, use = 0;    // ring positions (continue across tickets)
    constexpr int kGroup = 32 * T;

    for (;;) {
        int item = 0;
        if (lane == 0) item = static_cast<int>(atomicAdd(&p.st->counter, 1u));
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        Ticket tk;
        if (!decode_ticket<kGroup>(p, item, tk)) continue;
        const HaloDesc *hd = &p.halo[tk.h];
        const int64_t tbase = p.tgt_members ? hd->poff : 0;
        const int tsel = p.tgt_members ? parity : 0;

        float xi[T], yi[T], zi[T];
        double acc64[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int i = tk.group * kGroup + t * 32 + lane;
            const bool ok = i < tk.n_tgt;
            xi[t] = ok ? p.tx[tsel][tbase + i] : 0.f;
            yi[t] = ok ? p.ty[tsel][tbase + i] : 0.f;
            zi[t] = ok ? p.tz[tsel][tbase + i] : 0.f;
            acc64[t] = 0.0;
        }

        TileCursor prod, cons;
        prod.init(hd, tk.S, tk.s, tk.n_tgt, parity);
        cons = prod;
#pragma unroll 1
        for (int i = 0; i < kStages - 1 && prod.valid(); ++i) {
            issue_tile(p, prod, ring + fill * kStageFloats, &bars[fill], lane);
            fill = (fill + 1 == kStages) ? 0 : fill + 1;
            prod.next();
        }
#pragma unroll 1
        while (cons.valid()) {
            if (prod.valid()) {
                issue_tile(p, prod, ring + fill * kStageFloats, &bars[fill], lane);
                fill = (fill + 1 == kStages) ? 0 : fill + 1;
                prod.next();
            }
            mbar_wait(&bars[use], (par >> use) & 1u);
            par ^= 1u << use;
            tile_fast<T>(ring + use * kStageFloats, cons.len(), xi, yi, zi, acc64);
            __syncwarp();      // every lane is done with this stage before it is refilled
            use = (use + 1 == kStages) ? 0 : use + 1;
            cons.next();
        }

        double *out = p.phi_part + static_cast<int64_t>(tk.s) * p.phi_stride + tbase;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int i = tk.group * kGroup + t * 32 + lane;
            if (i < tk.n_tgt) out[i] = acc64[t];
        }
    }
}

// ---------------------------------------------------------------------------------------
// EXACT kernel
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float term_exact(float xs, float ys, float zs, float m, float xi, float yi, float zi)
{
    // particle_subroutines.f90:499-505.  Contraction of r^2 as GCC emits it for the
    // reference build flags: fma(dz,dz, fma(dx,dx, dy*dy)).  An excluded pair adds +0,
    // which leaves the (non-negative) float32 accumulator unchanged.
    const bool take = (xs != xi) && (ys != yi) && (zs != zi);
    const float dx = __fsub_rn(xs, xi);
    const float dy = __fsub_rn(ys, yi);
    const float dz = __fsub_rn(zs, zi);
    const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const float q = __fdiv_rn(m, __fsqrt_rn(r2));
    return take ? q : 0.f;
}

__device__ __forceinline__ void tile_exact(const float *__restrict__ stage, int len, float xi, float yi, float zi,
                                           float &acc)
{
    const float4 *X = reinterpret_cast<const float4 *>(stage);
    const float4 *Y = reinterpret_cast<const float4 *>(stage + kTileJ);
    const float4 *Z = reinterpret_cast<const float4 *>(stage + 2 * kTileJ);
    const float4 *M = reinterpret_cast<const float4 *>(stage + 3 * kTileJ);
    const int nq = (len + 3) >> 2;
#pragma unroll 2
    for (int q = 0; q < nq; ++q) {
        const float4 x4 = X[q], y4 = Y[q], z4 = Z[q], m4 = M[q];
        const float t0 = term_exact(x4.x, y4.x, z4.x, m4.x, xi, yi, zi);
        const float t1 = term_exact(x4.y, y4.y, z4.y, m4.y, xi, yi, zi);
        const float t2 = term_exact(x4.z, y4.z, z4.z, m4.z, xi, yi, zi);
        const float t3 = term_exact(x4.w, y4.w, z4.w, m4.w, xi, yi, zi);
        acc = __fadd_rn(acc, t0);      // ascending source order, one rounding per add
        acc = __fadd_rn(acc, t1);
        acc = __fadd_rn(acc, t2);
        acc = __fadd_rn(acc, t3);
    }
}

__global__ void __launch_bounds__(kPotentialBlock, 4) k_potential_exact(const PotParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LoopState *st = p.st;
    if (!st->any_active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ring = reinterpret_cast<float *>(smem_raw) + warp * (kStages * kStageFloats);
    uint64_t *bars =
        reinterpret_cast<uint64_t *>(smem_raw + kWarpsPerBlock * kStages * kStageFloats * 4) + warp * kStages;
    if (lane == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int n_items = st->n_items;
    const int parity = st->parity;
    uint32_t par = 0;
    int fill = 0, use = 0;

    for (;;) {
        int item = 0;
        if (lane == 0) item = static_cast<int>(atomicAdd(&p.st->counter, 1u));
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        Ticket tk;
        if (!decode_ticket<32>(p, item, tk)) continue;
        const HaloDesc *hd = &p.halo[tk.h];
        const int64_t tbase = p.tgt_members ? hd->poff : 0;
        const int tsel = p.tgt_members ? parity : 0;
        const int i = tk.group * 32 + lane;
        const bool ok = i < tk.n_tgt;
        const float xi = ok ? p.tx[tsel][tbase + i] : 0.f;
        const float yi = ok ? p.ty[tsel][tbase + i] : 0.f;
        const float zi = ok ? p.tz[tsel][tbase + i] : 0.f;

        float total = 0.f, cls = 0.f;      // halo_gas.py:301 `binding_energy = zeros(float32)`
        int cur_seg = -1;
        TileCursor prod, cons;
        prod.init(hd, 1, 0, tk.n_tgt, parity);      // EXACT never splits the source range
        cons = prod;
#pragma unroll 1
        for (int k = 0; k < kStages - 1 && prod.valid(); ++k) {
            issue_tile(p, prod, ring + fill * kStageFloats, &bars[fill], lane);
            fill = (fill + 1 == kStages) ? 0 : fill + 1;
            prod.next();
        }
#pragma unroll 1
        while (cons.valid()) {
            if (prod.valid()) {
                issue_tile(p, prod, ring + fill * kStageFloats, &bars[fill], lane);
                fill = (fill + 1 == kStages) ? 0 : fill + 1;
                prod.next();
            }
            if (cons.k != cur_seg) {
                // a new class: fold the finished class sum into the total in float32
                // (`binding_energy += binding_energy_<class>`, halo_gas.py:322,328,359,...)
                if (cons.flags & kSegNewClass) {
                    total = __fadd_rn(total, cls);
                    cls = 0.f;
                }
                cur_seg = cons.k;
            }
            mbar_wait(&bars[use], (par >> use) & 1u);
            par ^= 1u << use;
            tile_exact(ring + use * kStageFloats, cons.len(), xi, yi, zi, cls);
            __syncwarp();
            use = (use + 1 == kStages) ? 0 : use + 1;
            cons.next();
        }
        total = __fadd_rn(total, cls);
        if (ok) p.phi_part[tbase + i] = static_cast<double>(total);      // exact in float64
    }
}

// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
namespace {

// Shapes of the FAST kernel: T targets per lane (group = 32 T targets per ticket) and the
// minimum resident blocks per SM the register allocation is held to.  Variant 0 is the
// default; HALMA_FAST_VARIANT=<index> selects another one (tuning sweeps).
struct FastVariant {
    int targets, min_blocks;
    void (*kernel)(const PotParams);
};

const FastVariant kVariants[] = {
    {4, 6, k_potential_fast<4, 6>}, {4, 5, k_potential_fast<4, 5>}, {4, 4, k_potential_fast<4, 4>},
    {2, 8, k_potential_fast<2, 8>}, {2, 6, k_potential_fast<2, 6>}, {3, 5, k_potential_fast<3, 5>},
    {3, 6, k_potential_fast<3, 6>}, {6, 3, k_potential_fast<6, 3>}, {8, 2, k_potential_fast<8, 2>},
    {2, 10, k_potential_fast<2, 10>}, {1, 12, k_potential_fast<1, 12>},
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

const FastVariant &fast_variant()
{
    static int idx = [] {
        const char *e = getenv("HALMA_FAST_VARIANT");
        const int v = e ? atoi(e) : 0;
        return (v >= 0 && v < kNumVariants) ? v : 0;
    }();
    return kVariants[idx];
}

}  // namespace

int potential_group_size(int mode) { return mode == HALMA_MODE_EXACT ? 32 : 32 * fast_variant().targets; }

cudaError_t potential_configure(int mode, int *blocks_per_sm)
{
    cudaError_t e;
    if (mode == HALMA_MODE_EXACT) {
        e = cudaFuncSetAttribute(k_potential_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_potential_exact, kPotentialBlock,
                                                             kSmemBytes);
    }
    const FastVariant &v = fast_variant();
    e = cudaFuncSetAttribute(v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, v.kernel, kPotentialBlock, kSmemBytes);
}

cudaError_t potential_launch(const PotParams &p, int mode, int grid_blocks, cudaStream_t stream)
{
    if (mode == HALMA_MODE_EXACT)
        k_potential_exact<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    else
        fast_variant().kernel<<<grid_blocks, kPotentialBlock, kSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace halma
