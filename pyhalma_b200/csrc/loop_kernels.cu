// Kernels 2 and 3 of the unbinding path and the device-side scheduler, sm_100a.
//
// One pass of the loop is five launches on one stream, none of which
// needs the host:
//   k_potential_*   (potential.cu)  Phi for every current member of every active halo
//   k_energy_flag   energy step + bound flag + per-chunk survivor count and mass sums
//                   (halo_properties.py:342-359 / halo_gas.py:456-476; sums :16-60)
//   k_halo_decide   per halo: scan of chunk counts, reduction of chunk sums -> new count,
//                   M, CoM, bulk velocity, converged / active
//   k_compact       stable (order-preserving) warp-aggregated stream compaction of the
//                   float32 working set into the other buffer
//   k_schedule      commits the per-halo state, builds the ticket table of the next pass
// All are O(N) and HBM-bound; the potential kernel dominates for N >~ 1e3.
//
// Reductions are done in a fixed order (tree inside a block, then ascending chunks), so a
// run is bit-reproducible.
#include "halma_common.cuh"
#include "loop_kernels.h"

namespace halma {

namespace {

constexpr int kCh = kChunk;      // 256 members per chunk == threads per block

// Deterministic block-wide sum of NV doubles per thread (blockDim.x == kCh).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* [NV * 8] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) smem[k * (kCh / 32) + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (int w = 0; w < kCh / 32; ++w) s += smem[k * (kCh / 32) + w];
            v[k] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem /* [kCh/32 + 1] */, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kCh / 32 ? smem[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        if (lane < kCh / 32) smem[lane] = wi - w;
        if (lane == kCh / 32 - 1) smem[kCh / 32] = wi;
    }
    __syncthreads();
    const int r = smem[warp] + incl - v;
    total = smem[kCh / 32];
    __syncthreads();
    return r;
}

// Groups of `gs` targets of a halo with n members that belong to this rank (split mode).
__device__ __forceinline__ int my_groups(int n, int gs, int rank, int n_ranks)
{
    const int groups = (n + gs - 1) / gs;
    return (groups - rank + n_ranks - 1) / n_ranks;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Pack: float64 user arrays -> float32 working set (round to nearest, like np.float32()).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCh) k_pack_members(const LoopParams p)
{
    __shared__ double red[kChunkSums * (kCh / 32)];
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        const HaloDesc &hd = p.halo[h];
        const int q = p.chunk_p0[c] + threadIdx.x;
        double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (q < hd.n0) {
            const int64_t i = hd.poff + q, g = hd.uoff + q;
            const double x = p.x64[g], y = p.y64[g], z = p.z64[g], m = p.m64[g];
            p.wx[0][i] = __double2float_rn(x);
            p.wy[0][i] = __double2float_rn(y);
            p.wz[0][i] = __double2float_rn(z);
            p.wm[0][i] = __double2float_rn(m);
            p.widx[0][i] = static_cast<int32_t>(g);
            s[0] = m;
            s[1] = m * p.vx[g];
            s[2] = m * p.vy[g];
            s[3] = m * p.vz[g];
            s[4] = m * x;
            s[5] = m * y;
            s[6] = m * z;
        }
        block_sum<kChunkSums>(s, red);
        if (threadIdx.x == 0) {
            p.chunk_cnt[c] = min(kCh, max(hd.n0 - p.chunk_p0[c], 0));
#pragma unroll
            for (int k = 0; k < kChunkSums; ++k) p.chunk_sum[static_cast<int64_t>(c) * kChunkSums + k] = s[k];
            p.chunk_best[c] = -1.f;
            p.chunk_best_q[c] = -1;
        }
    }
}

// Symmetric mode: largest coordinate extent of every halo's members (float32 working set, right
// after the pack).  With the bound mass it sets the quantum that makes the two-sided sums exact.
__global__ void __launch_bounds__(kCh) k_halo_extent(const LoopParams p)
{
    __shared__ float smn[3][kCh / 32], smx[3][kCh / 32];
    for (int h = blockIdx.x; h < p.n_halo; h += gridDim.x) {
        const HaloDesc &hd = p.halo[h];
        const float inf = __int_as_float(0x7f800000);
        float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
        for (int q = threadIdx.x; q < hd.n0; q += kCh) {
            const float v[3] = {p.wx[0][hd.poff + q], p.wy[0][hd.poff + q], p.wz[0][hd.poff + q]};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                mn[a] = fminf(mn[a], v[a]);
                mx[a] = fmaxf(mx[a], v[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            for (int o = 16; o > 0; o >>= 1) {
                mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], o));
            }
            if ((threadIdx.x & 31) == 0) {
                smn[a][threadIdx.x >> 5] = mn[a];
                smx[a][threadIdx.x >> 5] = mx[a];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ext = 0.0;
            for (int a = 0; a < 3; ++a) {
                float lo = smn[a][0], hi = smx[a][0];
                for (int w = 1; w < kCh / 32; ++w) {
                    lo = fminf(lo, smn[a][w]);
                    hi = fmaxf(hi, smx[a][w]);
                }
                ext = fmax(ext, static_cast<double>(hi) - static_cast<double>(lo));
            }
            p.sym_ext[h] = ext;          // NaN coordinates are skipped by fmin/fmax; an empty halo gives -inf -> 0 below
        }
        __syncthreads();
    }
}

// One external group: user CSR layout (float64) -> padded float32 segment of each halo.
__global__ void __launch_bounds__(256) k_pack_group(const HaloDesc *halo, int n_halo, int seg_index,
                                                    const int64_t *ext_off, const double *m, const double *x,
                                                    const double *y, const double *z, float *em, float *ex,
                                                    float *ey, float *ez)
{
    for (int h = blockIdx.y; h < n_halo; h += gridDim.y) {
        const SegDesc sd = halo[h].seg[seg_index];
        const int64_t u0 = ext_off[h];
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < sd.count; k += gridDim.x * blockDim.x) {
            em[sd.begin + k] = __double2float_rn(m[u0 + k]);
            ex[sd.begin + k] = __double2float_rn(x[u0 + k]);
            ey[sd.begin + k] = __double2float_rn(y[u0 + k]);
            ez[sd.begin + k] = __double2float_rn(z[u0 + k]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel 2: energy step, bound flag, survivor counts and mass sums per chunk.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCh) k_energy_flag(const LoopParams p)
{
    __shared__ double red[kChunkSums * (kCh / 32)];
    __shared__ float sbest[kCh / 32];
    __shared__ int sbest_q[kCh / 32];
    if (!p.st->any_active) return;
    const int par = p.st->parity;
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        const int p0 = p.chunk_p0[c];
        if (p0 >= n) continue;
        const HaloDesc &hd = p.halo[h];
        const int q = p0 + threadIdx.x;
        double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        int bound = 0;
        float best = -1.f;          // potentials are >= 0; NaN never wins
        int best_q = -1;
        if (q < n) {
            const int64_t i = hd.poff + q;
            const int64_t g = p.widx[par][i];
            // Phi: ascending sum of the j-split partials, rounded once to the f2py output dtype
            const int S = p.nsplit[h];
            double phi = p.phi_part[i];
            for (int k = 1; k < S; ++k) phi += p.phi_part[static_cast<int64_t>(k) * p.n_pad + i];
            // a halo handed to the predicated kernel has its complete, final sum in the planes
            const bool redo = p.np_enabled && p.halo_redo[h];
            const bool inc = p.incr_enabled && p.incr[h];
            const int64_t slot = hd.poff + (g - hd.uoff);      // the member's original slot
            if (inc && !redo) {
                // incremental pass: the planes hold what the members removed by the previous pass contributed
                // (reference predicate applied); take it out of the potential kept from that pass
                phi = p.phi_keep[slot] - phi;
            } else if (!redo) {
                if (p.sym_enabled) phi += p.phi_sym[i];
                if (p.cache_ext && hd.n_ext > 0 && p.ext_ok[h]) {
                    // sum over the external sources, evaluated by the first pass only (potential.cu)
                    if (p.st->pass == 0) {
                        double e = p.phi_ext[slot];
                        for (int k = 1; k < S; ++k) e += p.phi_ext[static_cast<int64_t>(k) * p.n_pad + slot];
                        p.phi_ext[slot] = e;
                    }
                    phi += p.phi_ext[slot];
                }
                if (p.np_enabled) {
                    // predicate-free path: take out the pairs that share a coordinate (potential.cu)
                    phi -= (p.ax[0].corr[slot] + p.ax[1].corr[slot]) + p.ax[2].corr[slot];
                }
            }
            // the complete float64 potential, kept for a following incremental pass
            if (p.incr_enabled && !redo) p.phi_keep[slot] = phi;
            const float be = __double2float_rn(phi);
            // halo_properties.py:342-351 / halo_gas.py:456-465: float32 chain, two roundings
            float pe = -be;
            pe = __fmul_rn(pe, p.G32);
            pe = __fmul_rn(pe, p.kappa32);
            // :354 / :468  float64, no contraction: 0.5*((dvx^2 + dvy^2) + dvz^2)
            const double dvx = __dsub_rn(p.vx[g], p.hvb[3 * h + 0]);
            const double dvy = __dsub_rn(p.vy[g], p.hvb[3 * h + 1]);
            const double dvz = __dsub_rn(p.vz[g], p.hvb[3 * h + 2]);
            const double ke = __dmul_rn(
                0.5, __dadd_rn(__dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy)), __dmul_rn(dvz, dvz)));
            const double E = __dadd_rn(ke, static_cast<double>(pe));
            bound = (E <= 0.0) ? 1 : 0;           // :359 / :476 (NaN is neither bound nor unbound)
            p.flag[i] = static_cast<uint8_t>(bound);
            p.out_mask[g] = static_cast<uint8_t>(bound);
            p.out_be[g] = be;
            p.out_E[g] = E;
            if (be > best) {
                best = be;
                best_q = static_cast<int>(g - hd.uoff);
            }
            if (p.temp) {
                // halo_gas.py:479-490: cold = T < 5e4, hot = T >= 5e4 (NaN is neither)
                const double T = p.temp[g], m = p.m64[g];
                const bool cold = T < p.cold_T, hot = T >= p.cold_T;
                if (bound && cold) s[7] = m;
                if (E > 0.0 && cold) s[8] = m;
                if (E > 0.0 && hot) s[9] = m;
            }
            if (bound) {
                const double m = p.m64[g];
                s[0] = m;
                s[1] = m * p.vx[g];
                s[2] = m * p.vy[g];
                s[3] = m * p.vz[g];
                s[4] = m * p.x64[g];
                s[5] = m * p.y64[g];
                s[6] = m * p.z64[g];
            }
        }
        const int count = __syncthreads_count(bound);
        block_sum<kChunkSums>(s, red);
        // most bound member of the chunk: largest potential, lowest index on ties
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_down_sync(0xffffffffu, best, o);
            const int oq = __shfl_down_sync(0xffffffffu, best_q, o);
            if (ob > best || (ob == best && oq >= 0 && (best_q < 0 || oq < best_q))) {
                best = ob;
                best_q = oq;
            }
        }
        if ((threadIdx.x & 31) == 0) {
            sbest[threadIdx.x >> 5] = best;
            sbest_q[threadIdx.x >> 5] = best_q;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            p.chunk_cnt[c] = count;
#pragma unroll
            for (int k = 0; k < kChunkSums; ++k) p.chunk_sum[static_cast<int64_t>(c) * kChunkSums + k] = s[k];
            for (int w = 1; w < kCh / 32; ++w)
                if (sbest[w] > best || (sbest[w] == best && sbest_q[w] >= 0 && (best_q < 0 || sbest_q[w] < best_q))) {
                    best = sbest[w];
                    best_q = sbest_q[w];
                }
            p.chunk_best[c] = best;
            p.chunk_best_q[c] = best_q;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Per halo: exclusive scan of chunk counts, ordered reduction of the chunk sums, and the
// convergence decision.  init = 1 right after k_pack_members (no pass made yet).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCh) k_halo_decide(const LoopParams p, int init)
{
    __shared__ int sscan[kCh / 32 + 1];
    __shared__ double red[kChunkSums * (kCh / 32)];
    if (!init && !p.st->any_active) return;
    for (int h = blockIdx.x; h < p.n_halo; h += gridDim.x) {
        if (!init && !p.active[h]) continue;
        const HaloDesc &hd = p.halo[h];
        const int n_old = init ? hd.n0 : p.cnt[h];
        const int nch = (n_old + kCh - 1) / kCh;
        int carry = 0;
        double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        float best = -1.f;          // most bound member of this pass (largest potential, lowest index)
        int best_q = -1;
        for (int c0 = 0; c0 < nch; c0 += kCh) {
            const int c = c0 + threadIdx.x;
            const int v = (c < nch) ? p.chunk_cnt[hd.chunk_begin + c] : 0;
            int total;
            const int ex = block_exclusive_scan(v, sscan, total);
            if (c < nch) {
                p.chunk_off[hd.chunk_begin + c] = carry + ex;
                const double *cs = p.chunk_sum + static_cast<int64_t>(hd.chunk_begin + c) * kChunkSums;
#pragma unroll
                for (int k = 0; k < kChunkSums; ++k) s[k] += cs[k];
                const float b = p.chunk_best[hd.chunk_begin + c];
                const int bq = p.chunk_best_q[hd.chunk_begin + c];
                if (b > best || (b == best && bq >= 0 && (best_q < 0 || bq < best_q))) {
                    best = b;
                    best_q = bq;
                }
            }
            carry += total;
        }
        // block-wide (value, index) reduction with the same tie rule
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_down_sync(0xffffffffu, best, o);
            const int oq = __shfl_down_sync(0xffffffffu, best_q, o);
            if (ob > best || (ob == best && oq >= 0 && (best_q < 0 || oq < best_q))) {
                best = ob;
                best_q = oq;
            }
        }
        __shared__ float hb[kCh / 32];
        __shared__ int hq[kCh / 32];
        if ((threadIdx.x & 31) == 0) {
            hb[threadIdx.x >> 5] = best;
            hq[threadIdx.x >> 5] = best_q;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 1; w < kCh / 32; ++w)
                if (hb[w] > best || (hb[w] == best && hq[w] >= 0 && (best_q < 0 || hq[w] < best_q))) {
                    best = hb[w];
                    best_q = hq[w];
                }
        block_sum<kChunkSums>(s, red);
        if (threadIdx.x == 0) {
            const int n_new = carry;
            const double M = s[0];
            if (init) {
                p.hrps[4 * h + 0] = M;
                p.hrps[4 * h + 1] = p.hrps[4 * h + 2] = p.hrps[4 * h + 3] = 0.0;
                p.hbest[h] = -1;
            } else {
                p.hrps[4 * h + 1] = s[7];            // cold members bound after this pass
                p.hrps[4 * h + 2] += s[8];           // removed, cold
                p.hrps[4 * h + 3] += s[9];           // removed, hot
                p.hbest[h] = best_q;
            }
            p.hM[h] = M;
            if (p.sym_enabled) {
                // Quantum of the symmetric sums of the coming pass: every addend is rounded to a multiple of
                // q = 2^-37 * 2^ceil(log2(M / extent)), and a sum that stays below 2^52 q = 32768 * (1..2) * M / extent
                // is then EXACT in float64, so the order of the atomics cannot change it.  A sum that leaves the
                // window sends the halo to the one-sided kernel (potential.cu::sym_ticket).  0 = no quantisation.
                const double ext = p.sym_ext[h];
                double q = 0.0;
                if (M > 0.0 && ext > 0.0 && M <= 1.7976931348623157e308 && ext <= 1.7976931348623157e308) {
                    int e;
                    frexp(M / ext, &e);                   // M / ext = f * 2^e, 0.5 <= f < 1
                    if (e > -900 && e < 900) q = ldexp(1.0, e - 37);
                }
                p.sym_q[h] = q;
            }
            const double inv = M > 0.0 ? 1.0 / M : 0.0;
            // halo_properties.py:39-43, 56-60: sums divided by M, zeros when M == 0
            for (int k = 0; k < 3; ++k) {
                p.hcom[3 * h + k] = M > 0.0 ? s[4 + k] / M : 0.0;
                p.hvb_next[3 * h + k] = p.vb_fixed ? p.hvb[3 * h + k] : (M > 0.0 ? s[1 + k] / M : 0.0);
            }
            (void)inv;
            p.cnt_next[h] = n_new;
            int act;
            if (init) {
                p.iter[h] = 0;
                p.converged[h] = (n_new == 0) ? 1 : 0;
                p.pairs[h] = 0ull;
                p.evals[h] = 0ull;
                act = (n_new > 0 && p.max_iter > 0) ? 1 : 0;
                if (p.cache_ext) p.ext_ok[h] = 1;
                if (p.incr_enabled) {
                    p.incr[h] = 0;
                    p.rem_cnt[h] = 0;
                }
            } else {
                const int it = p.iter[h] + 1;
                p.iter[h] = it;
                const unsigned long long nn = static_cast<unsigned long long>(n_old);
                p.pairs[h] += nn * static_cast<unsigned long long>(n_old + hd.n_ext);
                const unsigned long long tiles = (nn + p.group_size - 1) / p.group_size;
                const bool redo = p.np_enabled && p.halo_redo[h];
                const bool was_incr = p.incr_enabled && p.incr[h] && !redo;
                // externals: evaluated unless their first-pass sum was reused (cache) or kept (incremental)
                const bool ext_reused = p.cache_ext && p.st->pass > 0 && p.ext_ok[h] && !redo;
                const unsigned long long ext_ev = ext_reused ? 0ull : nn * static_cast<unsigned long long>(hd.n_ext);
                if (was_incr) {
                    // survivors x the members the previous pass removed
                    p.evals[h] += nn * static_cast<unsigned long long>(p.rem_cnt[h]);
                } else if (p.sym_enabled && tiles >= 2 && !redo) {
                    // diagonal tiles one-sided, every other member pair once
                    const unsigned long long last = nn - (tiles - 1) * p.group_size;
                    const unsigned long long diag = (tiles - 1) * p.group_size * p.group_size + last * last;
                    p.evals[h] += ext_ev + (nn * nn + diag) / 2;
                } else {
                    p.evals[h] += nn * nn + ext_ev;
                }
                // a first pass that fell back to the predicated kernel leaves no usable external sums
                if (p.cache_ext && p.st->pass == 0 && redo) p.ext_ok[h] = 0;
                if (p.incr_enabled) {
                    // The coming pass is incremental when this one left a valid potential behind (k_energy_flag,
                    // phi_keep) and removed at most a third of the members: survivors x removed is then cheaper
                    // than a full pass even with the symmetric self-term.
                    const int n_rem = n_old - n_new;
                    p.rem_cnt[h] = n_rem;
                    p.incr[h] = (!redo && n_rem > 0 && 2ll * n_rem <= n_new) ? 1 : 0;
                }
                const int changed = n_new != n_old;
                p.converged[h] = (!changed || n_new == 0) ? 1 : 0;
                act = (changed && n_new > 0 && it < p.max_iter) ? 1 : 0;
            }
            p.active_next[h] = act;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Kernel 3: stable stream compaction (ballot + popc inside a warp, warp offsets through
// shared memory, chunk offsets from k_halo_decide).  Order-preserving, so the member
// indices stay ascending like part_list[bound] (halo_properties.py:359-361).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCh) k_compact(const LoopParams p)
{
    __shared__ int woff[kCh / 32];
    if (!p.st->any_active) return;
    const int par = p.st->parity, nxt = par ^ 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        const int p0 = p.chunk_p0[c];
        if (p0 >= n) continue;
        const HaloDesc &hd = p.halo[h];
        const int q = p0 + threadIdx.x;
        const int64_t i = hd.poff + q;
        const int f = (q < n) ? p.flag[i] : 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, f);
        const int rank_in_warp = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) woff[warp] = __popc(ballot);
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += woff[w];
        if (p.np_enabled && q < n && !f) {
            // a removed member stops being a source of the correction tickets
            const int64_t slot = hd.poff + (p.widx[par][i] - hd.uoff);
#pragma unroll
            for (int a = 0; a < 3; ++a) p.ax[a].m[p.ax[a].inv[slot]] = 0.f;
        }
        if (p.incr_enabled && p.incr[h] && q < n && !f) {
            // the coming pass is incremental: keep the removed members, in order, as its sources
            const int64_t d = hd.poff + (q - (p.chunk_off[c] + base + rank_in_warp));
            p.rx[d] = p.wx[par][i];
            p.ry[d] = p.wy[par][i];
            p.rz[d] = p.wz[par][i];
            p.rm[d] = p.wm[par][i];
        }
        if (f) {
            const int64_t d = hd.poff + p.chunk_off[c] + base + rank_in_warp;
            p.wx[nxt][d] = p.wx[par][i];
            p.wy[nxt][d] = p.wy[par][i];
            p.wz[nxt][d] = p.wz[par][i];
            p.wm[nxt][d] = p.wm[par][i];
            p.widx[nxt][d] = p.widx[par][i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Commit per-halo state and build the ticket table of the next potential pass.
// Single block.  Tickets are laid out in `order` (largest halo first).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_schedule(const LoopParams p, int init)
{
    __shared__ int sred[33];
    __shared__ long long spair[32];
    __shared__ int s_total_groups, s_carry, s_any, s_sym_chunk;
    LoopState *st = p.st;
    if (!init && !st->any_active) {
        // graph driver: nothing left to do, leave the WHILE node
        if (threadIdx.x == 0 && p.cond_handle) cudaGraphSetConditional(p.cond_handle, 0u);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // commit
    int groups = 0, any = 0;
    long long tile_pairs = 0;          // symmetric tickets: off-diagonal tile pairs of the whole plan
    for (int h = threadIdx.x; h < p.n_halo; h += blockDim.x) {
        if (init || p.active[h]) {
            // the halo took part in this pass: its members now live in the other buffer
            if (!init) p.halo_buf[h] = st->parity ^ 1;
            p.cnt[h] = p.cnt_next[h];
            p.active[h] = p.active_next[h];
            for (int k = 0; k < 3; ++k) p.hvb[3 * h + k] = p.hvb_next[3 * h + k];
        }
        if (init) p.halo_buf[h] = 0;
        if (p.np_enabled) p.halo_redo[h] = 0;
        if (p.active[h]) {
            // all ranks' groups: the j-split must depend on the problem only, so that a split
            // run sums its partial potentials in the same grouping as a single-GPU run
            const long long tiles = (p.cnt[h] + p.group_size - 1) / p.group_size;
            groups += static_cast<int>(tiles);
            if (!(p.incr_enabled && !init && p.incr[h])) tile_pairs += tiles * (tiles - 1) / 2;
            any = 1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        groups += __shfl_down_sync(0xffffffffu, groups, o);
        tile_pairs += __shfl_down_sync(0xffffffffu, tile_pairs, o);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
        sred[warp] = groups;
        spair[warp] = tile_pairs;
    }
    if (threadIdx.x == 0) {
        s_any = 0;
        s_carry = 0;
    }
    __syncthreads();
    if (lane == 0 && any) atomicOr(&s_any, 1);
    if (threadIdx.x == 0) {
        int t = 0;
        long long tp = 0;
        for (int w = 0; w < 32; ++w) {
            t += sred[w];
            tp += spair[w];
        }
        s_total_groups = t;
        // column tiles per symmetric ticket: about kNominalTickets tickets over the whole plan, between
        // 2 and 32 -- a lone mid-size halo gets short tickets that fill the machine, a catalogue or a giant
        // halo long ones that amortise the per-ticket work.  Depends on the plan only, not on the GPU.
        const long long c = tp / kNominalTickets;
        s_sym_chunk = c < 2 ? 2 : (c > 32 ? 32 : static_cast<int>(c));
        st->sym_chunk = s_sym_chunk;
    }
    __syncthreads();
    const int total_groups = s_total_groups;
    const int sym_chunk = s_sym_chunk;
    int want = 1;
    if (p.mode == HALMA_MODE_FAST && total_groups > 0 && total_groups < p.target_items)
        want = (p.target_items + total_groups - 1) / total_groups;

    // ticket counts in `order` space, tile-wise block scan with a running carry
    int max_split = 1;
    for (int k0 = 0; k0 < p.n_halo; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int items = 0, h = -1;
        if (k < p.n_halo) {
            h = p.order[k];
            if (p.active[h]) {
                const int n = p.cnt[h];
                // sources a main ticket of the coming pass streams (potential.cu): the removed members in an
                // incremental pass, else the members plus the externals unless their sum is cached
                const bool inc = p.incr_enabled && !init && p.incr[h];
                const bool ext_cached = p.cache_ext && !init && p.ext_ok[h];
                const int n_src = inc ? p.rem_cnt[h] : n + (ext_cached ? 0 : p.halo[h].n_ext);
                int S = min(want, min(p.max_split, max(1, n_src / kMinSplitSources)));
                p.nsplit[h] = S;
                max_split = max(max_split, S);
                items = my_groups(n, p.group_size, p.rank, p.n_ranks) * S;
                // correction tickets: three axes x blocks of the (static) sorted member list
                if (p.np_enabled && !inc) items += 3 * my_groups(p.halo[h].n0, p.group_size, p.rank, p.n_ranks);
                // symmetric tickets: row tiles x chunks of column tiles (potential.cu::decode_ticket)
                if (p.sym_enabled && !inc) {
                    const int tiles = (n + p.group_size - 1) / p.group_size;
                    if (tiles >= 2) items += tiles * ((tiles - 1 + sym_chunk - 1) / sym_chunk);
                }
            }
        }
        // inclusive warp scan
        int incl = items;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        __syncthreads();
        if (lane == 31) sred[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = sred[lane];
            int wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += y;
            }
            sred[lane] = wi - w;
            if (lane == 31) sred[32] = wi;
        }
        __syncthreads();
        const int carry = s_carry;
        if (k < p.n_halo) p.item_base[k] = carry + sred[warp] + incl - items;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + sred[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.item_base[p.n_halo] = s_carry;
        st->n_items = s_carry;
        st->any_active = s_any;
        st->counter = 0u;
        st->counter_redo = 0u;
        st->redo_any = 0;
        if (init) {
            st->parity = 0;
            st->pass = 0;
        } else {
            st->parity ^= 1;
            st->pass += 1;
        }
        // graph driver: run another pass of the WHILE body iff some halo is still active
        if (p.cond_handle) cudaGraphSetConditional(p.cond_handle, s_any ? 1u : 0u);
    }
    for (int o = 16; o > 0; o >>= 1) max_split = max(max_split, __shfl_down_sync(0xffffffffu, max_split, o));
    if (lane == 0 && max_split > 1) atomicMax(&st->n_split, max_split);
}

// ---------------------------------------------------------------------------------------
// Split mode: fold the j-split partials into plane 0 before the all-reduce, and clear the
// entries this rank does not own (they are filled by the reduction with exact zeros added).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fold_partials(const LoopParams p)
{
    if (!p.st->any_active) return;
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        const int q = p.chunk_p0[c] + threadIdx.x;
        if (q >= n) continue;
        const int64_t i = p.halo[h].poff + q;
        const int owner = (q / p.group_size) % p.n_ranks;
        double phi = 0.0;
        if (owner == p.rank) {
            const int S = p.nsplit[h];
            phi = p.phi_part[i];
            for (int k = 1; k < S; ++k) phi += p.phi_part[static_cast<int64_t>(k) * p.n_pad + i];
        }
        p.phi_part[i] = phi;
    }
}

// Split mode, predicate-free path: after the max-all-reduce of the per-halo fallback flags
// every rank derives the same redo_any.
__global__ void k_sync_redo(const LoopParams p)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int any = 0;
        for (int h = 0; h < p.n_halo; ++h) any |= p.halo_redo[h];
        p.st->redo_any = any;
    }
}

// After the all-reduce plane 0 holds the complete Phi: tell k_energy_flag to read only it.
__global__ void k_set_nsplit_one(const LoopParams p)
{
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < p.n_halo; h += gridDim.x * blockDim.x)
        p.nsplit[h] = 1;
}

// ---------------------------------------------------------------------------------------
// Final member lists: ascending local indices of the bound members of each halo.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCh) k_finalize(const LoopParams p)
{
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        const HaloDesc &hd = p.halo[h];
        const int q = p.chunk_p0[c] + threadIdx.x;
        if (q >= hd.n0) continue;
        const int n = p.cnt[h];
        const int b = p.halo_buf[h];
        p.out_idx[hd.uoff + q] = (q < n) ? static_cast<int32_t>(p.widx[b][hd.poff + q] - hd.uoff) : -1;
    }
}

// ---------------------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------------------
static inline int chunk_grid(const LoopParams &p, int sm_count)
{
    const int want = sm_count * 8;
    return p.n_chunks < want ? (p.n_chunks > 0 ? p.n_chunks : 1) : want;
}

cudaError_t launch_pack_members(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_pack_members<<<chunk_grid(p, sm_count), kCh, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_pack_group(const HaloDesc *halo, int n_halo, int seg_index, int max_count, const int64_t *ext_off,
                              const double *m, const double *x, const double *y, const double *z, float *em,
                              float *ex, float *ey, float *ez, cudaStream_t s)
{
    int gx = (max_count + 255) / 256;
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    int gy = n_halo < 1 ? 1 : (n_halo > 4096 ? 4096 : n_halo);
    if (gx * gy > 65536) gx = 65536 / gy > 0 ? 65536 / gy : 1;
    k_pack_group<<<dim3(gx, gy), 256, 0, s>>>(halo, n_halo, seg_index, ext_off, m, x, y, z, em, ex, ey, ez);
    return cudaGetLastError();
}

cudaError_t launch_energy_flag(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_energy_flag<<<chunk_grid(p, sm_count), kCh, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_halo_extent(const LoopParams &p, cudaStream_t s)
{
    const int g = p.n_halo < 4096 ? (p.n_halo > 0 ? p.n_halo : 1) : 4096;
    k_halo_extent<<<g, kCh, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_halo_decide(const LoopParams &p, int init, int sm_count, cudaStream_t s)
{
    const int want = sm_count * 8;
    const int g = p.n_halo < want ? (p.n_halo > 0 ? p.n_halo : 1) : want;
    k_halo_decide<<<g, kCh, 0, s>>>(p, init);
    return cudaGetLastError();
}

cudaError_t launch_compact(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_compact<<<chunk_grid(p, sm_count), kCh, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_schedule(const LoopParams &p, int init, cudaStream_t s)
{
    k_schedule<<<1, 1024, 0, s>>>(p, init);
    return cudaGetLastError();
}

cudaError_t launch_fold_partials(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_fold_partials<<<chunk_grid(p, sm_count), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_sync_redo(const LoopParams &p, cudaStream_t s)
{
    k_sync_redo<<<1, 32, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_set_nsplit_one(const LoopParams &p, cudaStream_t s)
{
    const int g = (p.n_halo + 255) / 256;
    k_set_nsplit_one<<<g < 1 ? 1 : g, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_finalize<<<chunk_grid(p, sm_count), kCh, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace halma
