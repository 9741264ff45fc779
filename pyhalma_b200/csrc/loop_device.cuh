// Device side of kernels 2 and 3 of the unbinding path and of the device-side scheduler, sm_100a.
//
// Every phase is a block-wide device function over 128-thread blocks, used both by the stand-alone
// kernels of the multi-launch drivers (loop_kernels.cu) and by the persistent loop kernel (fused.cu),
// so that all drivers produce the same bits:
//   energy_phase    per 256-member chunk: energy step + bound flag + survivor count and mass sums
//                   (halo_properties.py:342-359 / halo_gas.py:456-476; sums :16-60); the block that
//                   finishes the LAST chunk of a halo also takes the halo's decision (decide_halo):
//                   scan of chunk counts, reduction of chunk sums -> new count, M, CoM, bulk velocity,
//                   converged / active, the kind of the coming pass, and the halo's scheduling record
//   compact_phase   stable (order-preserving) warp-aggregated stream compaction of the float32
//                   working set into the other buffer; inside the persistent kernel a chunk waits
//                   for its halo's decision through a per-halo stamp instead of a kernel boundary
//   commit_phase    per halo: next-pass state becomes current
//   schedule_block  one block: ticket table of the next potential pass from the scheduling records
// All are O(N) and HBM-bound; the potential kernel dominates for N >~ 1e3.
//
// Reductions are done in a fixed order (two members per thread, tree inside a warp, ascending
// warps, ascending chunks), so a run is bit-reproducible and independent of the driver.
#pragma once
#include "halma_common.cuh"
#include "loop_kernels.h"

namespace halma {

constexpr int kLT = 128;                 // threads per block of every loop phase
constexpr int kLW = kLT / 32;
constexpr int kRounds = kChunk / kLT;    // members per thread and chunk
static_assert(kRounds * kLT == kChunk, "a chunk is a whole number of rounds");

struct LoopSmem {
    double red[kChunkSums * kLW];
    float best[kLW];
    int best_q[kLW];
    int scan[kLW + 1];
    int woff[kRounds * kLW];
    int bcast;
    long long lred[kLW];
    float fmn[3][kLW], fmx[3][kLW];
};

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Generic-proxy writes to global memory become visible to later TMA (async proxy) reads.
__device__ __forceinline__ void fence_proxy_async_global()
{
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// float <-> int whose signed order is the float order (for atomicMin / atomicMax)
__device__ __forceinline__ int float_order(float f)
{
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

// Deterministic block-wide sum of NV doubles per thread.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* [NV * kLW] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) smem[k * kLW + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (int w = 0; w < kLW; ++w) s += smem[k * kLW + w];
            v[k] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem /* [kLW + 1] */, int &total)
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
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < kLW; ++w) {
            const int t = smem[w];
            smem[w] = run;
            run += t;
        }
        smem[kLW] = run;
    }
    __syncthreads();
    const int r = smem[warp] + incl - v;
    total = smem[kLW];
    __syncthreads();
    return r;
}

// (value, index) arg-max with "largest value, lowest index on ties" (halo_gas.py:627-632)
__device__ __forceinline__ void best_merge(float &best, int &best_q, float ob, int oq)
{
    if (ob > best || (ob == best && oq >= 0 && (best_q < 0 || oq < best_q))) {
        best = ob;
        best_q = oq;
    }
}
__device__ __forceinline__ void block_best(float &best, int &best_q, LoopSmem &sm)
{
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oq = __shfl_down_sync(0xffffffffu, best_q, o);
        best_merge(best, best_q, ob, oq);
    }
    if ((threadIdx.x & 31) == 0) {
        sm.best[threadIdx.x >> 5] = best;
        sm.best_q[threadIdx.x >> 5] = best_q;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < kLW; ++w) best_merge(best, best_q, sm.best[w], sm.best_q[w]);
    __syncthreads();
}

// Groups of `gs` targets of a halo with n members that belong to this rank (split mode).
__device__ __forceinline__ int my_groups(int n, int gs, int rank, int n_ranks)
{
    const int groups = (n + gs - 1) / gs;
    return (groups - rank + n_ranks - 1) / n_ranks;
}

// ---------------------------------------------------------------------------------------
// Pack: float64 user arrays -> float32 working set (round to nearest, like np.float32()), the
// initial chunk sums, and (symmetric mode) the coordinate range of every halo.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_chunk(const LoopParams &p, LoopSmem &sm, int c)
{
    const int h = p.chunk_halo[c];
    const HaloDesc &hd = p.halo[h];
    const int p0 = p.chunk_p0[c];
    double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float inf = __int_as_float(0x7f800000);
    float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
    double in[kRounds][7];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {          // every load before the first store (see energy_chunk)
        const int q = p0 + r * kLT + threadIdx.x;
        const int64_t g = hd.uoff + q;
        const bool ok = q < hd.n0;
        in[r][0] = ok ? p.x64[g] : 0.0;
        in[r][1] = ok ? p.y64[g] : 0.0;
        in[r][2] = ok ? p.z64[g] : 0.0;
        in[r][3] = ok ? p.m64[g] : 0.0;
        in[r][4] = ok ? p.vx[g] : 0.0;
        in[r][5] = ok ? p.vy[g] : 0.0;
        in[r][6] = ok ? p.vz[g] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int q = p0 + r * kLT + threadIdx.x;
        if (q < hd.n0) {
            const int64_t i = hd.poff + q, g = hd.uoff + q;
            const double x = in[r][0], y = in[r][1], z = in[r][2], m = in[r][3];
            const float xf = __double2float_rn(x), yf = __double2float_rn(y), zf = __double2float_rn(z);
            p.wx[0][i] = xf;
            p.wy[0][i] = yf;
            p.wz[0][i] = zf;
            p.wm[0][i] = __double2float_rn(m);
            p.widx[0][i] = static_cast<int32_t>(g);
            s[0] += m;
            s[1] += m * in[r][4];
            s[2] += m * in[r][5];
            s[3] += m * in[r][6];
            s[4] += m * x;
            s[5] += m * y;
            s[6] += m * z;
            const float v[3] = {xf, yf, zf};
#pragma unroll
            for (int a = 0; a < 3; ++a) {          // fminf / fmaxf skip NaN coordinates
                mn[a] = fminf(mn[a], v[a]);
                mx[a] = fmaxf(mx[a], v[a]);
            }
        }
    }
    block_sum<kChunkSums>(s, sm.red);
    if (p.sym_enabled) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            for (int o = 16; o > 0; o >>= 1) {
                mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], o));
            }
            if ((threadIdx.x & 31) == 0) {
                sm.fmn[a][threadIdx.x >> 5] = mn[a];
                sm.fmx[a][threadIdx.x >> 5] = mx[a];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.chunk_cnt[c] = min(kChunk, max(hd.n0 - p0, 0));
#pragma unroll
        for (int k = 0; k < kChunkSums; ++k) p.chunk_sum[static_cast<int64_t>(c) * kChunkSums + k] = s[k];
        p.chunk_best[c] = -1.f;
        p.chunk_best_q[c] = -1;
        if (p.sym_enabled) {
            for (int a = 0; a < 3; ++a) {
                float lo = sm.fmn[a][0], hi = sm.fmx[a][0];
                for (int w = 1; w < kLW; ++w) {
                    lo = fminf(lo, sm.fmn[a][w]);
                    hi = fmaxf(hi, sm.fmx[a][w]);
                }
                // min / max are exact, so the order of the atomics does not matter
                if (lo <= hi) {
                    atomicMin(&p.halo_rmin[3 * h + a], float_order(lo));
                    atomicMax(&p.halo_rmax[3 * h + a], float_order(hi));
                }
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Per halo: exclusive scan of chunk counts, ordered reduction of the chunk sums, the convergence
// decision and the scheduling record of the coming pass.  init = 1 right after the pack (no pass
// made yet).  Chunk results may come from other blocks of the same launch: read through L2.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void decide_halo(const LoopParams &p, LoopSmem &sm, int h, int init, int par, int pass)
{
    const HaloDesc &hd = p.halo[h];
    const int n_old = init ? hd.n0 : p.cnt[h];
    const int nch = (n_old + kChunk - 1) / kChunk;
    int carry = 0;
    double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float best = -1.f;          // most bound member of this pass (largest potential, lowest index)
    int best_q = -1;
    for (int c0 = 0; c0 < nch; c0 += kLT) {
        const int c = c0 + threadIdx.x;
        const int v = (c < nch) ? __ldcg(&p.chunk_cnt[hd.chunk_begin + c]) : 0;
        int total;
        const int ex = block_exclusive_scan(v, sm.scan, total);
        if (c < nch) {
            p.chunk_off[hd.chunk_begin + c] = carry + ex;
            const double *cs = p.chunk_sum + static_cast<int64_t>(hd.chunk_begin + c) * kChunkSums;
#pragma unroll
            for (int k = 0; k < kChunkSums; ++k) s[k] += __ldcg(&cs[k]);
            best_merge(best, best_q, __ldcg(&p.chunk_best[hd.chunk_begin + c]),
                       __ldcg(&p.chunk_best_q[hd.chunk_begin + c]));
        }
        carry += total;
    }
    block_best(best, best_q, sm);
    block_sum<kChunkSums>(s, sm.red);
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
            if (init) {
                // largest coordinate extent of the halo's members (float32 working set, from the pack)
                double ext = 0.0;
                for (int a = 0; a < 3; ++a) {
                    const float lo = order_float(p.halo_rmin[3 * h + a]);
                    const float hi = order_float(p.halo_rmax[3 * h + a]);
                    ext = fmax(ext, static_cast<double>(hi) - static_cast<double>(lo));
                }
                p.sym_ext[h] = ext;
            }
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
        // halo_properties.py:39-43, 56-60: sums divided by M, zeros when M == 0
        for (int k = 0; k < 3; ++k) {
            p.hcom[3 * h + k] = M > 0.0 ? s[4 + k] / M : 0.0;
            p.hvb_next[3 * h + k] = p.vb_fixed ? p.hvb[3 * h + k] : (M > 0.0 ? s[1 + k] / M : 0.0);
        }
        p.cnt_next[h] = n_new;
        int act, inc_next = 0;
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
            const bool redo = p.redo_enabled && p.halo_redo[h];
            const bool was_incr = p.incr_enabled && p.incr[h] && !redo;
            // externals: evaluated unless their first-pass sum was reused (cache) or kept (incremental)
            const bool ext_reused = p.cache_ext && pass > 0 && p.ext_ok[h] && !redo;
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
            if (p.cache_ext && pass == 0 && redo) p.ext_ok[h] = 0;
            if (p.incr_enabled) {
                // The coming pass is incremental when this one left a valid potential behind (energy_phase,
                // phi_keep) and removed at most a third of the members: survivors x removed is then cheaper
                // than a full pass even with the symmetric self-term.  Chains of incremental passes are cut
                // by energy_phase, which compares the kept potential with the last fully evaluated one.
                const int n_rem = n_old - n_new;
                p.rem_cnt[h] = n_rem;
                inc_next = (!redo && n_rem > 0 && 2ll * n_rem <= n_new) ? 1 : 0;
                p.incr[h] = inc_next;
            }
            const int changed = n_new != n_old;
            p.converged[h] = (!changed || n_new == 0) ? 1 : 0;
            act = (changed && n_new > 0 && it < p.max_iter) ? 1 : 0;
        }
        p.active_next[h] = act;
        // the halo took part in this pass: its members now live in the other buffer (read by finalize_phase only)
        if (!init) p.halo_buf[h] = par ^ 1;
        // Scheduling record of the coming pass, in `order` space so that schedule_block reads it coalesced:
        // members (0: no pass), sources a main ticket streams, incremental?, original member count.
        const bool ext_cached = p.cache_ext && !init && p.ext_ok[h];
        const int n_src = inc_next ? p.rem_cnt[h] : n_new + (ext_cached ? 0 : hd.n_ext);
        p.sched[p.rank_of[h]] = make_int4(act ? n_new : 0, n_src, inc_next, hd.n0);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Kernel 2: energy step, bound flag, survivor counts and mass sums of one chunk.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void energy_chunk(const LoopParams &p, LoopSmem &sm, int c, int h, int par, int pass)
{
    const int n = p.cnt[h];
    const int p0 = p.chunk_p0[c];
    const HaloDesc &hd = p.halo[h];
    double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float best = -1.f;          // potentials are >= 0; NaN never wins
    int best_q = -1;
    const int S = p.nsplit[p.rank_of[h]];
    // a halo handed to the predicated kernel has its complete, final sum in the planes
    const bool redo = p.redo_enabled && p.halo_redo[h];
    const bool inc = p.incr_enabled && p.incr[h] && !redo;
    const bool ext_cached = p.cache_ext && hd.n_ext > 0 && p.ext_ok[h] && !redo && !inc;
    const bool corr = p.np_enabled && !redo && !inc;
    const double vb0 = p.hvb[3 * h + 0], vb1 = p.hvb[3 * h + 1], vb2 = p.hvb[3 * h + 2];
    // The phase is bound by memory latency, not bandwidth, unless many loads are in flight: all loads of both
    // members of a thread are issued before the first store (stores could alias, so the compiler would not
    // move a load across them).
    bool ok[kRounds];
    int64_t gi[kRounds];
    double phi[kRounds], vx[kRounds], vy[kRounds], vz[kRounds], mm[kRounds], tt[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int q = p0 + r * kLT + threadIdx.x;
        ok[r] = q < n;
        gi[r] = ok[r] ? p.widx[par][hd.poff + q] : hd.uoff;
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int64_t i = hd.poff + p0 + r * kLT + threadIdx.x;
        const int64_t g = gi[r];
        const int64_t slot = hd.poff + (g - hd.uoff);      // the member's original slot
        phi[r] = 0.0;
        vx[r] = vy[r] = vz[r] = mm[r] = tt[r] = 0.0;
        if (ok[r]) {
            // Phi: ascending sum of the j-split partials
            double ph = p.phi_part[i];
            for (int k = 1; k < S; ++k) ph += p.phi_part[static_cast<int64_t>(k) * p.n_pad + i];
            // the two-sided sums of this pass (the slot is cleared for the next one below)
            const double ps = p.sym_enabled ? p.phi_sym[i] : 0.0;
            if (inc) {
                // incremental pass: the planes hold what the members removed by the previous pass contributed
                // (reference predicate applied); take it out of the potential kept from that pass
                ph = p.phi_keep[slot] - ph;
            } else if (!redo) {
                if (p.sym_enabled) ph += ps;
                if (ext_cached) {
                    // sum over the external sources, evaluated by the first pass only (potential.cu)
                    double e = p.phi_ext[slot];
                    if (pass == 0)
                        for (int k = 1; k < S; ++k) e += p.phi_ext[static_cast<int64_t>(k) * p.n_pad + slot];
                    tt[r] = e;          // parked here until the stores below
                    ph += e;
                }
                if (corr) {
                    // predicate-free path: take out the pairs that share a coordinate (potential.cu)
                    ph -= (p.ax[0].corr[slot] + p.ax[1].corr[slot]) + p.ax[2].corr[slot];
                }
            }
            phi[r] = ph;
            vx[r] = p.vx[g];
            vy[r] = p.vy[g];
            vz[r] = p.vz[g];
            mm[r] = p.m64[g];
        }
    }
    int bound[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        bound[r] = 0;
        if (ok[r]) {
            const int64_t i = hd.poff + p0 + r * kLT + threadIdx.x;
            const int64_t g = gi[r];
            const int64_t slot = hd.poff + (g - hd.uoff);
            if (p.sym_enabled) p.phi_sym[i] = 0.0;
            if (ext_cached && pass == 0) p.phi_ext[slot] = tt[r];      // the folded sum, read by the later passes
            // the complete float64 potential, kept for a following incremental pass; phi_full remembers the
            // last one that was evaluated in full (the incremental tickets bound the share removed since)
            if (p.incr_enabled && !redo) {
                p.phi_keep[slot] = phi[r];
                if (!inc) p.phi_full[slot] = phi[r];
            }
            // rounded once to the f2py output dtype
            const float be = __double2float_rn(phi[r]);
            // halo_properties.py:342-351 / halo_gas.py:456-465: float32 chain, two roundings
            float pe = -be;
            pe = __fmul_rn(pe, p.G32);
            pe = __fmul_rn(pe, p.kappa32);
            // :354 / :468  float64, no contraction: 0.5*((dvx^2 + dvy^2) + dvz^2)
            const double dvx = __dsub_rn(vx[r], vb0);
            const double dvy = __dsub_rn(vy[r], vb1);
            const double dvz = __dsub_rn(vz[r], vb2);
            const double ke = __dmul_rn(
                0.5, __dadd_rn(__dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy)), __dmul_rn(dvz, dvz)));
            const double E = __dadd_rn(ke, static_cast<double>(pe));
            bound[r] = (E <= 0.0) ? 1 : 0;        // :359 / :476 (NaN is neither bound nor unbound)
            p.flag[i] = static_cast<uint8_t>(bound[r]);
            p.out_mask[g] = static_cast<uint8_t>(bound[r]);
            p.out_be[g] = be;
            p.out_E[g] = E;
            if (be > best) {                      // q ascends with r, so the lowest index wins ties
                best = be;
                best_q = static_cast<int>(g - hd.uoff);
            }
            const double m = mm[r];
            if (p.temp) {
                // halo_gas.py:479-490: cold = T < 5e4, hot = T >= 5e4 (NaN is neither)
                const double T = p.temp[g];
                const bool cold = T < p.cold_T, hot = T >= p.cold_T;
                if (bound[r] && cold) s[7] += m;
                if (E > 0.0 && cold) s[8] += m;
                if (E > 0.0 && hot) s[9] += m;
            }
            if (bound[r]) {
                s[0] += m;
                s[1] += m * vx[r];
                s[2] += m * vy[r];
                s[3] += m * vz[r];
                s[4] += m * p.x64[g];
                s[5] += m * p.y64[g];
                s[6] += m * p.z64[g];
            }
        }
    }
    int count = 0;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) count += __syncthreads_count(bound[r]);
    block_sum<kChunkSums>(s, sm.red);
    block_best(best, best_q, sm);
    if (threadIdx.x == 0) {
        p.chunk_cnt[c] = count;
#pragma unroll
        for (int k = 0; k < kChunkSums; ++k) p.chunk_sum[static_cast<int64_t>(c) * kChunkSums + k] = s[k];
        p.chunk_best[c] = best;
        p.chunk_best_q[c] = best_q;
    }
}

// Energy step of every chunk this block owns; the block that completes a halo decides it.
__device__ __forceinline__ void energy_phase(const LoopParams &p, LoopSmem &sm, int par, int pass)
{
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        if (p.chunk_p0[c] >= n) continue;
        energy_chunk(p, sm, c, h, par, pass);
        if (threadIdx.x == 0) {
            const int nch = (n + kChunk - 1) / kChunk;
            int last = 1;
            if (nch > 1) {
                __threadfence();                                   // this chunk's results before the count
                last = atomicAdd(&p.halo_done[h], 1) + 1 == nch;
                if (last) {
                    p.halo_done[h] = 0;
                    __threadfence();                               // the other chunks' results after it
                }
            }
            sm.bcast = last;
        }
        __syncthreads();
        const int last = sm.bcast;
        __syncthreads();
        if (last) {
            decide_halo(p, sm, h, 0, par, pass);
            if (threadIdx.x == 0) {
                __threadfence();
                st_release(&p.halo_stamp[h], pass + 1);            // compact_phase may go ahead with this halo
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel 3: stable stream compaction (ballot + popc inside a warp, warp offsets through
// shared memory, chunk offsets from decide_halo).  Order-preserving, so the member
// indices stay ascending like part_list[bound] (halo_properties.py:359-361).
// wait: inside the persistent kernel, spin until the halo's decision of this pass is published.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void compact_phase(const LoopParams &p, LoopSmem &sm, int par, int pass, bool wait)
{
    const int nxt = par ^ 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        const int p0 = p.chunk_p0[c];
        if (p0 >= n) continue;
        if (wait) {
            if (threadIdx.x == 0)
                while (ld_acquire(&p.halo_stamp[h]) != pass + 1) __nanosleep(64);
            __syncthreads();
        }
        const HaloDesc &hd = p.halo[h];
        const int coff = __ldcg(&p.chunk_off[c]);
        const int inc_next = p.incr_enabled ? __ldcg(&p.incr[h]) : 0;
        int f[kRounds], rk[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int q = p0 + r * kLT + threadIdx.x;
            f[r] = (q < n) ? p.flag[hd.poff + q] : 0;
            const unsigned ballot = __ballot_sync(0xffffffffu, f[r]);
            rk[r] = __popc(ballot & ((1u << lane) - 1u));
            if (lane == 0) sm.woff[r * kLW + warp] = __popc(ballot);
        }
        __syncthreads();
        float vals[kRounds][4];
        int32_t wid[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {          // every load before the first store (see energy_chunk)
            const int q = p0 + r * kLT + threadIdx.x;
            const int64_t i = hd.poff + min(q, n - 1);
            vals[r][0] = p.wx[par][i];
            vals[r][1] = p.wy[par][i];
            vals[r][2] = p.wz[par][i];
            vals[r][3] = p.wm[par][i];
            wid[r] = p.widx[par][i];
        }
        int32_t inv[kRounds][3];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int q = p0 + r * kLT + threadIdx.x;
            const int64_t slot = hd.poff + (wid[r] - hd.uoff);
#pragma unroll
            for (int a = 0; a < 3; ++a) inv[r][a] = (p.np_enabled && q < n && !f[r]) ? p.ax[a].inv[slot] : 0;
        }
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int q = p0 + r * kLT + threadIdx.x;
            if (q >= n) continue;
            int base = 0;
            for (int w = 0; w < r * kLW + warp; ++w) base += sm.woff[w];
            const int dst = coff + base + rk[r];                  // survivors before this member, whole halo
            if (!f[r]) {
                if (p.np_enabled) {
                    // a removed member stops being a source of the correction tickets
#pragma unroll
                    for (int a = 0; a < 3; ++a) p.ax[a].m[inv[r][a]] = 0.f;
                }
                if (inc_next) {
                    // the coming pass is incremental: keep the removed members, in order, as its sources
                    const int64_t d = hd.poff + (q - dst);
                    p.rx[d] = vals[r][0];
                    p.ry[d] = vals[r][1];
                    p.rz[d] = vals[r][2];
                    p.rm[d] = vals[r][3];
                }
            } else {
                const int64_t d = hd.poff + dst;
                p.wx[nxt][d] = vals[r][0];
                p.wy[nxt][d] = vals[r][1];
                p.wz[nxt][d] = vals[r][2];
                p.wm[nxt][d] = vals[r][3];
                p.widx[nxt][d] = wid[r];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Per halo: the state decided by the pass becomes current.  Grid-stride over haloes.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void commit_phase(const LoopParams &p, int init)
{
    for (int h = blockIdx.x * kLT + threadIdx.x; h < p.n_halo; h += gridDim.x * kLT) {
        if (init || p.active[h]) {
            p.cnt[h] = p.cnt_next[h];
            p.active[h] = p.active_next[h];
            for (int k = 0; k < 3; ++k) p.hvb[3 * h + k] = p.hvb_next[3 * h + k];
        }
        if (init) p.halo_buf[h] = 0;
        if (p.redo_enabled) p.halo_redo[h] = 0;
    }
}

// Tickets of one halo in the coming pass (must match potential_device.cuh::decode_ticket).
__device__ __forceinline__ int sched_items(const LoopParams &p, const int4 rec, int want, int sym_chunk, int &S)
{
    const int n = rec.x;
    S = 1;
    if (n <= 0) return 0;
    const bool inc = rec.z != 0;
    S = min(want, min(p.max_split, max(1, rec.y / kMinSplitSources)));
    int items = my_groups(n, p.group_size, p.rank, p.n_ranks) * S;
    // correction tickets: three axes x blocks of the (static) sorted member list
    if (p.np_enabled && !inc) items += 3 * my_groups(rec.w, p.group_size, p.rank, p.n_ranks);
    // symmetric tickets: row tiles x chunks of column tiles
    if (p.sym_enabled && !inc) {
        const int tiles = (n + p.group_size - 1) / p.group_size;
        if (tiles >= 2) items += tiles * ((tiles - 1 + sym_chunk - 1) / sym_chunk);
    }
    return items;
}

// ---------------------------------------------------------------------------------------
// Ticket table of the next potential pass, from the scheduling records.  ONE block.
// Tickets are laid out in `order` (largest halo first).  Does not touch the per-halo state the
// other phases read, so it may run next to commit_phase.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void schedule_block(const LoopParams &p, LoopSmem &sm, int init)
{
    LoopState *st = p.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // all ranks' groups: the j-split must depend on the problem only, so that a split
    // run sums its partial potentials in the same grouping as a single-GPU run
    int groups = 0, any = 0;
    long long tile_pairs = 0;          // symmetric tickets: off-diagonal tile pairs of the whole plan
    for (int k = threadIdx.x; k < p.n_halo; k += kLT) {
        const int4 rec = p.sched[k];
        if (rec.x > 0) {
            const long long tiles = (rec.x + p.group_size - 1) / p.group_size;
            groups += static_cast<int>(tiles);
            if (!rec.z) tile_pairs += tiles * (tiles - 1) / 2;
            any = 1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        groups += __shfl_down_sync(0xffffffffu, groups, o);
        tile_pairs += __shfl_down_sync(0xffffffffu, tile_pairs, o);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
        sm.scan[warp] = groups;
        sm.lred[warp] = tile_pairs;
        sm.woff[warp] = any;
    }
    __syncthreads();
    int total_groups = 0;
    long long tp = 0;
    any = 0;
    for (int w = 0; w < kLW; ++w) {
        total_groups += sm.scan[w];
        tp += sm.lred[w];
        any |= sm.woff[w];
    }
    __syncthreads();
    // column tiles per symmetric ticket: about kNominalTickets tickets over the whole plan, between
    // 2 and 32 -- a lone mid-size halo gets short tickets that fill the machine, a catalogue or a giant
    // halo long ones that amortise the per-ticket work.  Depends on the plan only, not on the GPU.
    const long long cc = tp / kNominalTickets;
    const int sym_chunk = cc < 2 ? 2 : (cc > 32 ? 32 : static_cast<int>(cc));
    int want = 1;
    if (p.mode == HALMA_MODE_FAST && total_groups > 0 && total_groups < p.target_items)
        want = (p.target_items + total_groups - 1) / total_groups;

    // every thread takes a contiguous run of haloes (in `order`): sum, block scan, then the prefix
    const int per = (p.n_halo + kLT - 1) / kLT;
    const int k0 = min(threadIdx.x * per, p.n_halo), k1 = min(k0 + per, p.n_halo);
    int mine = 0, max_split = 1;
    for (int k = k0; k < k1; ++k) {
        int S;
        mine += sched_items(p, p.sched[k], want, sym_chunk, S);
    }
    int total;
    int run = block_exclusive_scan(mine, sm.scan, total);
    for (int k = k0; k < k1; ++k) {
        int S;
        const int items = sched_items(p, p.sched[k], want, sym_chunk, S);
        p.item_base[k] = run;
        p.nsplit[k] = S;
        max_split = max(max_split, S);
        run += items;
    }
    if (threadIdx.x == 0) {
        p.item_base[p.n_halo] = total;
        st->n_items = total;
        st->any_active = any;
        st->counter = 0u;
        st->counter_redo = 0u;
        st->redo_any = 0;
        st->sym_chunk = sym_chunk;
        if (init) {
            st->parity = 0;
            st->pass = 0;
        } else {
            st->parity ^= 1;
            st->pass += 1;
        }
        // graph driver: run another pass of the WHILE body iff some halo is still active
        if (p.cond_handle) cudaGraphSetConditional(p.cond_handle, any ? 1u : 0u);
    }
    for (int o = 16; o > 0; o >>= 1) max_split = max(max_split, __shfl_down_sync(0xffffffffu, max_split, o));
    if (lane == 0 && max_split > 1) atomicMax(&st->n_split, max_split);
}

// ---------------------------------------------------------------------------------------
// Final member lists: ascending local indices of the bound members of each halo.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void finalize_phase(const LoopParams &p)
{
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        const HaloDesc &hd = p.halo[h];
        const int n = p.cnt[h];
        const int b = p.halo_buf[h];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int q = p.chunk_p0[c] + r * kLT + threadIdx.x;
            if (q < hd.n0)
                p.out_idx[hd.uoff + q] = (q < n) ? static_cast<int32_t>(p.widx[b][hd.poff + q] - hd.uoff) : -1;
        }
    }
}

}  // namespace halma
