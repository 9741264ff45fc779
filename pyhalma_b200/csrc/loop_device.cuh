// Device side of kernels 2 and 3 of the unbinding path and of the device-side scheduler, sm_100a.
//
// Every phase is a device function used both by the stand-alone kernels of the multi-launch drivers
// (loop_kernels.cu) and by the persistent loop kernel (fused.cu), so that all drivers produce the same bits.
// The unit of work is a WARP and a chunk of a halo (256 members, 8 per lane; 64 in small plans): no block barrier and
// no shared memory anywhere, so the warps of a block work on different chunks -- and different haloes --
// independently, like the ticket-taking warps of the potential kernel.
//   energy_phase    per chunk: energy step + bound flag + survivor count and mass sums
//                   (halo_properties.py:342-359 / halo_gas.py:456-476; sums :16-60); the warp that
//                   finishes the LAST chunk of a halo also takes the halo's decision (decide_halo):
//                   scan of chunk counts, reduction of chunk sums -> new count, M, CoM, bulk velocity,
//                   converged / active, the kind of the coming pass, and the halo's scheduling record
//   compact_phase   stable (order-preserving) warp-aggregated stream compaction of the float32
//                   working set into the other buffer; inside the persistent kernel a chunk waits
//                   for its halo's decision through a per-halo stamp instead of a kernel boundary
//   commit_phase    per halo (one thread each): next-pass state becomes current
//   schedule_block  one block: ticket table of the next potential pass from the scheduling records
// All are O(N) and HBM-bound; the potential kernel dominates for N >~ 1e3.  They are bound by memory
// LATENCY unless many loads are in flight, so the loads of two rounds (64 members per warp) are issued
// before the first store of either (stores could alias, the compiler would not move a load across them).
//
// Reductions are done in a fixed order (a lane's members in ascending order, tree across the lanes,
// ascending chunks), so a run is bit-reproducible and independent of the driver.
#pragma once
#include "halma_common.cuh"
#include "loop_kernels.h"

namespace halma {

constexpr int kLT = 128;                 // threads per block of every loop kernel (= the potential kernel's)
constexpr int kLW = kLT / 32;
// members per lane and chunk = p.chunk / 32: 8 (256-member chunks) or, for small plans, 2 (64-member chunks)
constexpr int kHoist = 1;                // rounds whose loads are in flight together (registers: the persistent
                                         // kernel holds these phases to the potential code's 80 registers)
constexpr int kIncPieceSources = 512;    // j-split piece of an incremental halo's removed members (sched_items)
constexpr int kHoistC = 2;               // ... in the compaction, which needs few registers per member
static_assert(kChunk % (32 * kHoist) == 0 && kChunkSmall % (32 * kHoist) == 0, "rounds come in groups");
static_assert(kChunk % (32 * kHoistC) == 0 && kChunkSmall % (32 * kHoistC) == 0, "rounds come in groups");

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

// Deterministic warp-wide sums of NV doubles per lane; the totals land in lane 0.
template <int NV>
__device__ __forceinline__ void warp_sum(double (&v)[NV])
{
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    }
}

// (value, index) arg-max with "largest value, lowest index on ties" (halo_gas.py:627-632)
__device__ __forceinline__ void best_merge(float &best, int &best_q, float ob, int oq)
{
    if (ob > best || (ob == best && oq >= 0 && (best_q < 0 || oq < best_q))) {
        best = ob;
        best_q = oq;
    }
}
__device__ __forceinline__ void warp_best(float &best, int &best_q)
{
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oq = __shfl_down_sync(0xffffffffu, best_q, o);
        best_merge(best, best_q, ob, oq);
    }
}

// Groups of `gs` targets of a halo with n members that belong to this rank (split mode).
__device__ __forceinline__ int my_groups(int n, int gs, int rank, int n_ranks)
{
    const int groups = (n + gs - 1) / gs;
    return (groups - rank + n_ranks - 1) / n_ranks;
}

// Results of one chunk, reduced over the warp (valid in lane 0).
struct ChunkSums {
    double s[kChunkSums];
    int count;
    float best;
    int best_q;
};

// ---------------------------------------------------------------------------------------
// Pack: float64 user arrays -> float32 working set (round to nearest, like np.float32()), the
// initial chunk sums, and (symmetric mode) the coordinate range of every halo.  One warp per chunk.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_chunk(const LoopParams &p, int c)
{
    const int lane = threadIdx.x & 31;
    const int h = p.chunk_halo[c];
    const HaloDesc &hd = p.halo[h];
    const int p0 = p.chunk_p0[c];
    double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float inf = __int_as_float(0x7f800000);
    float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
#pragma unroll 1
    for (int r0 = 0; r0 < (p.chunk >> 5); r0 += kHoist) {
        if (p0 + r0 * 32 >= hd.n0) break;
        double in[kHoist][7];
#pragma unroll
        for (int r = 0; r < kHoist; ++r) {
            const int q = p0 + (r0 + r) * 32 + lane;
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
        for (int r = 0; r < kHoist; ++r) {
            const int q = p0 + (r0 + r) * 32 + lane;
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
    }
    warp_sum<kChunkSums>(s);
    if (p.sym_enabled) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int o = 16; o > 0; o >>= 1) {
                mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], o));
            }
    }
    if (lane == 0) {
        p.chunk_cnt[c] = min(p.chunk, max(hd.n0 - p0, 0));
#pragma unroll
        for (int k = 0; k < kChunkSums; ++k) p.chunk_sum[static_cast<int64_t>(c) * kChunkSums + k] = s[k];
        p.chunk_best[c] = -1.f;
        p.chunk_best_q[c] = -1;
        if (p.sym_enabled) {
            for (int a = 0; a < 3; ++a)
                if (mn[a] <= mx[a]) {          // min / max are exact, so the order of the atomics does not matter
                    atomicMin(&p.halo_rmin[3 * h + a], float_order(mn[a]));
                    atomicMax(&p.halo_rmax[3 * h + a], float_order(mx[a]));
                }
        }
    }
}

__device__ __forceinline__ void pack_phase(const LoopParams &p)
{
    const int w = blockIdx.x * kLW + (threadIdx.x >> 5), nw = gridDim.x * kLW;
    for (int c = w; c < p.n_chunks; c += nw) pack_chunk(p, c);
}

// ---------------------------------------------------------------------------------------
// Per halo: exclusive scan of chunk counts, ordered reduction of the chunk sums, the convergence
// decision and the scheduling record of the coming pass.  One warp.  init = 1 right after the pack
// (no pass made yet).  `own`: the results of the halo's only chunk, still in lane 0's registers
// (haloes of up to 256 members, most of a catalogue); otherwise the chunk results are read back --
// through L2, they may come from other SMs in the same launch.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void decide_halo(const LoopParams &p, int h, int init, int par, int pass, const ChunkSums *own)
{
    const int lane = threadIdx.x & 31;
    const HaloDesc &hd = p.halo[h];
    const int n_old = init ? hd.n0 : p.cnt[h];
    const int nch = (n_old + p.chunk - 1) / p.chunk;
    int n_new = 0;
    double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float best = -1.f;          // most bound member of this pass (largest potential, lowest index)
    int best_q = -1;
    if (own) {
        n_new = own->count;
#pragma unroll
        for (int k = 0; k < kChunkSums; ++k) s[k] = own->s[k];
        best = own->best;
        best_q = own->best_q;
        if (lane == 0) p.chunk_off[hd.chunk_begin] = 0;
    } else {
        int carry = 0;
        for (int c0 = 0; c0 < nch; c0 += 32) {
            const int c = c0 + lane;
            const int v = (c < nch) ? __ldcg(&p.chunk_cnt[hd.chunk_begin + c]) : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (c < nch) {
                p.chunk_off[hd.chunk_begin + c] = carry + incl - v;
                const double *cs = p.chunk_sum + static_cast<int64_t>(hd.chunk_begin + c) * kChunkSums;
#pragma unroll
                for (int k = 0; k < kChunkSums; ++k) s[k] += __ldcg(&cs[k]);
                best_merge(best, best_q, __ldcg(&p.chunk_best[hd.chunk_begin + c]),
                           __ldcg(&p.chunk_best_q[hd.chunk_begin + c]));
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        n_new = carry;
        warp_best(best, best_q);
        warp_sum<kChunkSums>(s);
    }
    if (lane == 0) {
        // every load of the serial part first: their latencies overlap instead of adding up
        const int it_old = init ? 0 : p.iter[h];
        const unsigned long long pairs_old = init ? 0ull : p.pairs[h], evals_old = init ? 0ull : p.evals[h];
        const double rps2 = init ? 0.0 : p.hrps[4 * h + 2], rps3 = init ? 0.0 : p.hrps[4 * h + 3];
        const bool redo = !init && p.redo_enabled && p.halo_redo[h];
        const bool was_incr = !init && p.incr_enabled && p.incr[h] && !redo;
        const int rem_old = (!init && p.incr_enabled) ? p.rem_cnt[h] : 0;
        int ext_ok = p.cache_ext ? (init ? 1 : p.ext_ok[h]) : 0;
        double ext = 0.0;
        if (p.sym_enabled) {
            if (init) {
                // largest coordinate extent of the halo's members (float32 working set, from the pack)
                for (int a = 0; a < 3; ++a) {
                    const float lo = order_float(p.halo_rmin[3 * h + a]);
                    const float hi = order_float(p.halo_rmax[3 * h + a]);
                    ext = fmax(ext, static_cast<double>(hi) - static_cast<double>(lo));
                }
            } else {
                ext = p.sym_ext[h];
            }
        }
        double vbf[3] = {0.0, 0.0, 0.0};
        if (p.vb_fixed)
            for (int k = 0; k < 3; ++k) vbf[k] = p.hvb[3 * h + k];

        const double M = s[0];
        if (init) {
            p.hrps[4 * h + 0] = M;
            p.hrps[4 * h + 1] = p.hrps[4 * h + 2] = p.hrps[4 * h + 3] = 0.0;
            p.hbest[h] = -1;
        } else {
            p.hrps[4 * h + 1] = s[7];            // cold members bound after this pass
            p.hrps[4 * h + 2] = rps2 + s[8];     // removed, cold
            p.hrps[4 * h + 3] = rps3 + s[9];     // removed, hot
            p.hbest[h] = best_q;
        }
        p.hM[h] = M;
        if (p.sym_enabled) {
            if (init) p.sym_ext[h] = ext;
            // Quantum of the symmetric sums of the coming pass: every addend is rounded to a multiple of
            // q = 2^-37 * 2^ceil(log2(M / extent)), and a sum that stays below 2^52 q = 32768 * (1..2) * M / extent
            // is then EXACT in float64, so the order of the atomics cannot change it.  A sum that leaves the
            // window sends the halo to the one-sided kernel (potential.cu::sym_ticket).  0 = no quantisation.
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
            p.hvb_next[3 * h + k] = p.vb_fixed ? vbf[k] : (M > 0.0 ? s[1 + k] / M : 0.0);
        }
        p.cnt_next[h] = n_new;
        int act, inc_next = 0, n_rem = 0;
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
            const int it = it_old + 1;
            p.iter[h] = it;
            const unsigned long long nn = static_cast<unsigned long long>(n_old);
            p.pairs[h] = pairs_old + nn * static_cast<unsigned long long>((p.targets_only ? 0 : n_old) + hd.n_ext);
            const unsigned long long tiles = (nn + p.group_size - 1) / p.group_size;
            // externals: evaluated unless their first-pass sum was reused (cache) or kept (incremental)
            const bool ext_reused = p.cache_ext && pass > 0 && ext_ok && !redo;
            const unsigned long long ext_ev = ext_reused ? 0ull : nn * static_cast<unsigned long long>(hd.n_ext);
            unsigned long long ev;
            if (was_incr) {
                // survivors x the members the previous pass removed
                ev = nn * static_cast<unsigned long long>(rem_old);
            } else if (p.sym_enabled && tiles >= 2 && !redo) {
                // diagonal tiles one-sided, every other member pair once
                const unsigned long long last = nn - (tiles - 1) * p.group_size;
                const unsigned long long diag = (tiles - 1) * p.group_size * p.group_size + last * last;
                ev = ext_ev + (nn * nn + diag) / 2;
            } else {
                ev = (p.targets_only ? 0ull : nn * nn) + ext_ev;
            }
            p.evals[h] = evals_old + ev;
            // a first pass that fell back to the predicated kernel leaves no usable external sums
            if (p.cache_ext && pass == 0 && redo) {
                ext_ok = 0;
                p.ext_ok[h] = 0;
            }
            if (p.incr_enabled) {
                // The coming pass is incremental when this one left a valid potential behind (energy_phase,
                // phi_keep) and removed at most a third of the members: survivors x removed is then cheaper
                // than a full pass even with the symmetric self-term.  (Chains of incremental passes: the
                // incremental tickets compare with the last FULLY evaluated potential, potential_device.cuh.)
                n_rem = n_old - n_new;
                p.rem_cnt[h] = n_rem;
                inc_next = (!redo && n_rem > 0 && 2ll * n_rem <= n_new) ? 1 : 0;
                p.incr[h] = inc_next;
            }
            const int changed = n_new != n_old;
            p.converged[h] = (!changed || n_new == 0) ? 1 : 0;
            act = (changed && n_new > 0 && it < p.max_iter) ? 1 : 0;
            // the halo took part in this pass: its members now live in the other buffer (read by finalize_phase only)
            p.halo_buf[h] = par ^ 1;
        }
        p.active_next[h] = act;
        // Scheduling record of the coming pass, in `order` space so that schedule_block reads it coalesced:
        // members (0: no pass), sources a main ticket streams, incremental?, original member count.
        const bool ext_cached = p.cache_ext && !init && ext_ok;
        const int n_src = inc_next ? n_rem : n_new + (ext_cached ? 0 : hd.n_ext);
        p.sched[p.rank_of[h]] = make_int4(act ? n_new : 0, n_src, inc_next, hd.n0);
        if (act) {
            // plan-wide totals the ticket table needs (exact integer atomics: order-independent)
            const long long tiles = (n_new + p.group_size - 1) / p.group_size;
            atomicAdd(&p.st->next_groups, static_cast<int>(tiles));
            if (!inc_next) atomicAdd(&p.st->next_tile_pairs, static_cast<unsigned long long>(tiles * (tiles - 1) / 2));
        }
    }
}

__device__ __forceinline__ void decide_init_phase(const LoopParams &p)
{
    const int w = blockIdx.x * kLW + (threadIdx.x >> 5), nw = gridDim.x * kLW;
    for (int h = w; h < p.n_halo; h += nw) decide_halo(p, h, 1, 0, 0, nullptr);
}

// ---------------------------------------------------------------------------------------
// Kernel 2: energy step, bound flag, survivor counts and mass sums of one chunk.  One warp.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void energy_chunk(const LoopParams &p, int c, int h, int n, int par, int pass, ChunkSums &out)
{
    const int lane = threadIdx.x & 31;
    const int p0 = p.chunk_p0[c];
    const HaloDesc &hd = p.halo[h];
    double s[kChunkSums] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float best = -1.f;          // potentials are >= 0; NaN never wins
    int best_q = -1;
    int count = 0;
    const int S = p.nsplit[p.rank_of[h]];
    // a halo handed to the predicated kernel has its complete, final sum in the planes
    const bool redo = p.redo_enabled && p.halo_redo[h];
    const bool inc = p.incr_enabled && p.incr[h] && !redo;
    const bool ext_cached = p.cache_ext && hd.n_ext > 0 && p.ext_ok[h] && !redo && !inc;
    const bool corr = p.np_enabled && !redo && !inc;
    const double vb0 = p.hvb[3 * h + 0], vb1 = p.hvb[3 * h + 1], vb2 = p.hvb[3 * h + 2];
    // software pipeline: the slot -> user index lookups of the NEXT group of rounds are issued before this
    // group's loads, so that a group costs one memory latency, not two
    int64_t g_next[kHoist];
#pragma unroll
    for (int r = 0; r < kHoist; ++r) {
        const int q = p0 + r * 32 + lane;
        g_next[r] = q < n ? p.widx[par][hd.poff + q] : hd.uoff;
    }
#pragma unroll 1
    for (int r0 = 0; r0 < (p.chunk >> 5); r0 += kHoist) {
        if (p0 + r0 * 32 >= n) break;
        bool ok[kHoist];
        int64_t gi[kHoist];
        double phi[kHoist], vx[kHoist], vy[kHoist], vz[kHoist], mm[kHoist], ee[kHoist];
        double xx[kHoist], yy[kHoist], zz[kHoist];
#pragma unroll
        for (int r = 0; r < kHoist; ++r) {
            ok[r] = p0 + (r0 + r) * 32 + lane < n;
            gi[r] = g_next[r];
            const int qn = p0 + (r0 + kHoist + r) * 32 + lane;
            g_next[r] = (r0 + kHoist < (p.chunk >> 5) && qn < n) ? p.widx[par][hd.poff + qn] : hd.uoff;
        }
        // loads only, unconditional per lane (clamped index) and selected by warp-uniform flags, so that the
        // compiler can issue all of them before the first use: one memory latency per group of rounds
        double l_part[kHoist], l_sym[kHoist], l_keep[kHoist], l_ext[kHoist], l_c0[kHoist], l_c1[kHoist], l_c2[kHoist];
#pragma unroll
        for (int r = 0; r < kHoist; ++r) {
            const int64_t i = hd.poff + min(p0 + (r0 + r) * 32 + lane, n - 1);
            const int64_t g = gi[r];
            const int64_t slot = hd.poff + (g - hd.uoff);      // the member's original slot
            l_part[r] = p.phi_part[i];
            l_sym[r] = p.sym_enabled ? p.phi_sym[i] : 0.0;
            l_keep[r] = inc ? p.phi_keep[slot] : 0.0;
            l_ext[r] = ext_cached ? p.phi_ext[slot] : 0.0;
            l_c0[r] = corr ? p.ax[0].corr[slot] : 0.0;
            l_c1[r] = corr ? p.ax[1].corr[slot] : 0.0;
            l_c2[r] = corr ? p.ax[2].corr[slot] : 0.0;
            vx[r] = p.vx[g];
            vy[r] = p.vy[g];
            vz[r] = p.vz[g];
            mm[r] = p.m64[g];
            xx[r] = p.x64[g];          // (for the centre of mass of the survivors: most members are)
            yy[r] = p.y64[g];
            zz[r] = p.z64[g];
        }
#pragma unroll
        for (int r = 0; r < kHoist; ++r) {
            const int64_t i = hd.poff + min(p0 + (r0 + r) * 32 + lane, n - 1);
            const int64_t slot = hd.poff + (gi[r] - hd.uoff);
            // Phi: ascending sum of the j-split partials
            double ph = l_part[r];
            for (int k = 1; k < S; ++k) ph += p.phi_part[static_cast<int64_t>(k) * p.n_pad + i];
            ee[r] = 0.0;
            if (inc) {
                // incremental pass: the planes hold what the members removed by the previous pass contributed
                // (reference predicate applied); take it out of the potential kept from that pass
                ph = l_keep[r] - ph;
            } else if (!redo) {
                // the two-sided sums of this pass (the slot is cleared for the next one below)
                if (p.sym_enabled) ph += l_sym[r];
                if (ext_cached) {
                    // sum over the external sources, evaluated by the first pass only (potential.cu)
                    double e = l_ext[r];
                    if (pass == 0)
                        for (int k = 1; k < S; ++k) e += p.phi_ext[static_cast<int64_t>(k) * p.n_pad + slot];
                    ee[r] = e;
                    ph += e;
                }
                // predicate-free path: take out the pairs that share a coordinate (potential.cu)
                if (corr) ph -= (l_c0[r] + l_c1[r]) + l_c2[r];
            }
            phi[r] = ph;
        }
#pragma unroll
        for (int r = 0; r < kHoist; ++r) {
            int bound = 0;
            if (ok[r]) {
                const int64_t i = hd.poff + p0 + (r0 + r) * 32 + lane;
                const int64_t g = gi[r];
                const int64_t slot = hd.poff + (g - hd.uoff);
                if (p.sym_enabled) p.phi_sym[i] = 0.0;
                if (ext_cached && pass == 0) p.phi_ext[slot] = ee[r];      // the folded sum, read by the later passes
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
                bound = (E <= 0.0) ? 1 : 0;           // :359 / :476 (NaN is neither bound nor unbound)
                p.flag[i] = static_cast<uint8_t>(bound);
                p.out_mask[g] = static_cast<uint8_t>(bound);
                p.out_be[g] = be;
                p.out_E[g] = E;
                if (be > best) {                      // a lane's members ascend, so the lowest index wins ties
                    best = be;
                    best_q = static_cast<int>(g - hd.uoff);
                }
                const double m = mm[r];
                if (p.temp) {
                    // halo_gas.py:479-490: cold = T < 5e4, hot = T >= 5e4 (NaN is neither)
                    const double T = p.temp[g];
                    const bool cold = T < p.cold_T, hot = T >= p.cold_T;
                    if (bound && cold) s[7] += m;
                    if (E > 0.0 && cold) s[8] += m;
                    if (E > 0.0 && hot) s[9] += m;
                }
                if (bound) {
                    s[0] += m;
                    s[1] += m * vx[r];
                    s[2] += m * vy[r];
                    s[3] += m * vz[r];
                    s[4] += m * xx[r];
                    s[5] += m * yy[r];
                    s[6] += m * zz[r];
                }
            }
            count += __popc(__ballot_sync(0xffffffffu, bound));
        }
    }
    warp_sum<kChunkSums>(s);
    warp_best(best, best_q);
#pragma unroll
    for (int k = 0; k < kChunkSums; ++k) out.s[k] = s[k];
    out.count = count;
    out.best = best;
    out.best_q = best_q;
}

// Energy step of every chunk this warp owns; the warp that completes a halo decides it.
__device__ __forceinline__ void energy_phase(const LoopParams &p, int par, int pass)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kLW + (threadIdx.x >> 5), nw = gridDim.x * kLW;
    for (int c = w; c < p.n_chunks; c += nw) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        if (p.chunk_p0[c] >= n) continue;
        ChunkSums cs;
        energy_chunk(p, c, h, n, par, pass, cs);
        const int nch = (n + p.chunk - 1) / p.chunk;
        int last = 1;
        if (nch > 1) {
            if (lane == 0) {
                p.chunk_cnt[c] = cs.count;
#pragma unroll
                for (int k = 0; k < kChunkSums; ++k) p.chunk_sum[static_cast<int64_t>(c) * kChunkSums + k] = cs.s[k];
                p.chunk_best[c] = cs.best;
                p.chunk_best_q[c] = cs.best_q;
                __threadfence();                                   // this chunk's results before the count
                last = atomicAdd(&p.halo_done[h], 1) + 1 == nch;
                if (last) {
                    p.halo_done[h] = 0;
                    __threadfence();                               // the other chunks' results after it
                }
            }
            last = __shfl_sync(0xffffffffu, last, 0);
        }
        if (last) {
            decide_halo(p, h, 0, par, pass, nch == 1 ? &cs : nullptr);
            if (lane == 0) {
                __threadfence();
                st_release(&p.halo_stamp[h], pass + 1);            // compact_phase may go ahead with this halo
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel 3: stable stream compaction (ballot + popc inside a warp, running offset across its rounds,
// chunk offsets from decide_halo).  Order-preserving, so the member indices stay ascending like
// part_list[bound] (halo_properties.py:359-361).  One warp per chunk.
// wait: inside the persistent kernel, spin until the halo's decision of this pass is published.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void compact_phase(const LoopParams &p, int par, int pass, bool wait)
{
    const int nxt = par ^ 1;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kLW + (threadIdx.x >> 5), nw = gridDim.x * kLW;
    for (int c = w; c < p.n_chunks; c += nw) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        const int p0 = p.chunk_p0[c];
        if (p0 >= n) continue;
        if (wait) {
            if (lane == 0)
                while (ld_acquire(&p.halo_stamp[h]) != pass + 1) __nanosleep(64);
            __syncwarp();
        }
        const HaloDesc &hd = p.halo[h];
        int before = __ldcg(&p.chunk_off[c]);                    // survivors before this round, whole halo
        const int inc_next = p.incr_enabled ? __ldcg(&p.incr[h]) : 0;
#pragma unroll 1
        for (int r0 = 0; r0 < (p.chunk >> 5); r0 += kHoistC) {
            if (p0 + r0 * 32 >= n) break;
            int f[kHoistC], inv[kHoistC][3];
            float vals[kHoistC][4];
            int32_t wid[kHoistC];
#pragma unroll
            for (int r = 0; r < kHoistC; ++r) {
                const int q = p0 + (r0 + r) * 32 + lane;
                const int64_t i = hd.poff + min(q, n - 1);
                f[r] = (q < n) ? p.flag[i] : 0;
                vals[r][0] = p.wx[par][i];
                vals[r][1] = p.wy[par][i];
                vals[r][2] = p.wz[par][i];
                vals[r][3] = p.wm[par][i];
                wid[r] = p.widx[par][i];
            }
#pragma unroll
            for (int r = 0; r < kHoistC; ++r) {
                const int q = p0 + (r0 + r) * 32 + lane;
                const int64_t slot = hd.poff + (wid[r] - hd.uoff);
#pragma unroll
                for (int a = 0; a < 3; ++a) inv[r][a] = (p.np_enabled && q < n && !f[r]) ? p.ax[a].inv[slot] : 0;
            }
#pragma unroll
            for (int r = 0; r < kHoistC; ++r) {
                const int q = p0 + (r0 + r) * 32 + lane;
                const unsigned ballot = __ballot_sync(0xffffffffu, f[r]);
                const int dst = before + __popc(ballot & ((1u << lane) - 1u));
                before += __popc(ballot);
                if (q >= n) continue;
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
        }
    }
}

// ---------------------------------------------------------------------------------------
// Per halo: the state decided by the pass becomes current.  One thread per halo, grid-stride.
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
    // sources a main ticket streams: with the symmetric self-term only ONE tile of the members (its own) plus the
    // externals -- the other member tiles belong to the symmetric tickets, and cutting them into j-split pieces only
    // made empty tickets (60 % more tickets in the first pass of an eighth of the cfg3 catalogue, each fetched,
    // decoded and stored as a plane of zeros that the energy step then added up)
    int n_src = rec.y;
    if (p.sym_enabled && !inc && n > p.group_size) n_src = rec.y - n + p.group_size;
    S = min(want, min(p.max_split, max(1, n_src / p.min_split_sources)));
    // An incremental pass is short (survivors x removed), so a big halo's tickets -- 128 targets x all the members it
    // removed -- can each last as long as the rest of the pass: where the pass has few tickets (want > 1), cut its
    // removed list into pieces of ~kIncPieceSources (at least a tile).  (With plenty of tickets the extra pieces only
    // cost: every piece loads its targets and kept potentials again.)
    if (inc && want > 1) S = min(p.max_split, max(S, min(rec.y / kIncPieceSources, rec.y / 128)));
    S = max(S, 1);
    int items = my_groups(n, p.group_size, p.rank, p.n_ranks) * S;
    // correction tickets: blocks of the (static) sorted member lists, the three axes in one ticket
    if (p.np_enabled && !inc) items += my_groups(rec.w, p.group_size, p.rank, p.n_ranks);
    // symmetric tickets: row tiles x chunks of column tiles
    if (p.sym_enabled && !inc) {
        const int tiles = (n + p.group_size - 1) / p.group_size;
        const int ct = sym_chunk_of(sym_chunk, tiles);
        if (tiles >= 2) items += sym_units(tiles, p.sym_rows) * ((tiles - 1 + ct - 1) / ct);
    }
    return items;
}

// ---------------------------------------------------------------------------------------
// Ticket table of the next potential pass, from the scheduling records.  ONE block of kLT threads.
// Tickets are laid out in `order` (largest halo first).  Does not touch the per-halo state the
// other phases read, so it may run next to commit_phase.
// smem: kSchedSmemBytes of shared memory (the persistent kernel lends the idle TMA ring).
// ---------------------------------------------------------------------------------------
constexpr int kSchedPer = 8;                       // records per thread and tile
constexpr int kSchedTile = kSchedPer * kLT;        // 1024 records = 16 KB
constexpr int kSchedSmemBytes = kSchedTile * 16 + (kLW + 1) * 4;

__device__ __forceinline__ void schedule_block(const LoopParams &p, unsigned char *smem, int init)
{
    LoopState *st = p.st;
    int4 *stage = reinterpret_cast<int4 *>(smem);
    int *wsum = reinterpret_cast<int *>(smem + kSchedTile * 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // plan-wide totals of the coming pass, accumulated by decide_halo.  All ranks' groups: the j-split must
    // depend on the problem only, so that a split run sums its partial potentials in the same grouping as a
    // single-GPU run.
    const int total_groups = st->next_groups;
    const long long tp = static_cast<long long>(st->next_tile_pairs);
    const int any = total_groups > 0;
    // column tiles per symmetric ticket: about kNominalTickets tickets over the whole plan, between
    // 2 and 32 -- a lone mid-size halo gets short tickets that fill the machine, a catalogue or a giant
    // halo long ones that amortise the per-ticket work.  Depends on the plan only, not on the GPU.
    const long long cc = tp / kNominalTickets;
    const int sym_chunk = cc < 2 ? 2 : (cc > 32 ? 32 : static_cast<int>(cc));
    int want = 1;
    if (p.mode == HALMA_MODE_FAST && total_groups > 0 && total_groups < p.target_items)
        want = (p.target_items + total_groups - 1) / total_groups;

    int carry = 0, max_split = 1;
    for (int t0 = 0; t0 < p.n_halo; t0 += kSchedTile) {
        __syncthreads();
        // coalesced, independent loads of the tile's records
#pragma unroll
        for (int j = 0; j < kSchedPer; ++j) {
            const int k = t0 + j * kLT + threadIdx.x;
            stage[j * kLT + threadIdx.x] = k < p.n_halo ? p.sched[k] : make_int4(0, 0, 0, 0);
        }
        __syncthreads();
        // every thread takes a contiguous run of the tile: items, local prefix
        int items[kSchedPer], S[kSchedPer], mine = 0;
#pragma unroll
        for (int j = 0; j < kSchedPer; ++j) {
            items[j] = sched_items(p, stage[threadIdx.x * kSchedPer + j], want, sym_chunk, S[j]);
            mine += items[j];
            max_split = max(max_split, S[j]);
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int run = carry + incl - mine;
        int tile_total = 0;
        for (int w2 = 0; w2 < kLW; ++w2) {
            if (w2 < warp) run += wsum[w2];
            tile_total += wsum[w2];
        }
#pragma unroll
        for (int j = 0; j < kSchedPer; ++j) {
            const int k = t0 + threadIdx.x * kSchedPer + j;
            if (k < p.n_halo) {
                p.item_base[k] = run;
                p.nsplit[k] = S[j];
            }
            run += items[j];
        }
        carry += tile_total;
    }
    if (threadIdx.x == 0) {
        p.item_base[p.n_halo] = carry;
        st->n_items = carry;
        st->any_active = any;
        st->counter = 0u;
        st->counter_redo = 0u;
        st->redo_any = 0;
        st->sym_chunk = sym_chunk;
        st->next_groups = 0;
        st->next_tile_pairs = 0ull;
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
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kLW + (threadIdx.x >> 5), nw = gridDim.x * kLW;
    for (int c = w; c < p.n_chunks; c += nw) {
        const int h = p.chunk_halo[c];
        const HaloDesc &hd = p.halo[h];
        const int n = p.cnt[h];
        const int b = p.halo_buf[h];
        const int p0 = p.chunk_p0[c];
        for (int r = 0; r < (p.chunk >> 5); ++r) {
            const int q = p0 + r * 32 + lane;
            if (q < hd.n0)
                p.out_idx[hd.uoff + q] = (q < n) ? static_cast<int32_t>(p.widx[b][hd.poff + q] - hd.uoff) : -1;
        }
    }
}

}  // namespace halma
