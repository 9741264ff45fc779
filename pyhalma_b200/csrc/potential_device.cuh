// Device side of the all-pairs potential (kernel 1): ticket decoding, TMA tile ring, the FAST / EXACT
// bodies and the per-pass ticket loops.  Included by potential.cu (stand-alone kernels) and fused.cu (the
// whole unbinding loop as one persistent kernel).  See potential.cu for the description of the paths.
#pragma once
#include "halma_common.cuh"
#include "potential.h"

namespace halma {

constexpr int kTileJ = 128;                 // sources per shared-memory tile
constexpr int kStages = 3;                  // ring depth per warp
constexpr int kStageFloats = 4 * kTileJ;    // x | y | z | m
constexpr int kFlushQuads = 8;              // flush float32 partials every 8 quads = 32 sources ...
constexpr int kFlushQuadsNp = 16;           // ... or 64 in the predicate-free body, whose packed accumulator
                                            // keeps even and odd sources apart (32 terms per float32 sum)
constexpr double kIncrHeavy = 0.8;          // incremental pass: largest removed share of a member's kept potential
                                            // (error amplification 1 / (1 - 0.8) = 5 on ~1e-7)
constexpr int kWarpsPerBlock = kPotentialBlock / 32;
// ring stages + their mbarriers + the column partial sums of the symmetric tickets
constexpr int kSmemBytes = kWarpsPerBlock * (kStages * (kStageFloats * 4 + 8) + kTileJ * 4);

// ---------------------------------------------------------------------------------------
// Cursor over the source tiles of one ticket.  Main tickets: piece `s` of `S` of every
// segment of the halo, in segment order.  Correction tickets: one explicit range of a
// sorted copy.  Uniform across the warp.
// ---------------------------------------------------------------------------------------
struct TileCursor {
    const HaloDesc *hd;
    int nseg, k, S, s, n_members, parity;
    int64_t base;
    int set, pos, end, flags;
    int own_tile;       // >= 0: symmetric mode, the members segment is only this tile (taken by split 0)
    int filter;         // 0: every segment; 1: the members segment only; 2: the external segments only

    __device__ __forceinline__ void seek()
    {
        while (k < nseg) {
            const SegDesc sd = hd->seg[k];
            const bool is_members = (sd.flags & kSegMembers) != 0;
            if ((filter == 1 && !is_members) || (filter == 2 && is_members)) {
                ++k;
                continue;
            }
            const int c = (sd.flags & kSegMembers) ? n_members : sd.count;
            const int per = (((c + S - 1) / S) + kTileJ - 1) / kTileJ * kTileJ;
            int a = s * per;
            int b = min(a + per, c);
            if ((sd.flags & kSegMembers) && own_tile >= 0) {
                a = s == 0 ? own_tile * kTileJ : c;
                b = min(a + kTileJ, c);
            }
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
    __device__ __forceinline__ void init(const HaloDesc *h, int S_, int s_, int n_members_, int parity_,
                                         int own_tile_ = -1, int filter_ = 0)
    {
        own_tile = own_tile_;
        filter = filter_;
        hd = h;
        nseg = h->nseg;
        k = 0;
        S = S_;
        s = s_;
        n_members = n_members_;
        parity = parity_;
        seek();
    }
    __device__ __forceinline__ void init_range(int set_, int64_t begin, int count)
    {
        own_tile = -1;
        filter = 0;
        hd = nullptr;
        nseg = 1;
        k = count > 0 ? 0 : 1;
        S = 1;
        s = 0;
        n_members = 0;
        parity = 0;
        base = begin;
        set = set_;
        pos = 0;
        end = count;
        flags = 0;
    }
    __device__ __forceinline__ bool valid() const { return k < nseg; }
    __device__ __forceinline__ int len() const { return min(kTileJ, end - pos); }
    __device__ __forceinline__ void next()
    {
        pos += kTileJ;
        if (pos >= end) {
            ++k;
            if (hd) seek();
        }
    }
};

// Fill one stage with the cursor's tile.  Whole quads come by TMA; the last 1..3 sources of
// a segment are fetched with plain loads (never reading past the array) and the quad is
// padded with (inf, inf, inf, m = 0), which contributes exactly zero in every body.
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
// Tickets
// ---------------------------------------------------------------------------------------
struct Ticket {
    int h, group, s, S, n_tgt;
    int corr_axis;      // -1: main ticket; 0: correction ticket (all three axes), `group` = block;
                        // 3: symmetric ticket, `group` = row tile, `s` = chunk of column tiles
};

// Ticket layout of a halo (must match loop_kernels.cu::k_schedule):
//   [0, mine * S)                    main tickets, s-major
//   [mine * S, mine * S + myblk)     correction tickets (NP path), one per block and all three axes; myblk = this
//                                    rank's share of the ceil(n0 / group) blocks of the sorted member lists
template <int kGroup>
__device__ __forceinline__ bool decode_ticket(const PotParams &p, int item, Ticket &t)
{
    // largest k with item_base[k] <= item (item_base[0] = 0): 32 probes per round, three dependent loads for 1e4
    // haloes instead of fourteen.  The whole warp decodes the same item.
    int lo = 0;
    {
        const int lane = threadIdx.x & 31;
        int n = p.n_halo;
        while (n > 1) {
            const int step = (n + 31) >> 5;
            const int off = lane * step;
            const int v = off < n ? p.item_base[lo + off] : 0x7fffffff;
            const int c = __popc(__ballot_sync(0xffffffffu, v <= item));      // a prefix of the lanes, lane 0 included
            lo += (c - 1) * step;
            n = min(step, n - (c - 1) * step);
        }
    }
    int local = item - p.item_base[lo];
    t.h = p.order[lo];
    t.n_tgt = p.cnt ? p.cnt[t.h] : p.halo[t.h].n0;
    t.S = p.nsplit[lo];
    t.corr_axis = -1;
    const int groups = (t.n_tgt + kGroup - 1) / kGroup;
    const int mine = (groups - p.rank + p.n_ranks - 1) / p.n_ranks;
    const int n_main = mine * t.S;
    if (local < n_main) {
        t.s = local / mine;
        t.group = (local % mine) * p.n_ranks + p.rank;
        return true;
    }
    if (!p.np_enabled) return false;
    // a halo whose coming pass is incremental has main tickets only: they apply the reference's predicate to
    // the few removed members themselves (loop_kernels.cu::k_schedule counts the same way)
    if (p.incr_enabled && p.incr[t.h]) return false;
    local -= n_main;
    // correction blocks are dealt round-robin to the ranks like the target groups
    const int nblk = (p.halo[t.h].n0 + kGroup - 1) / kGroup;
    const int myblk = (nblk - p.rank + p.n_ranks - 1) / p.n_ranks;
    if (myblk > 0 && local < myblk) {
        t.corr_axis = 0;
        t.group = local * p.n_ranks + p.rank;
        t.s = 0;
        return true;
    }
    if (!p.sym_enabled || groups < 2) return false;
    // symmetric tickets: (row unit u, chunk c of its column tiles); ids of empty chunks are skipped.  A row unit is
    // one 128-member tile (sym_rows = 4 members per lane) or a pair of consecutive tiles (sym_rows = 8).
    local -= max(myblk, 0);
    const int ct = sym_chunk_of(p.st->sym_chunk, groups);
    const int chunks = (groups - 1 + ct - 1) / ct;
    const int units = sym_units(groups, p.sym_rows);
    if (local >= units * chunks) return false;
    t.corr_axis = 3;
    t.group = local / chunks;
    t.s = local % chunks;
    return t.group % p.n_ranks == p.rank;      // split mode: row units are dealt round-robin like the target groups
}

// ---------------------------------------------------------------------------------------
// FAST bodies
// ---------------------------------------------------------------------------------------
enum Body { kPred = 0, kNp = 1, kGuard = 2, kCorrX = 3, kCorrY = 4, kCorrZ = 5 };

template <int BODY>
__device__ __forceinline__ void pair_body(uint64_t X, uint64_t Y, uint64_t Z, float m0, float m1, float xi, float yi,
                                          float zi, float &acc, uint64_t &acc2)
{
    const uint64_t dx = sub2(X, pack2(xi, xi));
    const uint64_t dy = sub2(Y, pack2(yi, yi));
    const uint64_t dz = sub2(Z, pack2(zi, zi));
    const uint64_t r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    float r20, r21;
    unpack2(r2, r20, r21);
    const float i0 = rsqrt_ftz(r20);
    const float i1 = rsqrt_ftz(r21);
    if (BODY == kNp) {
        acc2 = fma2(pack2(m0, m1), pack2(i0, i1), acc2);
        return;
    }
    float dx0, dx1, dy0, dy1, dz0, dz1;
    unpack2(dx, dx0, dx1);
    unpack2(dy, dy0, dy1);
    unpack2(dz, dz0, dz1);
    bool p0, p1;
    if (BODY == kPred) {
        // particle_subroutines.f90:499-501: all three coordinates differ
        p0 = min3abs(dx0, dy0, dz0) > 0.f;
        p1 = min3abs(dx1, dy1, dz1) > 0.f;
    } else if (BODY == kGuard) {
        p0 = r20 > 0.f;
        p1 = r21 > 0.f;
    } else if (BODY == kCorrX) {
        p0 = (dx0 == 0.f) && (r20 > 0.f);
        p1 = (dx1 == 0.f) && (r21 > 0.f);
    } else if (BODY == kCorrY) {
        p0 = (dy0 == 0.f) && (dx0 != 0.f);
        p1 = (dy1 == 0.f) && (dx1 != 0.f);
    } else {
        p0 = (dz0 == 0.f) && (dx0 != 0.f) && (dy0 != 0.f);
        p1 = (dz1 == 0.f) && (dx1 != 0.f) && (dy1 != 0.f);
    }
    if (p0) acc = fmaf(m0, i0, acc);
    if (p1) acc = fmaf(m1, i1, acc);
}

template <int T, int BODY>
__device__ __forceinline__ void tile_fast(const float *__restrict__ stage, int len, const float (&xi)[T],
                                          const float (&yi)[T], const float (&zi)[T], double (&acc64)[T])
{
    const float4 *X = reinterpret_cast<const float4 *>(stage);
    const float4 *Y = reinterpret_cast<const float4 *>(stage + kTileJ);
    const float4 *Z = reinterpret_cast<const float4 *>(stage + 2 * kTileJ);
    const float4 *M = reinterpret_cast<const float4 *>(stage + 3 * kTileJ);
    const int nq = (len + 3) >> 2;
    constexpr int kFlush = BODY == kNp ? kFlushQuadsNp : kFlushQuads;
    for (int qb = 0; qb < nq; qb += kFlush) {
        const int qe = min(qb + kFlush, nq);
        float acc[T];
        uint64_t acc2[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            acc[t] = 0.f;
            acc2[t] = 0ull;
        }
#pragma unroll 2
        for (int q = qb; q < qe; ++q) {
            const float4 x4 = X[q], y4 = Y[q], z4 = Z[q], m4 = M[q];
            const uint64_t x01 = pack2(x4.x, x4.y), x23 = pack2(x4.z, x4.w);
            const uint64_t y01 = pack2(y4.x, y4.y), y23 = pack2(y4.z, y4.w);
            const uint64_t z01 = pack2(z4.x, z4.y), z23 = pack2(z4.z, z4.w);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                pair_body<BODY>(x01, y01, z01, m4.x, m4.y, xi[t], yi[t], zi[t], acc[t], acc2[t]);
                pair_body<BODY>(x23, y23, z23, m4.z, m4.w, xi[t], yi[t], zi[t], acc[t], acc2[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (BODY == kNp) {
                float lo, hi;
                unpack2(acc2[t], lo, hi);
                acc[t] = lo + hi;
            }
            acc64[t] += static_cast<double>(acc[t]);
        }
    }
}

// Ring state of one warp (continues across tickets).
struct Ring {
    float *buf;
    uint64_t *bars;
    uint32_t par;
    int fill, use;
};

// Streams every tile of `prod` through the ring and applies the body chosen per tile.
// CTX 0: predicated main ticket; 1: predicate-free main ticket (own tile guarded);
// 2..4: correction ticket of axis CTX-2.
template <int T, int CTX>
__device__ __forceinline__ void run_tiles(const PotParams &p, TileCursor prod, Ring &rg, int lane, int own_lo,
                                          int own_hi, const float (&xi)[T], const float (&yi)[T],
                                          const float (&zi)[T], double (&acc64)[T])
{
    TileCursor cons = prod;
#pragma unroll 1
    for (int i = 0; i < kStages - 1 && prod.valid(); ++i) {
        issue_tile(p, prod, rg.buf + rg.fill * kStageFloats, &rg.bars[rg.fill], lane);
        rg.fill = (rg.fill + 1 == kStages) ? 0 : rg.fill + 1;
        prod.next();
    }
#pragma unroll 1
    while (cons.valid()) {
        if (prod.valid()) {
            issue_tile(p, prod, rg.buf + rg.fill * kStageFloats, &rg.bars[rg.fill], lane);
            rg.fill = (rg.fill + 1 == kStages) ? 0 : rg.fill + 1;
            prod.next();
        }
        mbar_wait(&rg.bars[rg.use], (rg.par >> rg.use) & 1u);
        rg.par ^= 1u << rg.use;
        const float *stage = rg.buf + rg.use * kStageFloats;
        const int len = cons.len();
        if (CTX == 0) {
            tile_fast<T, kPred>(stage, len, xi, yi, zi, acc64);
        } else if (CTX == 1) {
            // the tile that holds the group's own members can contain zero separations
            const bool own = (cons.flags & kSegMembers) && cons.pos < own_hi && cons.pos + kTileJ > own_lo;
            if (own)
                tile_fast<T, kGuard>(stage, len, xi, yi, zi, acc64);
            else
                tile_fast<T, kNp>(stage, len, xi, yi, zi, acc64);
        } else if (CTX == 2) {
            tile_fast<T, kCorrX>(stage, len, xi, yi, zi, acc64);
        } else if (CTX == 3) {
            tile_fast<T, kCorrY>(stage, len, xi, yi, zi, acc64);
        } else {
            tile_fast<T, kCorrZ>(stage, len, xi, yi, zi, acc64);
        }
        __syncwarp();      // every lane is done with this stage before it is refilled
        rg.use = (rg.use + 1 == kStages) ? 0 : rg.use + 1;
        cons.next();
    }
}

__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t *a, int64_t lo, int64_t hi, uint32_t v)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int64_t upper_bound_u32(const uint32_t *a, int64_t lo, int64_t hi, uint32_t v)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// Correction ticket: 32*T consecutive members of the axis-sorted order against the sources that share their AXIS
// coordinate.  Equal coordinates are equal keys (canonical bits), i.e. one contiguous run of the sorted copy around
// the member.  Particles have short runs (positions cast to float32 in a cosmological box collide now and then),
// lattice cells runs of a whole plane.  Short runs: every lane walks the runs of its own T members (a handful of
// evaluations per member).  Long runs: the range of sorted sources between the first and the last member's key is
// streamed through the ring against all 32*T members, 128 x (hi - lo) evaluations whatever the run lengths.
constexpr int kCorrRunMax = 16;      // longest run the per-member walk takes
template <int T, int AXIS>
__device__ __forceinline__ void correction_ticket(const PotParams &p, const Ticket &tk, Ring &rg, int lane)
{
    constexpr int kGroup = 32 * T;
    const HaloDesc *hd = &p.halo[tk.h];
    const SortedAxis &A = p.ax[AXIS];
    const int n0 = hd->n0;
    const int li0 = tk.group * kGroup;
    const int64_t sb = hd->sbegin, se = hd->sbegin + n0 + hd->n_ext;
    float xi[T], yi[T], zi[T];
    int slot[T];
    int jl[T], jr[T];          // the run of sources sharing member t's key (sorted positions fit 32 bits: tgt)
    double acc64[T];
    bool longrun = false;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int li = li0 + t * 32 + lane;
        const bool ok = li < n0;
        const int k = ok ? A.tgt[hd->uoff + li] : 0;
        xi[t] = ok ? A.x[k] : 0.f;
        yi[t] = ok ? A.y[k] : 0.f;
        zi[t] = ok ? A.z[k] : 0.f;
        slot[t] = ok ? A.slot[k] : -1;
        acc64[t] = 0.0;
        jl[t] = jr[t] = k;
        if (ok) {
            const uint32_t key = A.key[k];
            while (jl[t] > sb && A.key[jl[t] - 1] == key && k - jl[t] < kCorrRunMax) --jl[t];
            while (jr[t] + 1 < se && A.key[jr[t] + 1] == key && jr[t] - k < kCorrRunMax) ++jr[t];
            longrun |= k - jl[t] >= kCorrRunMax || jr[t] - k >= kCorrRunMax;
        } else {
            jr[t] = -1;      // an empty run
        }
    }
    if (!__any_sync(0xffffffffu, longrun)) {
        const float kInf = __int_as_float(0x7f800000);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float acc = 0.f;
            uint64_t unused = 0ull;
            for (int j = jl[t]; j <= jr[t]; ++j) {
                // the member itself is in its run: zero separation, which every correction body skips; the second
                // slot of the packed body is padding (infinitely far, massless)
                pair_body<kCorrX + AXIS>(pack2(A.x[j], kInf), pack2(A.y[j], kInf), pack2(A.z[j], kInf), A.m[j], 0.f,
                                         xi[t], yi[t], zi[t], acc, unused);
            }
            acc64[t] = static_cast<double>(acc);
        }
    } else {
        const int64_t k_first = A.tgt[hd->uoff + li0];
        const int64_t k_last = A.tgt[hd->uoff + min(li0 + kGroup, n0) - 1];
        int64_t lo = lower_bound_u32(A.key, sb, se, A.key[k_first]);
        const int64_t hi = upper_bound_u32(A.key, sb, se, A.key[k_last]);
        lo &= ~int64_t(3);       // 16-byte alignment for the bulk copies; sb is a multiple of 4
        TileCursor cur;
        cur.init_range(3 + AXIS, lo, static_cast<int>(hi - lo));
        run_tiles<T, 2 + AXIS>(p, cur, rg, lane, 0, 0, xi, yi, zi, acc64);
    }
#pragma unroll
    for (int t = 0; t < T; ++t)
        if (slot[t] >= 0) A.corr[slot[t]] = acc64[t];
}

// ---------------------------------------------------------------------------------------
// Symmetric tickets (opt-in, halma_unbind_config.symmetric).  A pair of members in different
// tiles is evaluated ONCE: m_j / r goes to the row particle's sum, m_i / r to the column
// particle's.  Half the MUFU.RSQ work for the member x member term.
// The warp holds the 128 members of row tile I in registers (4 per lane) and streams column
// tiles J > I (2 to 32 per ticket, LoopState::sym_chunk) through the TMA ring.  Inside a tile lanes sweep the 64 source PAIRS in rotation
// (lane l visits pair (k + l) mod 64 at step k), so the column partial sums live in shared memory
// as plain read-modify-writes without conflicts.  Row and column sums are added to phi_sym with
// float64 atomics whose order is not fixed; the addends are therefore rounded to a per-halo quantum
// inside whose window float64 addition is exact (sym_add), which keeps runs bit-reproducible.
// scripts/probes/symmetric_probe.cu: 3.4-3.5 T pair evaluations/s = 6.8-7.1 T interactions/s.
// ---------------------------------------------------------------------------------------
template <int TR>
__device__ __forceinline__ void sym_tile(const float *__restrict__ stage, float *__restrict__ col, const float (&xi)[TR],
                                         const float (&yi)[TR], const float (&zi)[TR], const float (&mi)[TR],
                                         double (&acc64)[TR], int lane)
{
    const uint64_t *X = reinterpret_cast<const uint64_t *>(stage);
    const uint64_t *Y = reinterpret_cast<const uint64_t *>(stage + kTileJ);
    const uint64_t *Z = reinterpret_cast<const uint64_t *>(stage + 2 * kTileJ);
    const uint64_t *M = reinterpret_cast<const uint64_t *>(stage + 3 * kTileJ);
    uint64_t *C2 = reinterpret_cast<uint64_t *>(col);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {          // row partials: 32 terms per float32 sum, then float64
        uint64_t acc2[TR];
#pragma unroll
        for (int t = 0; t < TR; ++t) acc2[t] = 0ull;
        auto step = [&](int j, uint64_t x01, uint64_t y01, uint64_t z01, uint64_t m01) {
            uint64_t c2 = C2[j];          // the column pair's running sums: the first accumulate below adds to them
#pragma unroll
            for (int t = 0; t < TR; ++t) {
                const uint64_t dx = sub2(x01, pack2(xi[t], xi[t]));
                const uint64_t dy = sub2(y01, pack2(yi[t], yi[t]));
                const uint64_t dz = sub2(z01, pack2(zi[t], zi[t]));
                const uint64_t r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                float r20, r21;
                unpack2(r2, r20, r21);
                const uint64_t inv = pack2(rsqrt_ftz(r20), rsqrt_ftz(r21));
                acc2[t] = fma2(m01, inv, acc2[t]);
                c2 = fma2(pack2(mi[t], mi[t]), inv, c2);
            }
            C2[j] = c2;
            __syncwarp();          // the next step hands pair j to the neighbouring lane
        };
        if constexpr (TR == 8) {
            // the sources of step k + 1 are fetched before the warp barrier of step k (they are read-only; only the
            // column sums are handed from lane to lane), so a step starts with its operands in registers: one
            // 1e6-star pass 159.5 -> 156.2 ms.  (The 4-row kernel sits at its 80-register cap; there it costs 0.5 %.)
            int j = (half * 32 + lane) & (kTileJ / 2 - 1);
            uint64_t x01 = X[j], y01 = Y[j], z01 = Z[j], m01 = M[j];
#pragma unroll 2
            for (int k = 0; k < 32; ++k) {
                const int jn = (j + 1) & (kTileJ / 2 - 1);
                const uint64_t xn = X[jn], yn = Y[jn], zn = Z[jn], mn = M[jn];
                step(j, x01, y01, z01, m01);
                j = jn;
                x01 = xn;
                y01 = yn;
                z01 = zn;
                m01 = mn;
            }
        } else {
#pragma unroll 2
            for (int k = half * 32; k < half * 32 + 32; ++k) {
                const int j = (k + lane) & (kTileJ / 2 - 1);
                step(j, X[j], Y[j], Z[j], M[j]);
            }
        }
#pragma unroll
        for (int t = 0; t < TR; ++t) {
            float lo, hi;
            unpack2(acc2[t], lo, hi);
            acc64[t] += static_cast<double>(lo + hi);
        }
    }
}

// Adds v, rounded to a multiple of q, to *dst; returns true when the sum left the window in which
// float64 addition of such multiples is exact (the caller then hands the halo to the one-sided kernel).
__device__ __forceinline__ bool sym_add(double *dst, double v, double q, double inv_q, double window)
{
    if (q > 0.0) {
        v = __dmul_rn(rint(__dmul_rn(v, inv_q)), q);
        const double old = atomicAdd(dst, v);
        return !(fabs(old + v) < window);
    }
    atomicAdd(dst, v);
    return false;
}

// TR = 4: the row unit is tile tk.group, columns are the tiles after it.  TR = 8: the row unit is a PAIR of tiles
// (2u, 2u + 1) held as 8 members per lane -- half the shared-memory and bookkeeping instructions per evaluation --
// or the single last row tile; chunk 0 of a pair first takes the pairs between its own two tiles (tile 2u + 1 as a
// column, with the rows of that tile masked out: far away and massless, which contributes exactly zero).
template <int TR>
__device__ __forceinline__ void sym_ticket(const PotParams &p, const Ticket &tk, Ring &rg, float *col, int lane,
                                           int parity)
{
    const HaloDesc *hd = &p.halo[tk.h];
    const int n = tk.n_tgt;
    const int G = (n + kTileJ - 1) / kTileJ;
    const int ct = sym_chunk_of(p.st->sym_chunk, G);
    int I0, jfirst;
    bool pair = false;
    if (TR == 8) {
        const int P = (G - 1) >> 1;
        pair = tk.group < P;
        I0 = pair ? 2 * tk.group : G - 2;
        jfirst = pair ? I0 + 2 : G - 1;
    } else {
        I0 = tk.group;
        jfirst = I0 + 1;
    }
    const bool intra = TR == 8 && pair && tk.s == 0;      // this ticket also does tile I0 x tile I0 + 1
    int j0 = jfirst + tk.s * ct;
    const int j1 = min(j0 + ct, G);
    if (intra) j0 = I0 + 1;                               // ... which sits right in front of its first column tile
    if (j0 >= G) return;                                  // an empty chunk of this row unit
    const F32Set sp = p.src[parity];
    const int64_t base = hd->poff;
    constexpr float kFar = 1e30f;      // (x - 1e30)^2 overflows to +inf, rsqrt(+inf) = +0: a masked row adds nothing
    float xi[TR], yi[TR], zi[TR], mi[TR];
    double acc64[TR];
#pragma unroll
    for (int t = 0; t < TR; ++t) {                        // row tiles are tiles < G - 1: always full
        const int64_t i = base + static_cast<int64_t>(I0 + (t >> 2)) * kTileJ + (t & 3) * 32 + lane;
        const bool live = t < 4 || (pair && !intra);
        xi[t] = live ? sp.x[i] : kFar;
        yi[t] = live ? sp.y[i] : kFar;
        zi[t] = live ? sp.z[i] : kFar;
        mi[t] = live ? sp.m[i] : 0.f;
        acc64[t] = 0.0;
    }
    bool bad = false, first = true;
    // addends are rounded to multiples of q; while every running sum stays inside +-2^52 q the float64
    // additions are exact and their order does not matter (loop_device.cuh::decide_halo)
    // (split mode: every rank keeps to its share of the window, so the all-reduced total fits too)
    const double q = p.sym_q[tk.h], inv_q = q > 0.0 ? 1.0 / q : 0.0, window = ldexp(q, 52) / p.n_ranks;
    TileCursor prod;
    prod.init_range(parity, base + static_cast<int64_t>(j0) * kTileJ, min(n, j1 * kTileJ) - j0 * kTileJ);
    TileCursor cons = prod;
#pragma unroll 1
    for (int i = 0; i < kStages - 1 && prod.valid(); ++i) {
        issue_tile(p, prod, rg.buf + rg.fill * kStageFloats, &rg.bars[rg.fill], lane);
        rg.fill = (rg.fill + 1 == kStages) ? 0 : rg.fill + 1;
        prod.next();
    }
#pragma unroll 1
    while (cons.valid()) {
        if (prod.valid()) {
            issue_tile(p, prod, rg.buf + rg.fill * kStageFloats, &rg.bars[rg.fill], lane);
            rg.fill = (rg.fill + 1 == kStages) ? 0 : rg.fill + 1;
            prod.next();
        }
        mbar_wait(&rg.bars[rg.use], (rg.par >> rg.use) & 1u);
        rg.par ^= 1u << rg.use;
        float *stage = rg.buf + rg.use * kStageFloats;
        const int len = cons.len();
        if (len < kTileJ) {
            // the rotation visits all 64 pairs: fill the rest of a short last tile with sources at
            // infinity and mass 0, which contribute exactly zero to rows and columns
            for (int e = ((len + 3) & ~3) + lane; e < kTileJ; e += 32) {
                stage[e] = __int_as_float(0x7f800000);
                stage[kTileJ + e] = __int_as_float(0x7f800000);
                stage[2 * kTileJ + e] = __int_as_float(0x7f800000);
                stage[3 * kTileJ + e] = 0.f;
            }
            fence_proxy_async_smem();
        }
#pragma unroll
        for (int e = lane; e < kTileJ; e += 32) col[e] = 0.f;
        __syncwarp();
        sym_tile<TR>(stage, col, xi, yi, zi, mi, acc64, lane);
        __syncwarp();
        if (TR == 8 && intra && first) {
            // the pairs inside the row unit are done: its second tile joins the rows for the remaining columns
#pragma unroll
            for (int t = 4; t < TR; ++t) {
                const int64_t i = base + static_cast<int64_t>(I0 + 1) * kTileJ + (t & 3) * 32 + lane;
                xi[t] = sp.x[i];
                yi[t] = sp.y[i];
                zi[t] = sp.z[i];
                mi[t] = sp.m[i];
            }
        }
        first = false;
        const int64_t tile0 = cons.base + cons.pos;
#pragma unroll
        for (int e = lane; e < kTileJ; e += 32)
            if (e < len) {
                const float vf = col[e];
                bad |= !(fabsf(vf) <= 3.4028234e38f);
                bad |= sym_add(&p.phi_sym[tile0 + e], static_cast<double>(vf), q, inv_q, window);
            }
        __syncwarp();      // every lane is done with this stage and with col before they are reused
        rg.use = (rg.use + 1 == kStages) ? 0 : rg.use + 1;
        cons.next();
    }
#pragma unroll
    for (int t = 0; t < TR; ++t) {
        if (t >= 4 && !pair) continue;                    // masked rows of a single-tile unit
        bad |= !(fabs(acc64[t]) <= 1.7976931348623157e308);
        bad |= sym_add(&p.phi_sym[base + static_cast<int64_t>(I0 + (t >> 2)) * kTileJ + (t & 3) * 32 + lane], acc64[t], q,
                       inv_q, window);
    }
    // a zero separation between different tiles (exact duplicates) or non-finite input: the
    // predicated kernel recomputes the halo from scratch and phi_sym is ignored for it
    if (__any_sync(0xffffffffu, bad) && lane == 0) {
        atomicExch(&p.halo_redo[tk.h], 1);
        atomicExch(&p.st->redo_any, 1);
    }
}

// Per-warp state of the potential code: the TMA ring (continues across tickets and passes) and the column
// partial sums of the symmetric tickets.  smem_raw: kSmemBytes of dynamic shared memory.
__device__ __forceinline__ void warp_ring_setup(unsigned char *smem_raw, Ring &rg, float *&col)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    rg.buf = reinterpret_cast<float *>(smem_raw) + warp * (kStages * kStageFloats);
    rg.bars = reinterpret_cast<uint64_t *>(smem_raw + kWarpsPerBlock * kStages * kStageFloats * 4) + warp * kStages;
    rg.par = 0;
    rg.fill = rg.use = 0;
    col = reinterpret_cast<float *>(smem_raw + kWarpsPerBlock * kStages * (kStageFloats * 4 + 8)) + warp * kTileJ;
    if (lane == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&rg.bars[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
}

// One potential pass of the FAST kernels: every warp takes tickets until the pass's table is exhausted.
// REUSE: the instantiations that know about the external-sum cache and the incremental passes; plans
// without them run the REUSE = false kernels, whose code is what it was before those existed.
// redo_only: the predicated re-evaluation of the haloes the predicate-free pass flagged.
// SYM: 0 = no symmetric tickets, 4 / 8 = symmetric tickets with that many row members per lane.
template <int T, bool NP, int SYM, bool REUSE>
__device__ __forceinline__ void potential_pass_fast(const PotParams &p, Ring &rg, float *col, const bool redo_only)
{
    const LoopState *st = p.st;
    const int lane = threadIdx.x & 31;
    static_assert(SYM == 0 || ((SYM == 4 || SYM == 8) && NP && T == 4), "symmetric tickets use 128-member tiles on the predicate-free kernel");

    const int n_items = st->n_items;
    const int parity = st->parity;
    constexpr int kGroup = 32 * T;

    for (;;) {
        int item = 0;
        if (lane == 0)
            item = static_cast<int>(atomicAdd(redo_only ? &p.st->counter_redo : &p.st->counter, 1u));
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        Ticket tk;
        if (!decode_ticket<kGroup>(p, item, tk)) continue;
        if (tk.corr_axis == 3) {
            if constexpr (SYM != 0) sym_ticket<SYM>(p, tk, rg, col, lane, parity);
            continue;
        }
        if (tk.corr_axis >= 0) {
            if (NP) {
                correction_ticket<T, 0>(p, tk, rg, lane);
                correction_ticket<T, 1>(p, tk, rg, lane);
                correction_ticket<T, 2>(p, tk, rg, lane);
            }
            continue;
        }
        if (redo_only && !p.halo_redo[tk.h]) continue;
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
        }
        // Source phases of a main ticket.  Normally one: every segment of the halo.
        // External cache: the members segment alone (filter 1) and, in the first pass only, the external
        // segments on their own (filter 2) into phi_ext, which k_energy_flag adds in every later pass.
        // Incremental pass: only the members the previous pass removed, WITH the reference's predicate, so the
        // sum is exactly what those members contributed to the potential k_energy_flag kept from the previous
        // pass (externals and corrections included) and no correction ticket is needed.
        const bool incr = REUSE && p.incr_enabled && p.incr[tk.h];
        int filt = 0, nphase = 1;
        if (p.targets_only) {
            // the f2py-level call with targets that are not sources (api.cu::potential_via_plan): the members
            // carry no mass, only the external segments are summed
            filt = 2;
        } else if (REUSE && p.cache_ext && hd->n_ext > 0 && !incr) {
            if (st->pass == 0) {
                filt = 1;
                nphase = 2;
            } else if (p.ext_ok[tk.h]) {
                filt = 1;
            }
        }
#pragma unroll 1
        for (int ph = 0; ph < nphase; ++ph) {
#pragma unroll
            for (int t = 0; t < T; ++t) acc64[t] = 0.0;
            TileCursor cur;
            if (incr) {
                const int nr = p.rem_cnt[tk.h];
                const int per = (((nr + tk.S - 1) / tk.S) + kTileJ - 1) / kTileJ * kTileJ;
                const int a = min(tk.s * per, nr), b = min(a + per, nr);
                cur.init_range(6, hd->poff + a, b - a);
            } else {
                cur.init(hd, tk.S, tk.s, tk.n_tgt, parity, (SYM && p.sym_enabled) ? tk.group : -1, filt + ph);
            }
            if (REUSE && incr)
                run_tiles<T, 0>(p, cur, rg, lane, 0, 0, xi, yi, zi, acc64);
            else if (NP)
                run_tiles<T, 1>(p, cur, rg, lane, tk.group * kGroup, tk.group * kGroup + kGroup, xi, yi, zi, acc64);
            else
                run_tiles<T, 0>(p, cur, rg, lane, 0, 0, xi, yi, zi, acc64);
            if (NP || REUSE) {
                bool bad = false;
                if (NP) {
                    // zero separations outside the own tile (exact duplicates) or non-finite input:
                    // hand the halo to the predicated kernel
#pragma unroll
                    for (int t = 0; t < T; ++t)
                        bad |= (tk.group * kGroup + t * 32 + lane < tk.n_tgt) && !(fabs(acc64[t]) <= 1.7976931348623157e308);
                }
                if (REUSE && incr) {
                    // energy_phase will subtract this sum from the potential it kept; the difference carries the
                    // rounding of the float32 partial sums in here, ~1e-7 of the sum.  That is far inside the
                    // tolerance unless the removed members made up most of a member's potential (a tight pair
                    // that lost its partner): removed > kIncrHeavy of the last fully evaluated potential, tested
                    // per j-split piece (conservatively: a total beyond the limit has a piece beyond limit / S).
                    // Such a halo is recomputed from scratch by the predicated kernel in this same pass.
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const int i = tk.group * kGroup + t * 32 + lane;
                        if (i < tk.n_tgt) {
                            const int64_t os = hd->poff + (p.widx[parity][tbase + i] - hd->uoff);
                            // what is left of the last fully evaluated potential once this pass's share is out:
                            // chains of incremental passes stay inside the same bound as a single one
                            bad |= p.phi_keep[os] - fabs(acc64[t]) * tk.S < (1.0 - kIncrHeavy) * p.phi_full[os];
                        }
                    }
                }
                if (__any_sync(0xffffffffu, bad) && lane == 0) {
                    atomicExch(&p.halo_redo[tk.h], 1);
                    atomicExch(&p.st->redo_any, 1);
                }
            }
            // (the first pass works on the uncompacted buffer, so tbase + i is the member's original slot)
            double *out = (ph == 0 ? p.phi_part : p.phi_ext) + static_cast<int64_t>(tk.s) * p.phi_stride + tbase;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int i = tk.group * kGroup + t * 32 + lane;
                if (i < tk.n_tgt) out[i] = acc64[t];
            }
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

// The same term through exact_term_fast (halma_common.cuh): bit-identical while the operands are
// in the safe window, which `bad` reports for pairs that pass the predicate.  The predicate stays
// the literal != on the inputs (a three-input min of |d| would differ for NaN and Inf coordinates).
__device__ __forceinline__ float term_exact_fast(float xs, float ys, float zs, float m, float xi, float yi, float zi,
                                                 bool &bad)
{
    const float dx = __fsub_rn(xs, xi);
    const float dy = __fsub_rn(ys, yi);
    const float dz = __fsub_rn(zs, zi);
    const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const bool take = (xs != xi) && (ys != yi) && (zs != zi);
    bad |= take && !exact_r2_safe(r2);
    const float q = exact_term_fast(m, r2);
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
    // masses are per source: every lane vets four of the tile's 128 once, instead of every
    // lane vetting every source
    const int lane = threadIdx.x & 31;
    const float4 mq = M[lane];
    const bool m_ok = lane >= nq || (exact_mass_safe(mq.x) && exact_mass_safe(mq.y) && exact_mass_safe(mq.z) &&
                                     exact_mass_safe(mq.w));
    const bool fast_ok = __all_sync(0xffffffffu, m_ok);
#pragma unroll 2
    for (int q = 0; q < nq; ++q) {
        const float4 x4 = X[q], y4 = Y[q], z4 = Z[q], m4 = M[q];
        float t0, t1, t2, t3;
        bool bad = !fast_ok;
        if (fast_ok) {
            t0 = term_exact_fast(x4.x, y4.x, z4.x, m4.x, xi, yi, zi, bad);
            t1 = term_exact_fast(x4.y, y4.y, z4.y, m4.y, xi, yi, zi, bad);
            t2 = term_exact_fast(x4.z, y4.z, z4.z, m4.z, xi, yi, zi, bad);
            t3 = term_exact_fast(x4.w, y4.w, z4.w, m4.w, xi, yi, zi, bad);
        }
        if (__any_sync(0xffffffffu, bad)) {          // an operand outside the window: IEEE library path
            t0 = term_exact(x4.x, y4.x, z4.x, m4.x, xi, yi, zi);
            t1 = term_exact(x4.y, y4.y, z4.y, m4.y, xi, yi, zi);
            t2 = term_exact(x4.z, y4.z, z4.z, m4.z, xi, yi, zi);
            t3 = term_exact(x4.w, y4.w, z4.w, m4.w, xi, yi, zi);
        }
        acc = __fadd_rn(acc, t0);      // ascending source order, one rounding per add
        acc = __fadd_rn(acc, t1);
        acc = __fadd_rn(acc, t2);
        acc = __fadd_rn(acc, t3);
    }
}

// One potential pass of the EXACT kernel (the ring is the same per-warp ring as the FAST kernels').
__device__ __forceinline__ void potential_pass_exact(const PotParams &p, Ring &rg)
{
    const LoopState *st = p.st;
    const int lane = threadIdx.x & 31;
    float *ring = rg.buf;
    uint64_t *bars = rg.bars;
    const int n_items = st->n_items;
    const int parity = st->parity;
    uint32_t par = rg.par;
    int fill = rg.fill, use = rg.use;

    for (;;) {
        int item = 0;
        if (lane == 0) item = static_cast<int>(atomicAdd(&p.st->counter, 1u));
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        Ticket tk;
        if (!decode_ticket<32>(p, item, tk) || tk.corr_axis >= 0) continue;
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
    rg.par = par;
    rg.fill = fill;
    rg.use = use;
}

}  // namespace halma
