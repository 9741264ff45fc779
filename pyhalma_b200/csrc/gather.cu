// The gather step before the hot path: AMR cells -> gas pseudo-particles and
// DM / star particles inside a sphere, from a snapshot that stays resident in HBM.
//
// Replaces, per halo, python_scripts/halo_gas.py:223-277 (st_gas_dm_particles_inside) with its
// callees :9-52 (patch_to_particles, a numba triple loop per patch), :56-141
// (AMRgrid_to_particles, np.append per patch), :216-218 (parallel_inside) and the two
// KDTree.query_ball_point calls (:255, :269).
//
// Layout: the cell fields of all patches of level >= 1 are concatenated (patch after patch, each
// C-ordered with ix slowest = the reference's loop nest) into 5 float32 + 2 uint8 arrays;
// particles are float64 SoA.  A snapshot is uploaded once and queried for every halo.
//
// A gather is three order-preserving selections (cells, DM, stars), each = count per 256-item
// chunk -> exclusive scan of the chunk counts -> emit with warp ballots, so the output keeps
// the reference's order (ascending patch, ix, iy, iz; ascending particle index).  Only the
// cells of the sub-box each patch shares with the query box are visited.  HBM-bound:
//   cells      2 B flags per candidate cell (twice) + 20 B fields + 64 B written per selected cell
//   particles  24 B per particle (twice) + 32..40 B read and written per selected particle
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/halma_unbind.h"
#include "aux_timer.h"
#include "halma_common.cuh"

int halma_internal_ctx(int device, int *sm_count, cudaStream_t *stream);
int halma_internal_fail(int code, const char *msg);

#define GA_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess) return halma_internal_fail(HALMA_ERR_CUDA, cudaGetErrorString(e__));          \
    } while (0)

namespace halma {
namespace {

constexpr int kSelBlock = 256;

struct PatchDesc {
    double x0, y0, z0, res;      // centre of the first cell (halo_gas.py:27-29) and the cell size
    int64_t cell_off;            // offset of the patch's cells in the concatenated fields
    int32_t nx, ny, nz, active;  // active: level >= 1 (halo_gas.py:107) and uploaded
};

struct SubBox {
    int32_t i0, j0, k0, ni, nj, nk;
};

struct Query {
    double box[6];               // cx-R, cx+R, cy-R, cy+R, cz-R, cz+R   (halo_gas.py:87-88)
    double cx, cy, cz, R, R2;
    double mass_factor[2];       // rho_B, rete^3
    int box_only;                // gas: the box test alone (AMRgrid_to_particles, halo_gas.py:56-141)
};

// One thread per patch: the index range of the cells whose centres can lie strictly inside
// the box (one cell of slack on both sides; the exact test is made per cell).
__global__ void k_patch_clip(const PatchDesc *__restrict__ pd, int64_t n_patch, const Query q,
                             SubBox *__restrict__ sub, int64_t *__restrict__ cand)
{
    const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (p >= n_patch) return;
    const PatchDesc d = pd[p];
    SubBox s = {0, 0, 0, 0, 0, 0};
    int64_t c = 0;
    if (d.active) {
        const double o[3] = {d.x0, d.y0, d.z0};
        const int n[3] = {d.nx, d.ny, d.nz};
        int lo[3], cnt[3];
        bool any = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double fl = floor((q.box[2 * a] - o[a]) / d.res) - 1.0;
            const double fh = ceil((q.box[2 * a + 1] - o[a]) / d.res) + 1.0;
            const double l = fmax(fl, 0.0), h = fmin(fh, static_cast<double>(n[a] - 1));
            if (!(h >= l)) {
                any = false;
                lo[a] = cnt[a] = 0;
            } else {
                lo[a] = static_cast<int>(l);
                cnt[a] = static_cast<int>(h) - lo[a] + 1;
            }
        }
        if (any) {
            s.i0 = lo[0]; s.j0 = lo[1]; s.k0 = lo[2];
            s.ni = cnt[0]; s.nj = cnt[1]; s.nk = cnt[2];
            c = static_cast<int64_t>(cnt[0]) * cnt[1] * cnt[2];
        }
    }
    sub[p] = s;
    cand[p] = c;
}

// ---- selections ------------------------------------------------------------------------------
struct CellSel {
    const PatchDesc *pd;
    const SubBox *sub;
    const int64_t *cand_off;      // exclusive scan of the candidate counts, [n_patch + 1]
    int64_t n_patch;
    const float *delta, *vx, *vy, *vz, *temp;
    const uint8_t *cr0amr, *solapst;
    Query q;
    double *out[8];               // x, y, z, vx, vy, vz, mass, temp

    struct Item {
        double x, y, z;
        int64_t cell;
        double res;
    };

    __device__ __forceinline__ bool test(int64_t t, Item &it) const
    {
        // patch of candidate t: largest p with cand_off[p] <= t
        int64_t lo = 0, hi = n_patch;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (cand_off[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        const PatchDesc d = pd[lo];
        const SubBox s = sub[lo];
        const int64_t r = t - cand_off[lo];
        const int plane = s.nj * s.nk;
        const int ix = s.i0 + static_cast<int>(r / plane);
        const int rem = static_cast<int>(r % plane);
        const int iy = s.j0 + rem / s.nk, iz = s.k0 + rem % s.nk;
        // halo_gas.py:31,34,37: x = x0 + ix*patch_res, no contraction
        it.x = __dadd_rn(d.x0, __dmul_rn(static_cast<double>(ix), d.res));
        it.y = __dadd_rn(d.y0, __dmul_rn(static_cast<double>(iy), d.res));
        it.z = __dadd_rn(d.z0, __dmul_rn(static_cast<double>(iz), d.res));
        if (!(it.x > q.box[0] && it.x < q.box[1] && it.y > q.box[2] && it.y < q.box[3] && it.z > q.box[4] &&
              it.z < q.box[5]))
            return false;                                                       // :32,35,38
        it.cell = d.cell_off + (static_cast<int64_t>(ix) * d.ny + iy) * d.nz + iz;
        it.res = d.res;
        if (!(cr0amr[it.cell] && solapst[it.cell])) return false;               // :40
        // :216-218, :236  sqrt((x-cx)^2 + (y-cy)^2 + (z-cz)^2) < R
        const double dx = __dsub_rn(it.x, q.cx), dy = __dsub_rn(it.y, q.cy), dz = __dsub_rn(it.z, q.cz);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        return q.box_only || __dsqrt_rn(d2) < q.R;
    }
    __device__ __forceinline__ void emit(const Item &it, int64_t o) const
    {
        out[0][o] = it.x;
        out[1][o] = it.y;
        out[2][o] = it.z;
        out[3][o] = __dmul_rn(static_cast<double>(vx[it.cell]), 3e5);            // :136-138
        out[4][o] = __dmul_rn(static_cast<double>(vy[it.cell]), 3e5);
        out[5][o] = __dmul_rn(static_cast<double>(vz[it.cell]), 3e5);
        // :48  (1 + delta) * rho_B * res**3 (numba: res*res*res, left to right), then :246 *= rete**3
        const double res3 = __dmul_rn(__dmul_rn(it.res, it.res), it.res);
        const double m = __dmul_rn(__dmul_rn(__dadd_rn(1.0, static_cast<double>(delta[it.cell])), q.mass_factor[0]), res3);
        out[6][o] = __dmul_rn(m, q.mass_factor[1]);
        out[7][o] = static_cast<double>(temp[it.cell]);
    }
};

struct BallSel {
    const double *x, *y, *z, *mass;
    const int64_t *id;            // may be null
    Query q;
    int species;                  // 0 all; 1 heavy: mass >= m_split; 2 light: !(mass >= m_split)   (halo_gas.py:347,362)
    double m_split;
    double *out[4];
    int64_t *out_id;

    struct Item {
        int64_t i;
    };
    __device__ __forceinline__ bool test(int64_t i, Item &it) const
    {
        it.i = i;
        // KDTree.query_ball_point: squared distance <= R^2 (halo_gas.py:255,269)
        const double dx = __dsub_rn(x[i], q.cx), dy = __dsub_rn(y[i], q.cy), dz = __dsub_rn(z[i], q.cz);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (!(d2 <= q.R2)) return false;
        if (species == 0) return true;
        const bool heavy = mass[i] >= m_split;    // NaN is light, like np.logical_not(condition_l01)
        return species == 1 ? heavy : !heavy;
    }
    __device__ __forceinline__ void emit(const Item &it, int64_t o) const
    {
        out[0][o] = x[it.i];
        out[1][o] = y[it.i];
        out[2][o] = z[it.i];
        out[3][o] = mass[it.i];
        if (out_id) out_id[o] = id ? id[it.i] : it.i;
    }
};

// ---- uniform-grid index over the resident particles ---------------------------------------------
// The brute-force ball query reads every resident particle for every halo.  Above
// kIndexMinParticles the particles get a cell list instead: G^3 cells over their bounding box,
// `perm` = particle indices sorted by cell (x fastest; stable, so ascending inside a cell) and
// `cell_start`.  The cells a query box overlaps are, per (iy, iz) row, ONE contiguous range of
// `perm`; only those candidates are tested, with the same predicate as BallSel, and the selected
// indices are sorted so the output is the brute-force path's (ascending index), bit for bit.
struct GridGeom {
    double lo[3], inv_h[3];       // cell = clamp(floor((x - lo) * inv_h), 0, G - 1)
    int32_t G;
};

__device__ __forceinline__ int grid_cell(const GridGeom &g, int a, double v)
{
    const double f = floor(__dmul_rn(__dsub_rn(v, g.lo[a]), g.inv_h[a]));
    return static_cast<int>(fmin(fmax(f, 0.0), static_cast<double>(g.G - 1)));      // NaN -> 0
}

__global__ void __launch_bounds__(256) k_minmax(const double *__restrict__ x, const double *__restrict__ y,
                                               const double *__restrict__ z, int64_t n, double *__restrict__ part6)
{
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double v[6] = {inf, inf, inf, -inf, -inf, -inf};
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double p[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            v[a] = fmin(v[a], p[a]);
            v[3 + a] = fmax(v[3 + a], p[a]);
        }
    }
    __shared__ double red[6][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double t = v[k];
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_down_sync(0xffffffffu, t, o);
            t = k < 3 ? fmin(t, u) : fmax(t, u);
        }
        if (lane == 0) red[k][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = red[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) t = threadIdx.x < 3 ? fmin(t, red[threadIdx.x][w]) : fmax(t, red[threadIdx.x][w]);
        part6[blockIdx.x * 6 + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_cell_keys(const double *__restrict__ x, const double *__restrict__ y,
                                                  const double *__restrict__ z, int64_t n, const GridGeom g,
                                                  uint32_t *__restrict__ key, int32_t *__restrict__ idx)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ix = grid_cell(g, 0, x[i]), iy = grid_cell(g, 1, y[i]), iz = grid_cell(g, 2, z[i]);
        key[i] = static_cast<uint32_t>((iz * g.G + iy) * g.G + ix);
        idx[i] = static_cast<int32_t>(i);
    }
}

// cell_start[c] = first position of the sorted keys with key >= c, for c = 0 .. G^3
__global__ void __launch_bounds__(256) k_cell_start(const uint32_t *__restrict__ sorted_key, int64_t n, int64_t n_cells,
                                                   int32_t *__restrict__ cell_start)
{
    for (int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c <= n_cells;
         c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_key[mid] < c)
                lo = mid + 1;
            else
                hi = mid;
        }
        cell_start[c] = static_cast<int32_t>(lo);
    }
}

struct RowRange {
    int32_t ix0, ix1, iy0, ny, iz0, nz, G;
};

// candidates of every (iy, iz) row of the query's cell box
__global__ void __launch_bounds__(256) k_row_len(const int32_t *__restrict__ cell_start, const RowRange r,
                                                int64_t *__restrict__ len)
{
    const int rows = r.ny * r.nz;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= rows; t += gridDim.x * blockDim.x) {
        int64_t v = 0;
        if (t < rows) {
            const int iy = r.iy0 + t % r.ny, iz = r.iz0 + t / r.ny;
            const int64_t base = (static_cast<int64_t>(iz) * r.G + iy) * r.G;
            v = cell_start[base + r.ix1 + 1] - cell_start[base + r.ix0];
        }
        len[t] = v;
    }
}

struct IndexedBallSel {
    BallSel b;
    const int32_t *perm, *cell_start;
    const int64_t *row_off;       // exclusive scan of the row lengths, [rows + 1]
    RowRange r;
    int32_t *out_idx;

    typedef BallSel::Item Item;
    __device__ __forceinline__ bool test(int64_t t, Item &it) const
    {
        int lo = 0, hi = r.ny * r.nz;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (row_off[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        const int iy = r.iy0 + lo % r.ny, iz = r.iz0 + lo / r.ny;
        const int64_t base = (static_cast<int64_t>(iz) * r.G + iy) * r.G;
        return b.test(perm[cell_start[base + r.ix0] + (t - row_off[lo])], it);
    }
    __device__ __forceinline__ void emit(const Item &it, int64_t o) const { out_idx[o] = static_cast<int32_t>(it.i); }
};

__global__ void __launch_bounds__(256) k_emit_rows(const BallSel b, const int32_t *__restrict__ idx, int64_t k)
{
    for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < k;
         o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        BallSel::Item it;
        it.i = idx[o];
        b.emit(it, o);
    }
}

template <class Sel>
__global__ void __launch_bounds__(kSelBlock) k_sel_count(const Sel sel, int64_t n, int64_t *__restrict__ chunk_cnt)
{
    const int64_t nchunk = (n + kSelBlock - 1) / kSelBlock;
    for (int64_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
        const int64_t i = c * kSelBlock + threadIdx.x;
        typename Sel::Item it;
        const bool f = i < n && sel.test(i, it);
        const int cnt = __syncthreads_count(f);
        if (threadIdx.x == 0) chunk_cnt[c] = cnt;
    }
}

// Stable: rank inside the chunk = survivors in lower warps + survivors in lower lanes.
template <class Sel>
__global__ void __launch_bounds__(kSelBlock) k_sel_emit(const Sel sel, int64_t n, const int64_t *__restrict__ chunk_off)
{
    __shared__ int warp_cnt[kSelBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nchunk = (n + kSelBlock - 1) / kSelBlock;
    for (int64_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
        const int64_t i = c * kSelBlock + threadIdx.x;
        typename Sel::Item it;
        const bool f = i < n && sel.test(i, it);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[warp] = __popc(b);
        __syncthreads();
        if (f) {
            int below = __popc(b & ((1u << lane) - 1u));
            for (int w = 0; w < warp; ++w) below += warp_cnt[w];
            sel.emit(it, chunk_off[c] + below);
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace halma

using namespace halma;

struct halma_snapshot {
    int device = 0, sm = 0;
    cudaStream_t stream = nullptr;
    double L = 0;
    int32_t ncoarse = 0;
    int64_t n_patch = 0, n_cells = 0;
    std::vector<PatchDesc> host_pd;
    std::vector<uint8_t> uploaded;
    PatchDesc *d_pd = nullptr;
    SubBox *d_sub = nullptr;
    int64_t *d_cand = nullptr;          // [2 * (n_patch + 1)]: counts, then their exclusive scan
    float *d_fields = nullptr;          // 5 x n_cells
    uint8_t *d_flags = nullptr;         // 2 x n_cells
    bool pd_dirty = true;
    // particles: 0 DM, 1 stars
    int64_t n_part[2] = {0, 0};
    double *d_part[2] = {nullptr, nullptr};     // 4 x n
    int64_t *d_id[2] = {nullptr, nullptr};
    // uniform-grid index per particle kind (built above kIndexMinParticles)
    bool indexed[2] = {false, false};
    GridGeom geom[2];
    int32_t *d_perm[2] = {nullptr, nullptr};
    int32_t *d_cell_start[2] = {nullptr, nullptr};
    // scratch
    void *d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int64_t *d_chunk = nullptr;         // counts then offsets
    size_t chunk_cap = 0;
    // results of the last gather
    // gas 8 x n; DM 4 x n (all of it, or the heavy species when a mass split was asked for);
    // light DM 4 x n; stars 4 x n
    int64_t n_out[4] = {0, 0, 0, 0};
    double *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t *d_out_id = nullptr;
    bool have_result = false;
};

namespace {

// Stream-ordered scratch that is returned to the pool on every exit path.
struct Scratch {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    explicit Scratch(cudaStream_t stream) : s(stream) {}
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, s); }
    template <class T>
    T *as() const { return static_cast<T *>(p); }
    ~Scratch()
    {
        if (p) cudaFreeAsync(p, s);
    }
};

int ensure_scan_tmp(halma_snapshot *s, int64_t n)
{
    size_t need = 0;
    GA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, static_cast<int64_t *>(nullptr), static_cast<int64_t *>(nullptr),
                                         static_cast<int>(n), s->stream));
    if (need > s->scan_tmp_bytes) {
        if (s->d_scan_tmp) GA_TRY(cudaFreeAsync(s->d_scan_tmp, s->stream));
        GA_TRY(cudaMallocAsync(&s->d_scan_tmp, need, s->stream));
        s->scan_tmp_bytes = need;
    }
    return HALMA_OK;
}

int ensure_chunks(halma_snapshot *s, int64_t nchunk)
{
    const size_t need = static_cast<size_t>(nchunk + 1);
    if (need > s->chunk_cap) {
        if (s->d_chunk) GA_TRY(cudaFreeAsync(s->d_chunk, s->stream));
        GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_chunk), 2 * need * sizeof(int64_t), s->stream));
        s->chunk_cap = need;
    }
    return HALMA_OK;
}

void free_results(halma_snapshot *s)
{
    for (int k = 0; k < 4; ++k) {
        if (s->d_out[k]) cudaFreeAsync(s->d_out[k], s->stream);
        s->d_out[k] = nullptr;
        s->n_out[k] = 0;
    }
    if (s->d_out_id) cudaFreeAsync(s->d_out_id, s->stream);
    s->d_out_id = nullptr;
    s->have_result = false;
}

// count -> scan -> (host learns the total) -> allocate -> emit
template <class Sel, class Alloc>
int run_selection(halma_snapshot *s, Sel &sel, int64_t n, int64_t *total, Alloc alloc_outputs)
{
    *total = 0;
    if (n <= 0) return alloc_outputs(sel, 0);
    if (n > (int64_t(1) << 40)) return halma_internal_fail(HALMA_ERR_TOO_LARGE, "selection too large");
    const int64_t nchunk = (n + kSelBlock - 1) / kSelBlock;
    if (nchunk + 1 > INT32_MAX) return halma_internal_fail(HALMA_ERR_TOO_LARGE, "selection too large");
    if (int rc = ensure_chunks(s, nchunk)) return rc;
    if (int rc = ensure_scan_tmp(s, nchunk + 1)) return rc;
    int64_t *cnt = s->d_chunk, *off = s->d_chunk + s->chunk_cap;
    const int blocks = static_cast<int>(std::min<int64_t>(nchunk, static_cast<int64_t>(s->sm) * 16));
    GA_TRY(cudaMemsetAsync(cnt + nchunk, 0, sizeof(int64_t), s->stream));
    k_sel_count<Sel><<<blocks, kSelBlock, 0, s->stream>>>(sel, n, cnt);
    GA_TRY(cudaGetLastError());
    size_t tmp = s->scan_tmp_bytes;
    GA_TRY(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, tmp, cnt, off, static_cast<int>(nchunk + 1), s->stream));
    GA_TRY(cudaMemcpyAsync(total, off + nchunk, sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
    GA_TRY(cudaStreamSynchronize(s->stream));
    if (int rc = alloc_outputs(sel, *total)) return rc;
    if (*total > 0) {
        k_sel_emit<Sel><<<blocks, kSelBlock, 0, s->stream>>>(sel, n, off);
        GA_TRY(cudaGetLastError());
    }
    return HALMA_OK;
}

constexpr int64_t kIndexMinParticles = 1 << 18;      // below this the brute-force pass is a few microseconds

int64_t index_min_particles()
{
    const char *e = getenv("HALMA_GATHER_INDEX_MIN");      // tests force the index on small inputs; -1 disables it
    return e ? static_cast<int64_t>(atoll(e)) : kIndexMinParticles;
}

int free_index(halma_snapshot *s, int kind)
{
    if (s->d_perm[kind]) GA_TRY(cudaFreeAsync(s->d_perm[kind], s->stream));
    if (s->d_cell_start[kind]) GA_TRY(cudaFreeAsync(s->d_cell_start[kind], s->stream));
    s->d_perm[kind] = nullptr;
    s->d_cell_start[kind] = nullptr;
    s->indexed[kind] = false;
    return HALMA_OK;
}

// Cell list of the particles of `kind` (positions already on the device).
int build_index(halma_snapshot *s, int kind)
{
    const int64_t n = s->n_part[kind];
    const int64_t min_n = index_min_particles();
    if (min_n < 0 || n < std::max<int64_t>(min_n, 1) || n >= INT32_MAX) return HALMA_OK;
    const size_t nn = static_cast<size_t>(n);
    const double *x = s->d_part[kind], *y = x + nn, *z = x + 2 * nn;
    // bounding box
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(s->sm) * 8));
    Scratch mm_buf(s->stream);
    GA_TRY(mm_buf.alloc(blocks * 6 * sizeof(double)));
    double *d_mm = mm_buf.as<double>();
    k_minmax<<<blocks, 256, 0, s->stream>>>(x, y, z, n, d_mm);
    GA_TRY(cudaGetLastError());
    std::vector<double> mm(static_cast<size_t>(blocks) * 6);
    GA_TRY(cudaMemcpyAsync(mm.data(), d_mm, mm.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    GA_TRY(cudaStreamSynchronize(s->stream));
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = mm[a];
        hi[a] = mm[3 + a];
        for (int b = 1; b < blocks; ++b) {
            lo[a] = std::fmin(lo[a], mm[static_cast<size_t>(b) * 6 + a]);
            hi[a] = std::fmax(hi[a], mm[static_cast<size_t>(b) * 6 + 3 + a]);
        }
        if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return HALMA_OK;      // infinite extent: keep brute force
    }
    GridGeom g;
    int G = static_cast<int>(std::lround(std::cbrt(static_cast<double>(n) / 4.0)));
    g.G = std::min(256, std::max(1, G));
    for (int a = 0; a < 3; ++a) {
        g.lo[a] = lo[a];
        const double ext = hi[a] - lo[a];
        g.inv_h[a] = ext > 0 && std::isfinite(g.G / ext) ? g.G / ext : 0.0;
    }
    const int64_t n_cells = static_cast<int64_t>(g.G) * g.G * g.G;
    Scratch key_buf(s->stream), key2_buf(s->stream), idx_buf(s->stream), tmp_buf(s->stream);
    GA_TRY(key_buf.alloc(nn * sizeof(uint32_t)));
    GA_TRY(key2_buf.alloc(nn * sizeof(uint32_t)));
    GA_TRY(idx_buf.alloc(nn * sizeof(int32_t)));
    uint32_t *d_key = key_buf.as<uint32_t>(), *d_key2 = key2_buf.as<uint32_t>();
    int32_t *d_idx = idx_buf.as<int32_t>();
    GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_perm[kind]), nn * sizeof(int32_t), s->stream));
    GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_cell_start[kind]), (n_cells + 1) * sizeof(int32_t), s->stream));
    k_cell_keys<<<blocks, 256, 0, s->stream>>>(x, y, z, n, g, d_key, d_idx);
    GA_TRY(cudaGetLastError());
    int bits = 1;
    while ((int64_t(1) << bits) < n_cells) ++bits;
    size_t tmp_bytes = 0;
    GA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_idx, s->d_perm[kind], static_cast<int>(n), 0,
                                           bits, s->stream));
    GA_TRY(tmp_buf.alloc(tmp_bytes));
    GA_TRY(cub::DeviceRadixSort::SortPairs(tmp_buf.p, tmp_bytes, d_key, d_key2, d_idx, s->d_perm[kind], static_cast<int>(n), 0,
                                           bits, s->stream));
    const int cb = static_cast<int>(std::min<int64_t>((n_cells + 256) / 256, static_cast<int64_t>(s->sm) * 16));
    k_cell_start<<<cb, 256, 0, s->stream>>>(d_key2, n, n_cells, s->d_cell_start[kind]);
    GA_TRY(cudaGetLastError());
    GA_TRY(cudaStreamSynchronize(s->stream));
    s->geom[kind] = g;
    s->indexed[kind] = true;
    return HALMA_OK;
}

int host_cell(const GridGeom &g, int a, double v)
{
    const double f = std::floor((v - g.lo[a]) * g.inv_h[a]);
    return static_cast<int>(std::fmin(std::fmax(f, 0.0), static_cast<double>(g.G - 1)));
}

// Ball selection through the cell list: candidates -> selected indices -> sorted -> rows.
template <class Alloc>
int run_indexed_ball(halma_snapshot *s, int kind, BallSel &bs, int64_t *total, Alloc alloc_outputs)
{
    *total = 0;
    const GridGeom &g = s->geom[kind];
    const Query &q = bs.q;
    if (!(q.R >= 0)) return alloc_outputs(bs, 0);          // negative or NaN radius selects nothing
    RowRange r;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        // one cell of slack: the exact test below decides
        lo[a] = std::max(0, host_cell(g, a, q.box[2 * a]) - 1);
        hi[a] = std::min(g.G - 1, host_cell(g, a, q.box[2 * a + 1]) + 1);
    }
    r.ix0 = lo[0]; r.ix1 = hi[0];
    r.iy0 = lo[1]; r.ny = hi[1] - lo[1] + 1;
    r.iz0 = lo[2]; r.nz = hi[2] - lo[2] + 1;
    r.G = g.G;
    const int64_t rows = static_cast<int64_t>(r.ny) * r.nz;
    Scratch len_buf(s->stream), sel_buf(s->stream), sort_buf(s->stream);
    GA_TRY(len_buf.alloc(2 * (rows + 1) * sizeof(int64_t)));
    int64_t *d_len = len_buf.as<int64_t>();
    int64_t *d_off = d_len + rows + 1;
    k_row_len<<<static_cast<int>((rows + 256) / 256), 256, 0, s->stream>>>(s->d_cell_start[kind], r, d_len);
    GA_TRY(cudaGetLastError());
    if (int rc = ensure_scan_tmp(s, rows + 1)) return rc;
    size_t tmp = s->scan_tmp_bytes;
    GA_TRY(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, tmp, d_len, d_off, static_cast<int>(rows + 1), s->stream));
    int64_t n_cand = 0;
    GA_TRY(cudaMemcpyAsync(&n_cand, d_off + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
    GA_TRY(cudaStreamSynchronize(s->stream));
    // a query that covers a large part of the particles is cheaper as one streaming pass
    if (n_cand * 3 > s->n_part[kind]) return run_selection(s, bs, s->n_part[kind], total, alloc_outputs);
    IndexedBallSel is;
    is.b = bs;
    is.perm = s->d_perm[kind];
    is.cell_start = s->d_cell_start[kind];
    is.row_off = d_off;
    is.r = r;
    is.out_idx = nullptr;
    int32_t *d_sel = nullptr, *d_sorted = nullptr;
    int rc = run_selection(s, is, n_cand, total, [&](IndexedBallSel &sel, int64_t k) -> int {
        if (k > 0) {
            GA_TRY(sel_buf.alloc(2 * k * sizeof(int32_t)));
            d_sel = sel_buf.as<int32_t>();
            d_sorted = d_sel + k;
            sel.out_idx = d_sel;
        }
        return HALMA_OK;
    });
    if (rc) return rc;
    const int64_t k = *total;
    if (int rc2 = alloc_outputs(bs, k)) return rc2;
    if (k > 0) {
        if (k >= INT32_MAX) return halma_internal_fail(HALMA_ERR_TOO_LARGE, "selection too large");
        size_t tb = 0;
        int bits = 1;
        while ((int64_t(1) << bits) < s->n_part[kind]) ++bits;
        const uint32_t *kin = reinterpret_cast<const uint32_t *>(d_sel);
        uint32_t *kout = reinterpret_cast<uint32_t *>(d_sorted);
        GA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tb, kin, kout, static_cast<int>(k), 0, bits, s->stream));
        GA_TRY(sort_buf.alloc(tb));
        GA_TRY(cub::DeviceRadixSort::SortKeys(sort_buf.p, tb, kin, kout, static_cast<int>(k), 0, bits, s->stream));
        const int eb = static_cast<int>(std::min<int64_t>((k + 255) / 256, static_cast<int64_t>(s->sm) * 16));
        k_emit_rows<<<eb, 256, 0, s->stream>>>(bs, d_sorted, k);
        GA_TRY(cudaGetLastError());
    }
    return HALMA_OK;
}

}  // namespace

extern "C" int halma_snapshot_create(int device, double L, int32_t ncoarse, int64_t n_patch, const int32_t *level,
                                     const int32_t *nx, const int32_t *ny, const int32_t *nz, const double *rx,
                                     const double *ry, const double *rz, halma_snapshot **out)
{
    if (!out) return halma_internal_fail(HALMA_ERR_INVALID, "out is null");
    *out = nullptr;
    if (n_patch < 0 || ncoarse <= 0 || !(L > 0)) return halma_internal_fail(HALMA_ERR_INVALID, "bad grid description");
    if (n_patch > 0 && (!level || !nx || !ny || !nz || !rx || !ry || !rz))
        return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    if (n_patch >= INT32_MAX) return halma_internal_fail(HALMA_ERR_TOO_LARGE, "too many patches");
    std::vector<PatchDesc> pd(static_cast<size_t>(n_patch));
    int64_t cells = 0;
    for (int64_t p = 0; p < n_patch; ++p) {
        PatchDesc &d = pd[p];
        memset(&d, 0, sizeof d);
        if (level[p] < 0 || level[p] > 60) return halma_internal_fail(HALMA_ERR_INVALID, "bad patch level");
        if (level[p] >= 1) {                                  // halo_gas.py:107: the base grid is never read
            if (nx[p] < 0 || ny[p] < 0 || nz[p] < 0) return halma_internal_fail(HALMA_ERR_INVALID, "negative patch extent");
            d.res = std::ldexp(L / ncoarse, -level[p]);       // :108  (L/ncoarse)/2**l, exact scaling
            d.x0 = rx[p] - d.res / 2;                         // :27-29
            d.y0 = ry[p] - d.res / 2;
            d.z0 = rz[p] - d.res / 2;
            d.nx = nx[p];
            d.ny = ny[p];
            d.nz = nz[p];
            d.cell_off = cells;
            cells += static_cast<int64_t>(nx[p]) * ny[p] * nz[p];
        }
    }
    halma_snapshot *s = new halma_snapshot;
    if (int rc = halma_internal_ctx(device, &s->sm, &s->stream)) {
        delete s;
        return rc;
    }
    s->device = device;
    s->L = L;
    s->ncoarse = ncoarse;
    s->n_patch = n_patch;
    s->n_cells = cells;
    s->host_pd.swap(pd);
    s->uploaded.assign(static_cast<size_t>(n_patch), 0);
    const size_t np1 = static_cast<size_t>(n_patch) + 1, nc = static_cast<size_t>(std::max<int64_t>(cells, 1));
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&s->d_pd), np1 * sizeof(PatchDesc), s->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void **>(&s->d_sub), np1 * sizeof(SubBox), s->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void **>(&s->d_cand), 2 * np1 * sizeof(int64_t), s->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void **>(&s->d_fields), 5 * nc * sizeof(float), s->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void **>(&s->d_flags), 2 * nc, s->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_flags, 0, 2 * nc, s->stream);
    if (e != cudaSuccess) {
        halma_snapshot_destroy(s);
        return halma_internal_fail(HALMA_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = s;
    return HALMA_OK;
}

extern "C" void halma_snapshot_destroy(halma_snapshot *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    free_results(s);
    void *ptrs[] = {s->d_pd, s->d_sub, s->d_cand, s->d_fields, s->d_flags, s->d_part[0], s->d_part[1],
                    s->d_id[0], s->d_id[1], s->d_scan_tmp, s->d_chunk, s->d_perm[0], s->d_perm[1],
                    s->d_cell_start[0], s->d_cell_start[1]};
    for (void *p : ptrs)
        if (p) cudaFreeAsync(p, s->stream);
    cudaStreamSynchronize(s->stream);
    delete s;
}

extern "C" int64_t halma_snapshot_cells(const halma_snapshot *s) { return s ? s->n_cells : -1; }

extern "C" int halma_snapshot_upload_patch(halma_snapshot *s, int64_t patch, const float *delta, const float *vx,
                                           const float *vy, const float *vz, const float *temp,
                                           const uint8_t *cr0amr, const uint8_t *solapst)
{
    if (!s) return halma_internal_fail(HALMA_ERR_INVALID, "snapshot is null");
    if (patch < 0 || patch >= s->n_patch) return halma_internal_fail(HALMA_ERR_INVALID, "patch index out of range");
    const PatchDesc &d = s->host_pd[patch];
    if (d.res == 0.0) return HALMA_OK;                        // level 0: never read
    const size_t n = static_cast<size_t>(d.nx) * d.ny * d.nz;
    if (n && (!delta || !vx || !vy || !vz || !temp || !cr0amr || !solapst))
        return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    GA_TRY(cudaSetDevice(s->device));
    const size_t nc = static_cast<size_t>(std::max<int64_t>(s->n_cells, 1));
    const float *src[5] = {delta, vx, vy, vz, temp};
    for (int k = 0; k < 5 && n; ++k)
        GA_TRY(cudaMemcpyAsync(s->d_fields + k * nc + d.cell_off, src[k], n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    if (n) {
        GA_TRY(cudaMemcpyAsync(s->d_flags + d.cell_off, cr0amr, n, cudaMemcpyHostToDevice, s->stream));
        GA_TRY(cudaMemcpyAsync(s->d_flags + nc + d.cell_off, solapst, n, cudaMemcpyHostToDevice, s->stream));
    }
    GA_TRY(cudaStreamSynchronize(s->stream));                 // the caller may reuse its buffers
    s->uploaded[patch] = 1;
    s->pd_dirty = true;
    return HALMA_OK;
}

extern "C" int halma_snapshot_upload_particles(halma_snapshot *s, int kind, int64_t n, const double *x, const double *y,
                                               const double *z, const double *mass, const int64_t *id)
{
    if (!s) return halma_internal_fail(HALMA_ERR_INVALID, "snapshot is null");
    if (kind < 0 || kind > 1 || n < 0) return halma_internal_fail(HALMA_ERR_INVALID, "bad kind or size");
    if (n && (!x || !y || !z || !mass)) return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    GA_TRY(cudaSetDevice(s->device));
    if (s->d_part[kind]) GA_TRY(cudaFreeAsync(s->d_part[kind], s->stream));
    if (s->d_id[kind]) GA_TRY(cudaFreeAsync(s->d_id[kind], s->stream));
    s->d_part[kind] = nullptr;
    s->d_id[kind] = nullptr;
    s->n_part[kind] = 0;
    if (int rc = free_index(s, kind)) return rc;
    if (n) {
        const size_t nn = static_cast<size_t>(n);
        GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_part[kind]), 4 * nn * sizeof(double), s->stream));
        const double *src[4] = {x, y, z, mass};
        for (int k = 0; k < 4; ++k)
            GA_TRY(cudaMemcpyAsync(s->d_part[kind] + k * nn, src[k], nn * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (id) {
            GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_id[kind]), nn * sizeof(int64_t), s->stream));
            GA_TRY(cudaMemcpyAsync(s->d_id[kind], id, nn * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
        }
        GA_TRY(cudaStreamSynchronize(s->stream));
    }
    s->n_part[kind] = n;
    return build_index(s, kind);
}

static int snapshot_gather(halma_snapshot *s, double cx, double cy, double cz, double R, double rho_B,
                           double mass_scale, double dm_heavy_min, int64_t *counts4, bool box_only);

extern "C" int halma_snapshot_gather(halma_snapshot *s, double cx, double cy, double cz, double R, double rho_B,
                                     double mass_scale, double dm_heavy_min, int64_t *counts4)
{
    return snapshot_gather(s, cx, cy, cz, R, rho_B, mass_scale, dm_heavy_min, counts4, false);
}

// AMRgrid_to_particles on its own (halo_gas.py:56-141, called at halo_properties.py:318): the gas cells whose
// centres lie strictly inside the BOX of half-width R, without the sphere test; no particles are selected.
extern "C" int halma_snapshot_gather_box(halma_snapshot *s, double cx, double cy, double cz, double R, double rho_B,
                                         double mass_scale, int64_t *counts4)
{
    return snapshot_gather(s, cx, cy, cz, R, rho_B, mass_scale, -std::numeric_limits<double>::infinity(), counts4, true);
}

static int snapshot_gather(halma_snapshot *s, double cx, double cy, double cz, double R, double rho_B,
                           double mass_scale, double dm_heavy_min, int64_t *counts4, bool box_only)
{
    if (!s || !counts4) return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    if (dm_heavy_min != dm_heavy_min) return halma_internal_fail(HALMA_ERR_INVALID, "dm_heavy_min is NaN");
    GA_TRY(cudaSetDevice(s->device));
    free_results(s);
    Query q;
    q.box[0] = cx - R; q.box[1] = cx + R;                     // halo_gas.py:87-88
    q.box[2] = cy - R; q.box[3] = cy + R;
    q.box[4] = cz - R; q.box[5] = cz + R;
    q.cx = cx; q.cy = cy; q.cz = cz;
    q.R = R;
    q.R2 = R * R;
    q.mass_factor[0] = rho_B;
    q.mass_factor[1] = mass_scale;
    q.box_only = box_only ? 1 : 0;
    KernelTimer timer(s->stream);
    // ---- gas cells ----
    if (s->pd_dirty) {
        for (int64_t p = 0; p < s->n_patch; ++p) s->host_pd[p].active = s->host_pd[p].res != 0.0 && s->uploaded[p];
        if (s->n_patch)
            GA_TRY(cudaMemcpyAsync(s->d_pd, s->host_pd.data(), s->n_patch * sizeof(PatchDesc), cudaMemcpyHostToDevice, s->stream));
        s->pd_dirty = false;
    }
    int64_t n_cand = 0;
    const int64_t np = s->n_patch;
    int64_t *cand = s->d_cand, *cand_off = s->d_cand + np + 1;
    if (np > 0) {
        k_patch_clip<<<static_cast<int>((np + 255) / 256), 256, 0, s->stream>>>(s->d_pd, np, q, s->d_sub, cand);
        GA_TRY(cudaGetLastError());
        GA_TRY(cudaMemsetAsync(cand + np, 0, sizeof(int64_t), s->stream));
        if (int rc = ensure_scan_tmp(s, np + 1)) return rc;
        size_t tmp = s->scan_tmp_bytes;
        GA_TRY(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, tmp, cand, cand_off, static_cast<int>(np + 1), s->stream));
        GA_TRY(cudaMemcpyAsync(&n_cand, cand_off + np, sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
        GA_TRY(cudaStreamSynchronize(s->stream));
    }
    const size_t nc = static_cast<size_t>(std::max<int64_t>(s->n_cells, 1));
    CellSel cs;
    cs.pd = s->d_pd;
    cs.sub = s->d_sub;
    cs.cand_off = cand_off;
    cs.n_patch = np;
    cs.delta = s->d_fields;
    cs.vx = s->d_fields + nc;
    cs.vy = s->d_fields + 2 * nc;
    cs.vz = s->d_fields + 3 * nc;
    cs.temp = s->d_fields + 4 * nc;
    cs.cr0amr = s->d_flags;
    cs.solapst = s->d_flags + nc;
    cs.q = q;
    int rc = run_selection(s, cs, n_cand, &s->n_out[0], [&](CellSel &sel, int64_t total) -> int {
        if (total > 0) {
            GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_out[0]), 8 * total * sizeof(double), s->stream));
            for (int k = 0; k < 8; ++k) sel.out[k] = s->d_out[0] + k * total;
        }
        return HALMA_OK;
    });
    if (rc) return rc;
    // ---- DM (one or two species) and stars ----
    const double inf = std::numeric_limits<double>::infinity();
    const bool split = dm_heavy_min > -inf;
    for (int sel_id = 1; sel_id < 4 && !box_only; ++sel_id) {
        if (sel_id == 2 && !split) continue;
        const int kind = sel_id == 3 ? 1 : 0;
        const int64_t n = s->n_part[kind];
        BallSel bs;
        const size_t nn = static_cast<size_t>(std::max<int64_t>(n, 1));
        bs.x = s->d_part[kind];
        bs.y = s->d_part[kind] ? s->d_part[kind] + nn : nullptr;
        bs.z = s->d_part[kind] ? s->d_part[kind] + 2 * nn : nullptr;
        bs.mass = s->d_part[kind] ? s->d_part[kind] + 3 * nn : nullptr;
        bs.id = s->d_id[kind];
        bs.q = q;
        bs.species = (split && sel_id != 3) ? sel_id : 0;
        bs.m_split = dm_heavy_min;
        bs.out_id = nullptr;
        auto alloc = [&](BallSel &sel, int64_t total) -> int {
            if (total > 0) {
                GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_out[sel_id]), 4 * total * sizeof(double), s->stream));
                for (int k = 0; k < 4; ++k) sel.out[k] = s->d_out[sel_id] + k * total;
                if (kind == 1) {
                    GA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&s->d_out_id), total * sizeof(int64_t), s->stream));
                    sel.out_id = s->d_out_id;
                }
            }
            return HALMA_OK;
        };
        rc = s->indexed[kind] ? run_indexed_ball(s, kind, bs, &s->n_out[sel_id], alloc)
                              : run_selection(s, bs, n, &s->n_out[sel_id], alloc);
        if (rc) return rc;
    }
    timer.stop();
    GA_TRY(cudaStreamSynchronize(s->stream));
    timer.publish();
    for (int k = 0; k < 4; ++k) counts4[k] = s->n_out[k];
    s->have_result = true;
    return HALMA_OK;
}

extern "C" int halma_snapshot_fetch(halma_snapshot *s, double *const *gas8, double *const *dm4, double *const *dml4,
                                    double *const *st4, int64_t *st_id)
{
    if (!s) return halma_internal_fail(HALMA_ERR_INVALID, "snapshot is null");
    if (!s->have_result) return halma_internal_fail(HALMA_ERR_STATE, "halma_snapshot_fetch before halma_snapshot_gather");
    GA_TRY(cudaSetDevice(s->device));
    double *const *dst[4] = {gas8, dm4, dml4, st4};
    const int ncol[4] = {8, 4, 4, 4};
    for (int g = 0; g < 4; ++g) {
        const int64_t n = s->n_out[g];
        if (!dst[g] || n == 0) continue;
        for (int k = 0; k < ncol[g]; ++k)
            if (dst[g][k])
                GA_TRY(cudaMemcpyAsync(dst[g][k], s->d_out[g] + k * n, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    if (st_id && s->n_out[3] > 0)
        GA_TRY(cudaMemcpyAsync(st_id, s->d_out_id, s->n_out[3] * sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
    GA_TRY(cudaStreamSynchronize(s->stream));
    return HALMA_OK;
}

// Device addresses of the last gather's result (column k of group g at ptr[g] + k * n_g), valid
// until the next gather or destroy.  The halma_plan_upload_* entry points take them as they are.
extern "C" int halma_snapshot_result_device(halma_snapshot *s, double **ptr4, int64_t **st_id, int64_t *counts4)
{
    if (!s || !ptr4) return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    if (!s->have_result) return halma_internal_fail(HALMA_ERR_STATE, "no gather result");
    for (int g = 0; g < 4; ++g) {
        ptr4[g] = s->d_out[g];
        if (counts4) counts4[g] = s->n_out[g];
    }
    if (st_id) *st_id = s->d_out_id;
    return HALMA_OK;
}

// Position, mass and id of gathered star k (the most bound one, halo_gas.py:629-634).
extern "C" int halma_snapshot_fetch_star(halma_snapshot *s, int64_t k, double *xyzm4, int64_t *id)
{
    if (!s || !xyzm4) return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    if (!s->have_result) return halma_internal_fail(HALMA_ERR_STATE, "no gather result");
    const int64_t n = s->n_out[3];
    if (k < 0 || k >= n) return halma_internal_fail(HALMA_ERR_INVALID, "star index out of range");
    GA_TRY(cudaSetDevice(s->device));
    for (int c = 0; c < 4; ++c)
        GA_TRY(cudaMemcpyAsync(xyzm4 + c, s->d_out[3] + c * n + k, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (id) GA_TRY(cudaMemcpyAsync(id, s->d_out_id + k, sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
    GA_TRY(cudaStreamSynchronize(s->stream));
    return HALMA_OK;
}
