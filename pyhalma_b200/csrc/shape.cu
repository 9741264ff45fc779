// halo_shape and sigma_projections: the two O(N) routines that share the
// f2py module `particle` with the potential kernel.  Both are HBM-bound reductions:
//   halo_shape          16 B per particle read once, 7 float64 sums
//   sigma_projections   (28 B + 4 B index) per particle read three times, 3 x 4 maps of
//                       n_cell^2 float64 updated with atomics (L2-resident)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/halma_unbind.h"
#include "aux_timer.h"
#include "halma_common.cuh"

int halma_internal_ctx(int device, int *sm_count, cudaStream_t *stream);
int halma_internal_fail(int code, const char *msg);

#define SH_TRY(expr)                                                          \
    do {                                                                      \
        cudaError_t e__ = (expr);                                             \
        if (e__ != cudaSuccess) return halma_internal_fail(HALMA_ERR_CUDA, cudaGetErrorString(e__)); \
    } while (0)

namespace halma {
namespace {

constexpr int kShBlock = 256;

template <int NV>
__device__ __forceinline__ void block_sum_atomic(double (&v)[NV], double *out)
{
    __shared__ double red[NV][kShBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < kShBlock / 32; ++w) s += red[threadIdx.x][w];
        atomicAdd(&out[threadIdx.x], s);
    }
}

// particle_subroutines.f90:188-199: sum m r_i r_j (upper triangle) and sum m.
// The four arrays start on 16-byte boundaries (the host pads the stride): 128-bit loads, four
// particles per thread and trip, 64 B in flight per thread.
__device__ __forceinline__ void inertia_add(double (&s)[7], float mf, float xf, float yf, float zf)
{
    const double mx = mf, a = xf, b = yf, c = zf;
    s[0] += mx * a * a;
    s[1] += mx * a * b;
    s[2] += mx * a * c;
    s[3] += mx * b * b;
    s[4] += mx * b * c;
    s[5] += mx * c * c;
    s[6] += mx;
}

__global__ void __launch_bounds__(kShBlock) k_inertia(const float *__restrict__ x, const float *__restrict__ y,
                                                      const float *__restrict__ z, const float *__restrict__ m,
                                                      int64_t n, double *out7)
{
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    const int64_t n4 = n >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x), *y4 = reinterpret_cast<const float4 *>(y),
                 *z4 = reinterpret_cast<const float4 *>(z), *m4 = reinterpret_cast<const float4 *>(m);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 a = x4[i], b = y4[i], c = z4[i], w = m4[i];
        inertia_add(s, w.x, a.x, b.x, c.x);
        inertia_add(s, w.y, a.y, b.y, c.y);
        inertia_add(s, w.z, a.z, b.z, c.z);
        inertia_add(s, w.w, a.w, b.w, c.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        inertia_add(s, m[i], x[i], y[i], z[i]);
    }
    // one partial per block, summed on the host in block order: bit-reproducible
    __shared__ double red[7][kShBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0.0;
        for (int w = 0; w < kShBlock / 32; ++w) v += red[threadIdx.x][w];
        out7[blockIdx.x * 7 + threadIdx.x] = v;
    }
}

struct SigmaParams {
    const float *grid;
    const int32_t *part_list;
    const float *x, *y, *z, *vx, *vy, *vz, *m;
    int64_t npart;
    int32_t n_cell;
    float cx, cy, cz, r05[3];
    double *vcm[3], *sd[3], *sig[3];     // final maps
    int32_t *cnt[3];
    // Replicated accumulation targets: block b adds into replica (b & rep_mask), so the float64
    // atomics of different blocks mostly hit different L2 lines (scripts/probes/atomic_hist_probe.cu:
    // 4.6 G updates/s with one copy of a 25 x 25 map, 94 G/s with 64); k_fold sums the replicas.
    double *racc;           // replica r: racc + r * 6 n^2: vcm x3 | sd x3 (pass 1), sig x3 (pass 2)
    int32_t *rcnt;          // replica r: rcnt + r * 3 n^2
    int32_t rep_mask;
    int32_t grid_sorted;    // grid is strictly increasing and free of NaN: binary search
    int16_t *cell;          // 3 per listed particle
    double *s05;            // 3 sums
    int32_t *c05;           // 3 counters
};

// minloc(abs(grid - d), dim = 1): nearest grid point, first minimum on ties (:267-269)
__device__ __forceinline__ int nearest_cell(const float *sgrid, int n_cell, float d)
{
    int best = 0;
    float bv = fabsf(sgrid[0] - d);
    for (int k = 1; k < n_cell; ++k) {
        const float v = fabsf(sgrid[k] - d);
        if (v < bv) {
            bv = v;
            best = k;
        }
    }
    return best;
}

// The same for a strictly increasing grid (checked on the host; pyHALMA.py:1001 builds it with
// np.arange): |grid - d| falls, then rises, so the first minimum is next to the first grid
// point >= d; a plateau of equal rounded differences to its left is walked back so the
// result is the linear scan's in every case.
__device__ __forceinline__ int nearest_cell_sorted(const float *sgrid, int n_cell, float d)
{
    int lo = 0, hi = n_cell;                 // first k with grid[k] >= d
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sgrid[mid] < d)
            lo = mid + 1;
        else
            hi = mid;
    }
    int best = lo > 0 ? lo - 1 : 0;
    float bv = fabsf(sgrid[best] - d);
    if (lo < n_cell && lo != best) {
        const float v = fabsf(sgrid[lo] - d);
        if (v < bv) {
            bv = v;
            best = lo;
        }
    }
    while (best > 0 && fabsf(sgrid[best - 1] - d) <= bv) {
        --best;
        bv = fabsf(sgrid[best] - d);
    }
    return best;
}

// pass 1 (:264-281): cell of every listed particle, mass-weighted velocity, mass and count maps
__global__ void __launch_bounds__(kShBlock) k_sigma_bin(const SigmaParams p)
{
    extern __shared__ float sgrid[];
    for (int k = threadIdx.x; k < p.n_cell; k += blockDim.x) sgrid[k] = p.grid[k];
    __syncthreads();
    const int nc = p.n_cell;
    const size_t nn = static_cast<size_t>(nc) * nc;
    double *acc = p.racc + (blockIdx.x & p.rep_mask) * (6 * nn);
    int32_t *cnt = p.rcnt + (blockIdx.x & p.rep_mask) * (3 * nn);
    for (int64_t ip = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; ip < p.npart;
         ip += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t q = p.part_list[ip];
        const float dx = p.x[q] - p.cx, dy = p.y[q] - p.cy, dz = p.z[q] - p.cz;
        int ix, iy, iz;
        if (p.grid_sorted && dx == dx && dy == dy && dz == dz) {
            ix = nearest_cell_sorted(sgrid, nc, dx);
            iy = nearest_cell_sorted(sgrid, nc, dy);
            iz = nearest_cell_sorted(sgrid, nc, dz);
        } else {
            ix = nearest_cell(sgrid, nc, dx);
            iy = nearest_cell(sgrid, nc, dy);
            iz = nearest_cell(sgrid, nc, dz);
        }
        p.cell[3 * ip] = static_cast<int16_t>(ix);
        p.cell[3 * ip + 1] = static_cast<int16_t>(iy);
        p.cell[3 * ip + 2] = static_cast<int16_t>(iz);
        const int64_t k3[3] = {iy + static_cast<int64_t>(nc) * iz, ix + static_cast<int64_t>(nc) * iz,
                               ix + static_cast<int64_t>(nc) * iy};
        const double m = p.m[q];
        const double v3[3] = {p.vx[q], p.vy[q], p.vz[q]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicAdd(&acc[a * nn + k3[a]], v3[a] * m);
            atomicAdd(&acc[(3 + a) * nn + k3[a]], m);
            atomicAdd(&cnt[a * nn + k3[a]], 1);
        }
    }
}

// dst[k] = sum over replicas, in replica order
template <class T>
__global__ void k_fold(T *__restrict__ dst, const T *__restrict__ src, int64_t n, int reps, int64_t stride)
{
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        T v = 0;
        for (int r = 0; r < reps; ++r) v += src[r * stride + k];
        dst[k] = v;
    }
}

// :286-288  VCM = VCM / SD where SD /= 0        :310-312  SIG = sqrt(SIG / count) where count /= 0
__global__ void k_sigma_maps(double *a, const double *den_d, const int32_t *den_i, int64_t n, int take_sqrt)
{
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (den_d) {
            if (den_d[k] != 0.0) a[k] = a[k] / den_d[k];
        } else if (den_i[k] != 0) {
            const double v = a[k] / den_i[k];
            a[k] = take_sqrt ? sqrt(v) : v;
        }
    }
}

// pass 2 (:296-306): squared deviation from the cell's mean line-of-sight velocity
__global__ void __launch_bounds__(kShBlock) k_sigma_dev(const SigmaParams p)
{
    const int nc = p.n_cell;
    const size_t nn = static_cast<size_t>(nc) * nc;
    double *acc = p.racc + (blockIdx.x & p.rep_mask) * (6 * nn);
    for (int64_t ip = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; ip < p.npart;
         ip += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t q = p.part_list[ip];
        const int ix = p.cell[3 * ip], iy = p.cell[3 * ip + 1], iz = p.cell[3 * ip + 2];
        const int64_t k3[3] = {iy + static_cast<int64_t>(nc) * iz, ix + static_cast<int64_t>(nc) * iz,
                               ix + static_cast<int64_t>(nc) * iy};
        const double v3[3] = {p.vx[q], p.vy[q], p.vz[q]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double d = v3[a] - p.vcm[a][k3[a]];
            atomicAdd(&acc[a * nn + k3[a]], d * d);
        }
    }
}

// pass 3 (:326-352): mean of the cell dispersions over the particles projected inside R05
__global__ void __launch_bounds__(kShBlock) k_sigma_r05(const SigmaParams p)
{
    const int nc = p.n_cell;
    double s[3] = {0, 0, 0};
    int c[3] = {0, 0, 0};
    for (int64_t ip = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; ip < p.npart;
         ip += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t q = p.part_list[ip];
        const float dx = p.x[q] - p.cx, dy = p.y[q] - p.cy, dz = p.z[q] - p.cz;
        const int ix = p.cell[3 * ip], iy = p.cell[3 * ip + 1], iz = p.cell[3 * ip + 2];
        const float dist[3] = {sqrtf(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz))),
                               sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz))),
                               sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)))};
        const int64_t k3[3] = {iy + static_cast<int64_t>(nc) * iz, ix + static_cast<int64_t>(nc) * iz,
                               ix + static_cast<int64_t>(nc) * iy};
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (dist[a] < p.r05[a]) {
                s[a] += p.sig[a][k3[a]];
                c[a] += 1;
            }
    }
    block_sum_atomic<3>(s, p.s05);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int v = c[a];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&p.c05[a], v);
    }
}

// particle_subroutines.f90:12-126: cyclic Jacobi, eigenvalues only, double precision
void jacobi_eigenvalues(const float in[3][3], float out[3])
{
    double a[3][3], d[3], b[3], z[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = in[i][j];
    double total = 0.0;
    for (int i = 0; i < 3; ++i) {
        b[i] = d[i] = a[i][i];
        z[i] = 0.0;
        for (int j = i; j < 3; ++j) total += std::fabs(a[i][j]);
    }
    const double tol = static_cast<double>(1.e-4f) * total;       // REAL*4 literal in the reference (:57)
    for (int sweep = 1; sweep <= 100; ++sweep) {
        const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
        if (off < tol) break;
        const double limit = sweep < 4 ? 0.2 * off * off : 0.0;
        for (int i = 0; i < 2; ++i)
            for (int j = i + 1; j < 3; ++j) {
                const double g = 100.0 * std::fabs(a[i][j]);
                if (sweep > 4 && std::fabs(d[i]) + g == std::fabs(d[i]) && std::fabs(d[j]) + g == std::fabs(d[j])) {
                    a[i][j] = 0.0;
                    continue;
                }
                if (!(std::fabs(a[i][j]) > limit)) continue;
                double h = d[j] - d[i], t;
                if (std::fabs(h) + g == std::fabs(h)) {
                    t = a[i][j] / h;
                } else {
                    const double theta = 0.5 * h / a[i][j];
                    t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
                    if (theta < 0.0) t = -t;
                }
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
                h = t * a[i][j];
                z[i] -= h;
                z[j] += h;
                d[i] -= h;
                d[j] += h;
                a[i][j] = 0.0;
                auto rot = [&](double &p, double &q) {
                    const double u = p, v = q;
                    p = u - s * (v + u * tau);
                    q = v + s * (u - v * tau);
                };
                for (int k = 0; k < i; ++k) rot(a[k][i], a[k][j]);
                for (int k = i + 1; k < j; ++k) rot(a[i][k], a[k][j]);
                for (int k = j + 1; k < 3; ++k) rot(a[i][k], a[j][k]);
            }
        for (int i = 0; i < 3; ++i) {
            b[i] += z[i];
            d[i] = b[i];
            z[i] = 0.0;
        }
    }
    for (int i = 0; i < 3; ++i) out[i] = static_cast<float>(d[i]);
}

// particle_subroutines.f90:129-157: largest first
void sort_largest_first(float *e)
{
    for (int i = 0; i < 2; ++i) {
        int k = i;
        float v = e[i];
        for (int j = i + 1; j < 3; ++j)
            if (e[j] >= v) {
                k = j;
                v = e[j];
            }
        if (k != i) {
            e[k] = e[i];
            e[i] = v;
        }
    }
}

}  // namespace
}  // namespace halma

using namespace halma;

extern "C" int halma_halo_shape_f32(int device, const float *x, const float *y, const float *z, const float *mass,
                                    int64_t npart, float *eig)
{
    if (!eig) return halma_internal_fail(HALMA_ERR_INVALID, "null output");
    if (npart < 0) return halma_internal_fail(HALMA_ERR_INVALID, "negative size");
    if (npart > 0 && (!x || !y || !z || !mass)) return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    int sm = 0;
    cudaStream_t s = nullptr;
    if (int rc = halma_internal_ctx(device, &sm, &s)) return rc;
    double h7[7] = {0, 0, 0, 0, 0, 0, 0};
    if (npart > 0) {
        float *d = nullptr;
        double *d7 = nullptr;
        const size_t n = (static_cast<size_t>(npart) + 3) & ~size_t(3);      // stride: 16-byte aligned arrays
        const int blocks = static_cast<int>(std::min<int64_t>((npart / 4 + kShBlock) / kShBlock, sm * 8));
        std::vector<double> part(static_cast<size_t>(blocks) * 7);
        SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d), 4 * n * sizeof(float), s));
        SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d7), part.size() * sizeof(double), s));
        const float *src[4] = {x, y, z, mass};
        for (int k = 0; k < 4; ++k)
            SH_TRY(cudaMemcpyAsync(d + k * n, src[k], npart * sizeof(float), cudaMemcpyHostToDevice, s));
        KernelTimer timer(s);
        k_inertia<<<blocks, kShBlock, 0, s>>>(d, d + n, d + 2 * n, d + 3 * n, npart, d7);
        timer.stop();
        SH_TRY(cudaGetLastError());
        SH_TRY(cudaMemcpyAsync(part.data(), d7, part.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        SH_TRY(cudaFreeAsync(d, s));
        SH_TRY(cudaFreeAsync(d7, s));
        SH_TRY(cudaStreamSynchronize(s));
        timer.publish();
        for (int b = 0; b < blocks; ++b)
            for (int k = 0; k < 7; ++k) h7[k] += part[static_cast<size_t>(b) * 7 + k];
    }
    // :203 normalise, :206-211 eigenvalues, largest first, square roots
    const double M = h7[6];
    float t[3][3];
    const double sym[3][3] = {{h7[0], h7[1], h7[2]}, {h7[1], h7[3], h7[4]}, {h7[2], h7[4], h7[5]}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[i][j] = static_cast<float>(sym[i][j] / M);
    jacobi_eigenvalues(t, eig);
    sort_largest_first(eig);
    for (int i = 0; i < 3; ++i) eig[i] = std::sqrt(eig[i]);
    return HALMA_OK;
}

extern "C" int halma_sigma_projections_f32(int device, int64_t npart, const float *grid, int32_t n_cell,
                                           const int32_t *part_list, int64_t n_all, const float *st_x,
                                           const float *st_y, const float *st_z, const float *st_vx,
                                           const float *st_vy, const float *st_vz, const float *st_mass, float cx,
                                           float cy, float cz, float R05x, float R05y, float R05z, float ll,
                                           float *out5)
{
    if (!out5 || !grid) return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    if (npart < 0 || n_all < 0 || n_cell < 1 || n_cell > 8192) return halma_internal_fail(HALMA_ERR_INVALID, "bad size");
    if (npart > 0 && (!part_list || !st_x || !st_y || !st_z || !st_vx || !st_vy || !st_vz || !st_mass))
        return halma_internal_fail(HALMA_ERR_INVALID, "null pointer");
    for (int64_t i = 0; i < npart; ++i)
        if (part_list[i] < 0 || part_list[i] >= n_all) return halma_internal_fail(HALMA_ERR_INVALID, "part_list out of range");
    int sm = 0;
    cudaStream_t s = nullptr;
    if (int rc = halma_internal_ctx(device, &sm, &s)) return rc;
    const size_t nn = static_cast<size_t>(n_cell) * n_cell, na = static_cast<size_t>(std::max<int64_t>(n_all, 1)),
                 np = static_cast<size_t>(std::max<int64_t>(npart, 1));
    float *d_f = nullptr;
    double *d_maps = nullptr;
    int32_t *d_i = nullptr;
    int16_t *d_cell = nullptr;
    SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_f), (7 * na + n_cell) * sizeof(float), s));
    SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_maps), (9 * nn + 3) * sizeof(double), s));
    SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_i), (3 * nn + 3 + np) * sizeof(int32_t), s));
    SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_cell), 3 * np * sizeof(int16_t), s));
    SH_TRY(cudaMemsetAsync(d_maps, 0, (9 * nn + 3) * sizeof(double), s));
    SH_TRY(cudaMemsetAsync(d_i, 0, (3 * nn + 3) * sizeof(int32_t), s));
    int reps = 64;                                     // power of two, at most 2^22 map cells in all
    while (reps > 1 && static_cast<size_t>(reps) * nn > (size_t(1) << 22)) reps >>= 1;
    double *d_racc = nullptr;
    int32_t *d_rcnt = nullptr;
    SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_racc), reps * 6 * nn * sizeof(double), s));
    SH_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_rcnt), reps * 3 * nn * sizeof(int32_t), s));
    SH_TRY(cudaMemsetAsync(d_racc, 0, reps * 6 * nn * sizeof(double), s));
    SH_TRY(cudaMemsetAsync(d_rcnt, 0, reps * 3 * nn * sizeof(int32_t), s));
    const float *src[7] = {st_x, st_y, st_z, st_vx, st_vy, st_vz, st_mass};
    if (n_all > 0)
        for (int k = 0; k < 7; ++k)
            SH_TRY(cudaMemcpyAsync(d_f + k * na, src[k], n_all * sizeof(float), cudaMemcpyHostToDevice, s));
    SH_TRY(cudaMemcpyAsync(d_f + 7 * na, grid, n_cell * sizeof(float), cudaMemcpyHostToDevice, s));
    if (npart > 0)
        SH_TRY(cudaMemcpyAsync(d_i + 3 * nn + 3, part_list, npart * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    SigmaParams p;
    memset(&p, 0, sizeof p);
    p.grid = d_f + 7 * na;
    p.part_list = d_i + 3 * nn + 3;
    p.x = d_f;
    p.y = d_f + na;
    p.z = d_f + 2 * na;
    p.vx = d_f + 3 * na;
    p.vy = d_f + 4 * na;
    p.vz = d_f + 5 * na;
    p.m = d_f + 6 * na;
    p.npart = npart;
    p.n_cell = n_cell;
    p.cx = cx;
    p.cy = cy;
    p.cz = cz;
    p.r05[0] = R05x;
    p.r05[1] = R05y;
    p.r05[2] = R05z;
    for (int a = 0; a < 3; ++a) {
        p.vcm[a] = d_maps + a * nn;
        p.sd[a] = d_maps + (3 + a) * nn;
        p.sig[a] = d_maps + (6 + a) * nn;
        p.cnt[a] = d_i + a * nn;
    }
    p.s05 = d_maps + 9 * nn;
    p.c05 = d_i + 3 * nn;
    p.racc = d_racc;
    p.rcnt = d_rcnt;
    p.rep_mask = reps - 1;
    p.grid_sorted = 1;
    for (int k = 1; k < n_cell; ++k)
        if (!(grid[k] > grid[k - 1])) p.grid_sorted = 0;
    if (!(grid[0] == grid[0])) p.grid_sorted = 0;
    p.cell = d_cell;
    const int pb = static_cast<int>(std::min<int64_t>((npart + kShBlock - 1) / kShBlock, sm * 8));
    const int mb = static_cast<int>(std::min<size_t>((nn + 255) / 256, static_cast<size_t>(sm) * 8));
    KernelTimer timer(s);
    if (npart > 0) {
        k_sigma_bin<<<pb, kShBlock, n_cell * sizeof(float), s>>>(p);
        k_fold<double><<<mb, 256, 0, s>>>(d_maps, d_racc, 6 * nn, reps, 6 * nn);          // vcm | sd
        k_fold<int32_t><<<mb, 256, 0, s>>>(d_i, d_rcnt, 3 * nn, reps, 3 * nn);
        for (int a = 0; a < 3; ++a) k_sigma_maps<<<mb, 256, 0, s>>>(p.vcm[a], p.sd[a], nullptr, nn, 0);
        SH_TRY(cudaMemsetAsync(d_racc, 0, reps * 6 * nn * sizeof(double), s));
        k_sigma_dev<<<pb, kShBlock, 0, s>>>(p);
        k_fold<double><<<mb, 256, 0, s>>>(d_maps + 6 * nn, d_racc, 3 * nn, reps, 6 * nn);   // sig
        for (int a = 0; a < 3; ++a) k_sigma_maps<<<mb, 256, 0, s>>>(p.sig[a], nullptr, p.cnt[a], nn, 1);
        k_sigma_r05<<<pb, kShBlock, 0, s>>>(p);
        SH_TRY(cudaGetLastError());
    }
    timer.stop();
    std::vector<double> maps(9 * nn + 3);
    int32_t c05[3];
    SH_TRY(cudaMemcpyAsync(maps.data(), d_maps, maps.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    SH_TRY(cudaMemcpyAsync(c05, d_i + 3 * nn, sizeof c05, cudaMemcpyDeviceToHost, s));
    SH_TRY(cudaFreeAsync(d_f, s));
    SH_TRY(cudaFreeAsync(d_maps, s));
    SH_TRY(cudaFreeAsync(d_i, s));
    SH_TRY(cudaFreeAsync(d_racc, s));
    SH_TRY(cudaFreeAsync(d_rcnt, s));
    SH_TRY(cudaFreeAsync(d_cell, s));
    SH_TRY(cudaStreamSynchronize(s));
    timer.publish();
    // :356-366 means inside R05; :370-454 V/sigma and lambda_R per projection, then the average.
    // The map loops are n_cell^2 long: done here, in the reference's loop order.
    const double *vcm[3] = {maps.data(), maps.data() + nn, maps.data() + 2 * nn};
    const double *sd[3] = {maps.data() + 3 * nn, maps.data() + 4 * nn, maps.data() + 5 * nn};
    const double *sig[3] = {maps.data() + 6 * nn, maps.data() + 7 * nn, maps.data() + 8 * nn};
    const double *s05 = maps.data() + 9 * nn;
    const float r05[3] = {R05x, R05y, R05z};
    double vs[3] = {0, 0, 0}, lam[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a) {
        out5[a] = static_cast<float>(c05[a] > 0 ? s05[a] / c05[a] : s05[a]);
        double sumV = 0, sumS = 0, up = 0, down = 0;
        for (int j = 0; j < n_cell; ++j)
            for (int i = 0; i < n_cell; ++i) {
                const float rbin = std::sqrt(grid[i] * grid[i] + grid[j] * grid[j]);
                if (rbin < r05[a] + 2 * ll) {
                    const size_t k = i + static_cast<size_t>(n_cell) * j;
                    sumV += vcm[a][k] * vcm[a][k] * sd[a][k];
                    sumS += sig[a][k] * sig[a][k] * sd[a][k];
                    up += sd[a][k] * rbin * std::fabs(vcm[a][k]);
                    down += sd[a][k] * rbin * std::sqrt(vcm[a][k] * vcm[a][k] + sig[a][k] * sig[a][k]);
                }
            }
        if (sumS > 0) vs[a] = std::sqrt(sumV / sumS);
        if (down > 0) lam[a] = up / down;
    }
    out5[3] = static_cast<float>((vs[0] + vs[1] + vs[2]) / 3);
    out5[4] = static_cast<float>((lam[0] + lam[1] + lam[2]) / 3);
    return HALMA_OK;
}
