// Probe: throughput of float64 weighted-histogram updates on B200 for the sigma_projections maps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ahp scripts/probes/atomic_hist_probe.cu && /tmp/ahp
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_global(const int *key, const double *w, int n, double *maps, int nb, int reps_mask)
{
    double *m = maps + (size_t)(blockIdx.x & reps_mask) * nb;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&m[key[i]], w[i]);
}

__global__ void k_match(const int *key, const double *w, int n, double *maps, int nb)
{
    const int lane = threadIdx.x & 31;
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        const bool ok = i < n;
        const int k = ok ? key[i] : -1 - lane;
        double v = ok ? w[i] : 0.0;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1;
        // sum over peers: every lane walks the peer mask
        double s = 0.0;
        unsigned mk = peers;
        while (mk) {
            const int src = __ffs(mk) - 1;
            mk &= mk - 1;
            s += __shfl_sync(peers, v, src);
        }
        if (ok && lane == leader) atomicAdd(&maps[k], s);
    }
}

__global__ void k_smem(const int *key, const double *w, int n, double *maps, int nb)
{
    extern __shared__ double s[];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&s[key[i]], w[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (s[i] != 0.0) atomicAdd(&maps[i], s[i]);
}

// fixed point in shared memory with native 64-bit integer atomics
__global__ void k_smem_fixed(const int *key, const double *w, int n, double *maps, int nb, double scale)
{
    extern __shared__ unsigned long long si[];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) si[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&si[key[i]], (unsigned long long)__double2ll_rn(w[i] * scale));
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (si[i]) atomicAdd(&maps[i], (double)(long long)si[i] / scale);
}

int main()
{
    const int n = 10000000;
    std::vector<int> key(n);
    std::vector<double> w(n);
    int dev_sm = 0;
    CK(cudaDeviceGetAttribute(&dev_sm, cudaDevAttrMultiProcessorCount, 0));
    int *d_key; double *d_w, *d_maps;
    CK(cudaMalloc(&d_key, n * sizeof(int)));
    CK(cudaMalloc(&d_w, n * sizeof(double)));
    CK(cudaMalloc(&d_maps, 64ull * 128 * 128 * sizeof(double)));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int nc : {25, 57, 101}) {
        const int nb = nc * nc;
        srand(1);
        for (int i = 0; i < n; ++i) {
            auto g = [&]() { double u = 0; for (int k = 0; k < 6; ++k) u += rand() / (double)RAND_MAX; return (u - 3.0) / 1.4; };
            int ix = (int)lrint(g() * nc / 6.0 + nc / 2), iy = (int)lrint(g() * nc / 6.0 + nc / 2);
            ix = ix < 0 ? 0 : ix >= nc ? nc - 1 : ix;
            iy = iy < 0 ? 0 : iy >= nc ? nc - 1 : iy;
            key[i] = ix + nc * iy;
            w[i] = 1.0 + (i % 7);
        }
        CK(cudaMemcpy(d_key, key.data(), n * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_w, w.data(), n * sizeof(double), cudaMemcpyHostToDevice));
        auto run = [&](const char *name, auto launch) {
            float best = 1e9;
            double sum = 0;
            for (int r = 0; r < 4; ++r) {
                CK(cudaMemset(d_maps, 0, 64ull * nb * sizeof(double)));
                cudaEventRecord(a);
                launch();
                cudaEventRecord(b);
                CK(cudaEventSynchronize(b));
                CK(cudaGetLastError());
                float ms; cudaEventElapsedTime(&ms, a, b);
                best = ms < best ? ms : best;
            }
            std::vector<double> h(64ull * nb);
            CK(cudaMemcpy(h.data(), d_maps, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (double v : h) sum += v;
            printf("n_cell=%3d %-22s %8.3f ms  %7.1f G upd/s  sum=%.6e\n", nc, name, best, n / best / 1e6, sum);
        };
        const int blocks = dev_sm * 8;
        run("global R=1", [&] { k_global<<<blocks, 256>>>(d_key, d_w, n, d_maps, nb, 0); });
        run("global R=8", [&] { k_global<<<blocks, 256>>>(d_key, d_w, n, d_maps, nb, 7); });
        run("global R=64", [&] { k_global<<<blocks, 256>>>(d_key, d_w, n, d_maps, nb, 63); });
        run("match_any + global", [&] { k_match<<<blocks, 256>>>(d_key, d_w, n, d_maps, nb); });
        if (nb * 8 <= 200 * 1024) {
            CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, nb * 8));
            CK(cudaFuncSetAttribute(k_smem_fixed, cudaFuncAttributeMaxDynamicSharedMemorySize, nb * 8));
            for (int bps : {1, 2, 4}) {
                if ((size_t)nb * 8 * bps > 220 * 1024) continue;
                char nm[64];
                snprintf(nm, sizeof nm, "smem CAS f64 x%d", bps);
                run(nm, [&] { k_smem<<<dev_sm * bps, 512, nb * 8>>>(d_key, d_w, n, d_maps, nb); });
                snprintf(nm, sizeof nm, "smem fixed u64 x%d", bps);
                run(nm, [&] { k_smem_fixed<<<dev_sm * bps, 512, nb * 8>>>(d_key, d_w, n, d_maps, nb, 1048576.0); });
            }
        }
    }
    return 0;
}
