// Pipe-rate microbenchmark for the roofline denominator: how many
// MUFU.RSQ, FFMA and packed FFMA2 instructions the device retires per second, and the SM
// clock while doing so (clock64 ticks against %globaltimer inside the kernel), from which
// the per-SM-per-clock rates follow.  Rates come from CUDA-event time of a single resident
// wave, so they do not depend on assumed occupancy.  See csrc/pipebench.cu for the full set
// of instruction forms.
#include <algorithm>
#include <cstring>
#include <vector>

#include "halma_common.cuh"
#include "../../include/halma_unbind.h"

namespace halma {

constexpr int kMbThreads = 256;
constexpr int kMbIlp = 8;
constexpr int kMbUnroll = 8;

__device__ __forceinline__ unsigned long long mb_gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int OP>
__global__ void __launch_bounds__(kMbThreads) k_pipe(int iters, float seed, long long *ticks, unsigned long long *ns,
                                                    float *sink)
{
    float v[kMbIlp];
    uint64_t w[kMbIlp];
    const float b = 1.0f + seed * 1e-3f, c = 0.25f + seed;
#pragma unroll
    for (int k = 0; k < kMbIlp; ++k) {
        v[k] = seed + 0.001f * (threadIdx.x + k);
        w[k] = pack2(v[k], v[k] + 1.f);
    }
    const uint64_t wb = pack2(b, b), wc = pack2(c, 2 * c);
    __syncthreads();
    const unsigned long long g0 = mb_gtimer();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < kMbUnroll; ++r) {
#pragma unroll
            for (int k = 0; k < kMbIlp; ++k) {
                if (OP == 0) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
                if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[k]) : "f"(b), "f"(c));
                if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[k]) : "l"(wb), "l"(wc));
            }
        }
    }
    const long long t1 = clock64();
    const unsigned long long g1 = mb_gtimer();
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kMbIlp; ++k) {
        float lo, hi;
        unpack2(w[k], lo, hi);
        acc += v[k] + lo + hi;
    }
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) {
        ticks[blockIdx.x] = t1 - t0;
        ns[blockIdx.x] = g1 - g0;
    }
}

template <int OP>
static int run_pipe(int sm, int iters, long long *d_ticks, unsigned long long *d_ns, float *d_sink, int max_blocks,
                    double *gops, double *mhz)
{
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_pipe<OP>, kMbThreads, 0) != cudaSuccess || bps < 1)
        return HALMA_ERR_CUDA;
    const int blocks = std::min(sm * bps, max_blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_pipe<OP><<<blocks, kMbThreads>>>(iters, 1.5f, d_ticks, d_ns, d_sink);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) return HALMA_ERR_CUDA;
        cudaEventElapsedTime(&ms, e0, e1);
    }
    std::vector<long long> t(blocks);
    std::vector<unsigned long long> n(blocks);
    cudaMemcpy(t.data(), d_ticks, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(n.data(), d_ns, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    std::sort(t.begin(), t.end());
    std::sort(n.begin(), n.end());
    *mhz = static_cast<double>(t[blocks / 2]) / static_cast<double>(n[blocks / 2]) * 1e3;
    const double total = static_cast<double>(blocks) * kMbThreads * iters * kMbUnroll * kMbIlp;
    *gops = total / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HALMA_OK;
}

}  // namespace halma

extern "C" int halma_microbench(int device, double *out8)
{
    using namespace halma;
    if (!out8) return HALMA_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return HALMA_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HALMA_ERR_CUDA;
    const int sm = prop.multiProcessorCount;
    const int max_blocks = sm * 16;
    long long *d_ticks = nullptr;
    unsigned long long *d_ns = nullptr;
    float *d_sink = nullptr;
    if (cudaMalloc(&d_ticks, max_blocks * 8) != cudaSuccess || cudaMalloc(&d_ns, max_blocks * 8) != cudaSuccess ||
        cudaMalloc(&d_sink, 16) != cudaSuccess)
        return HALMA_ERR_CUDA;
    double g[3] = {0, 0, 0}, mhz[3] = {0, 0, 0};
    int rc = run_pipe<0>(sm, 4000, d_ticks, d_ns, d_sink, max_blocks, &g[0], &mhz[0]);
    if (!rc) rc = run_pipe<1>(sm, 20000, d_ticks, d_ns, d_sink, max_blocks, &g[1], &mhz[1]);
    if (!rc) rc = run_pipe<2>(sm, 10000, d_ticks, d_ns, d_sink, max_blocks, &g[2], &mhz[2]);
    cudaFree(d_ticks);
    cudaFree(d_ns);
    cudaFree(d_sink);
    if (rc) return rc;
    memset(out8, 0, 8 * sizeof(double));
    for (int k = 0; k < 3; ++k) out8[k] = g[k] * 1e9 / (static_cast<double>(sm) * mhz[k] * 1e6);   // per clk per SM
    out8[3] = mhz[0];
    out8[4] = sm;
    out8[5] = g[0];      // absolute G instr/s
    out8[6] = g[1];
    out8[7] = g[2];
    return HALMA_OK;
}
