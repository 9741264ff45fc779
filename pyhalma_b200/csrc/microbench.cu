// Pipe-rate microbenchmark for the roofline denominator: how many
// MUFU.RSQ, FFMA and packed FFMA2 instructions one SM retires per clock, and the SM clock
// while doing so.  Every kernel runs one persistent block set per SM and times itself with
// clock64(), so the result is in instructions per SM-clock and does not depend on DVFS.
#include <algorithm>
#include <cstring>
#include <vector>

#include "halma_common.cuh"
#include "../../include/halma_unbind.h"

namespace halma {

constexpr int kMbThreads = 512;
constexpr int kMbIters = 4096;
constexpr int kMbIlp = 8;

template <int OP>
__global__ void __launch_bounds__(kMbThreads) k_pipe(float seed, long long *cycles, float *sink)
{
    float v[kMbIlp];
    uint64_t w[kMbIlp];
#pragma unroll
    for (int k = 0; k < kMbIlp; ++k) {
        v[k] = seed + 0.001f * (threadIdx.x + k);
        w[k] = pack2(v[k], v[k] + 1.f);
    }
    const uint64_t c2 = pack2(1.0001f, 0.9999f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
        for (int k = 0; k < kMbIlp; ++k) {
            if (OP == 0) v[k] = rsqrt_ftz(v[k]);
            if (OP == 1) v[k] = fmaf(v[k], 1.0001f, 0.5f);
            if (OP == 2) w[k] = fma2(w[k], c2, c2);
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kMbIlp; ++k) {
        float a, b;
        unpack2(w[k], a, b);
        acc += v[k] + a + b;
    }
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
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
    const int blocks = sm * (2048 / kMbThreads);      // fill every SM with resident threads
    long long *d_cyc = nullptr;
    float *d_sink = nullptr;
    if (cudaMalloc(&d_cyc, blocks * sizeof(long long)) != cudaSuccess) return HALMA_ERR_CUDA;
    if (cudaMalloc(&d_sink, 16) != cudaSuccess) return HALMA_ERR_CUDA;
    std::vector<long long> cyc(blocks);
    double rate[3] = {0, 0, 0}, mhz = 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k_pipe<0><<<blocks, kMbThreads>>>(1.5f, d_cyc, d_sink);
            if (op == 1) k_pipe<1><<<blocks, kMbThreads>>>(1.5f, d_cyc, d_sink);
            if (op == 2) k_pipe<2><<<blocks, kMbThreads>>>(1.5f, d_cyc, d_sink);
            cudaEventRecord(e1);
            if (cudaDeviceSynchronize() != cudaSuccess) return HALMA_ERR_CUDA;
        }
        cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        std::sort(cyc.begin(), cyc.end());
        const double med = static_cast<double>(cyc[blocks / 2]);
        // resident threads per SM * instructions per thread / cycles the SM took
        const double per_thread = static_cast<double>(kMbIters) * kMbIlp;
        rate[op] = 2048.0 * per_thread / med;
        if (op == 0) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            mhz = med / (ms * 1e-3) / 1e6;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_cyc);
    cudaFree(d_sink);
    memset(out8, 0, 8 * sizeof(double));
    out8[0] = rate[0];
    out8[1] = rate[1];
    out8[2] = rate[2];
    out8[3] = mhz;
    out8[4] = sm;
    return HALMA_OK;
}
