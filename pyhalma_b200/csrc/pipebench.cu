// pipebench: per-instruction issue throughput on sm_100a, for the roofline denominator and
// for choosing the instruction mix of the potential kernel.  Standalone tool:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipebench pipebench.cu && ./pipebench
// One resident wave (grid = SMs x occupancy), every block times itself with clock64() and
// %globaltimer, so the SM clock during the test is measured, not assumed.  Output:
// thread-instructions per SM clock per SM for each instruction form, JSON lines.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int kThreads = 256;
constexpr int kIlp = 8;
constexpr int kUnroll = 8;       // kIlp * kUnroll instructions per loop trip

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t gtimer() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

enum Op { RSQ, FFMA_REG, FFMA_IMM, FADD_REG, FADD2_REG, FADD2_BCAST, FMUL2, FFMA2, FMNMX3, FSETP_SEL, BODY_PACKED, BODY_SCALAR,
          MIX_FADD2_FMNMX3, MIX_FADD2_RSQ, MIX_FADD2_FFMA, MIX_FFMA_FMNMX3, MIX_FFMA_RSQ, MIX_FADD_FADD2, FSETP_ONLY, BODY_NOPRED, N_OPS };
static const char *kNames[N_OPS] = {"mufu_rsq", "ffma_reg", "ffma_imm", "fadd_reg", "fadd2_reg", "fadd2_bcast", "fmul2",
                                    "ffma2", "fmnmx3", "fsetp_ffma_pred", "body_packed(2 inter)", "body_scalar(1 inter)",
                                    "mix fadd2+fmnmx3", "mix fadd2+rsq", "mix fadd2+ffma", "mix ffma+fmnmx3", "mix ffma+rsq",
                                    "mix fadd+fadd2", "fsetp+sel", "body_nopred(2 inter)"};
// thread-instructions issued per "op" (for BODY_*: issue slots of one body)
static const int kSlots[N_OPS] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 14, 11, 2, 2, 2, 2, 2, 2, 2, 9};
static const int kInter[N_OPS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 2};

template <int OP>
__global__ void __launch_bounds__(kThreads) k(int iters, float seed, float4 *out, long long *ticks, unsigned long long *ns)
{
    float v[kIlp], u[kIlp];
    uint64_t w[kIlp];
    const float a = seed + threadIdx.x * 1e-3f, b = 1.0f + seed * 1e-3f, c = 0.25f + seed;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) { v[k] = a + k; u[k] = c + k; w[k] = pk(a + k, a - k); }
    const uint64_t wb = pk(b, b), wc = pk(c, 2 * c);
    __syncthreads();
    const unsigned long long g0 = gtimer();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < kUnroll; ++r) {
#pragma unroll
            for (int k = 0; k < kIlp; ++k) {
                if (OP == RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
                if (OP == FFMA_REG) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[k]) : "f"(b), "f"(c));
                if (OP == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, 0f3F800054, %1;" : "+f"(v[k]) : "f"(c));
                if (OP == FADD_REG) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[k]) : "f"(c));
                if (OP == FADD2_REG) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[k]) : "l"(wc));
                if (OP == FADD2_BCAST) asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %0, t;}" : "+l"(w[k]) : "f"(c));
                if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[k]) : "l"(wb));
                if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[k]) : "l"(wb), "l"(wc));
                if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(v[k]) : "f"(b), "f"(u[k]));
                if (OP == FSETP_SEL)
                    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; @p fma.rn.f32 %0, %0, %2, %3;}"
                                 : "+f"(v[k]) : "f"(u[k]), "f"(b), "f"(c));
                if (OP == BODY_PACKED) {
                    // two interactions: the hot-loop body of k_potential_fast, registers only
                    uint64_t dx, dy, dz, r2;
                    float dx0, dx1, dy0, dy1, dz0, dz1, r0, r1, t0_, t1_, i0, i1;
                    asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %2, t;}" : "=l"(dx) : "f"(a), "l"(w[k]));
                    asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %2, t;}" : "=l"(dy) : "f"(b), "l"(w[(k + 1) % kIlp]));
                    asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %2, t;}" : "=l"(dz) : "f"(c), "l"(w[(k + 2) % kIlp]));
                    asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(r2) : "l"(dy));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(r2) : "l"(dx));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(r2) : "l"(dz));
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(dx0), "=f"(dx1) : "l"(dx));
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(dy0), "=f"(dy1) : "l"(dy));
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(dz0), "=f"(dz1) : "l"(dz));
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r2));
                    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(t0_) : "f"(fabsf(dx0)), "f"(fabsf(dy0)), "f"(fabsf(dz0)));
                    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(t1_) : "f"(fabsf(dx1)), "f"(fabsf(dy1)), "f"(fabsf(dz1)));
                    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(r0));
                    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(r1));
                    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; @p fma.rn.f32 %0, %2, %3, %0;}" : "+f"(v[k]) : "f"(t0_), "f"(u[k]), "f"(i0));
                    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; @p fma.rn.f32 %0, %2, %3, %0;}" : "+f"(v[k]) : "f"(t1_), "f"(u[k]), "f"(i1));
                    w[k] = dx;      // loop-carried: nothing is invariant
                }
                if (OP == BODY_NOPRED) {
                    // two interactions without the exclusion predicate and with a packed accumulate
                    uint64_t dx, dy, dz, r2, inv;
                    float r0, r1, i0, i1;
                    asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %2, t;}" : "=l"(dx) : "f"(a), "l"(w[k]));
                    asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %2, t;}" : "=l"(dy) : "f"(b), "l"(w[(k + 1) % kIlp]));
                    asm volatile("{.reg .b64 t; mov.b64 t,{%1,%1}; sub.rn.f32x2 %0, %2, t;}" : "=l"(dz) : "f"(c), "l"(w[(k + 2) % kIlp]));
                    asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(r2) : "l"(dy));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(r2) : "l"(dx));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(r2) : "l"(dz));
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r2));
                    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(r0));
                    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(r1));
                    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(inv) : "f"(i0), "f"(i1));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(w[(k + 3) % kIlp]) : "l"(wb), "l"(inv));
                    w[k] = dx;
                }
                if (OP == MIX_FADD2_FMNMX3) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[k]) : "l"(wc)); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(v[k]) : "f"(b), "f"(u[k])); }
                if (OP == MIX_FADD2_RSQ) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[k]) : "l"(wc)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k])); }
                if (OP == MIX_FADD2_FFMA) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[k]) : "l"(wc)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[k]) : "f"(b), "f"(c)); }
                if (OP == MIX_FFMA_FMNMX3) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(u[k]) : "f"(b), "f"(c)); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(v[k]) : "f"(b), "f"(c)); }
                if (OP == MIX_FFMA_RSQ) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(u[k]) : "f"(b), "f"(c)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k])); }
                if (OP == MIX_FADD_FADD2) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[k]) : "f"(c)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[k]) : "l"(wc)); }
                if (OP == FSETP_ONLY) { asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(v[k]) : "f"(u[k]), "f"(b)); }
                if (OP == BODY_SCALAR) {
                    float dx, dy, dz, r2, t_, i_;
                    float lo, hi;
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(w[k]));
                    asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dx) : "f"(lo), "f"(a));
                    asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dy) : "f"(hi), "f"(b));
                    asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dz) : "f"(u[k]), "f"(c));
                    asm volatile("mul.rn.f32 %0, %1, %1;" : "=f"(r2) : "f"(dy));
                    asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(r2) : "f"(dx));
                    asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(r2) : "f"(dz));
                    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(t_) : "f"(fabsf(dx)), "f"(fabsf(dy)), "f"(fabsf(dz)));
                    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i_) : "f"(r2));
                    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; @p fma.rn.f32 %0, %2, %3, %0;}" : "+f"(v[k]) : "f"(t_), "f"(b), "f"(i_));
                    w[k] = pk(dx, dy);
                    u[k] = dz;
                }
            }
        }
    }
    const long long t1 = clock64();
    const unsigned long long g1 = gtimer();
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
        float lo, hi;
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(w[k]));
        s += v[k] + u[k];
        s2 += lo + hi;
    }
    if (s == 1234.5f) out[0] = make_float4(s, s2, 0, 0);
    if (threadIdx.x == 0) {
        ticks[blockIdx.x] = t1 - t0;
        ns[blockIdx.x] = g1 - g0;
    }
}

template <int OP>
static void run(int sm, int iters)
{
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k<OP>, kThreads, 0);
    const int blocks = sm * bps;
    float4 *out; long long *ticks; unsigned long long *ns;
    cudaMalloc(&out, 64); cudaMalloc(&ticks, blocks * 8); cudaMalloc(&ns, blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<OP><<<blocks, kThreads>>>(iters, 1.5f, out, ticks, ns);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
    }
    std::vector<long long> t(blocks); std::vector<unsigned long long> n(blocks);
    cudaMemcpy(t.data(), ticks, blocks * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(n.data(), ns, blocks * 8, cudaMemcpyDeviceToHost);
    std::sort(t.begin(), t.end()); std::sort(n.begin(), n.end());
    const double med_t = (double)t[blocks / 2], med_ns = (double)n[blocks / 2];
    const double mhz = med_t / med_ns * 1e3;
    const double ops_per_thread = (double)iters * kUnroll * kIlp;
    const double thr_per_sm = (double)bps * kThreads;
    const double op_per_clk_sm = thr_per_sm * ops_per_thread / med_t;          // "ops" (bodies) per clk per SM
    const double total_ops = (double)blocks * kThreads * ops_per_thread;
    printf("{\"op\": \"%s\", \"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"sm_mhz\": %.1f, \"event_ms\": %.3f, "
           "\"ops_per_clk_sm\": %.2f, \"issue_slots_per_clk_sm\": %.2f, \"interactions_per_clk_sm\": %.2f, "
           "\"Gops_per_s\": %.1f}\n",
           kNames[OP], bps, bps * kThreads / 32, mhz, ms, op_per_clk_sm, op_per_clk_sm * kSlots[OP],
           op_per_clk_sm * kInter[OP], total_ops / (ms * 1e-3) / 1e9);
    fflush(stdout);
    cudaFree(out); cudaFree(ticks); cudaFree(ns);
}

int main(int argc, char **argv)
{
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    printf("{\"device\": \"%s\", \"sm_count\": %d}\n", p.name, p.multiProcessorCount);
    const int sm = p.multiProcessorCount;
    run<RSQ>(sm, iters); run<FFMA_REG>(sm, iters); run<FFMA_IMM>(sm, iters); run<FADD_REG>(sm, iters);
    run<FADD2_REG>(sm, iters); run<FADD2_BCAST>(sm, iters); run<FMUL2>(sm, iters); run<FFMA2>(sm, iters);
    run<FMNMX3>(sm, iters); run<FSETP_SEL>(sm, iters); run<FSETP_ONLY>(sm, iters);
    run<MIX_FADD2_FMNMX3>(sm, iters); run<MIX_FADD2_RSQ>(sm, iters); run<MIX_FADD2_FFMA>(sm, iters); run<MIX_FFMA_FMNMX3>(sm, iters);
    run<MIX_FFMA_RSQ>(sm, iters); run<MIX_FADD_FADD2>(sm, iters);
    run<BODY_PACKED>(sm, iters / 4); run<BODY_SCALAR>(sm, iters / 4); run<BODY_NOPRED>(sm, iters / 4);
    return 0;
}
