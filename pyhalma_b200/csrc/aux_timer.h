// CUDA-event bracket around the kernels of an auxiliary entry point (shape.cu, gather.cu);
// the elapsed time is published through halma_last_kernel_ms().
#pragma once
#include <cuda_runtime.h>

void halma_internal_set_kernel_ms(double ms);

namespace halma {
struct KernelTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s = nullptr;
    explicit KernelTimer(cudaStream_t stream) : s(stream)
    {
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) a = b = nullptr;
        if (a) cudaEventRecord(a, s);
    }
    void stop()
    {
        if (a) cudaEventRecord(b, s);
    }
    // call after the stream was synchronised
    void publish()
    {
        float ms = 0.f;
        if (a && cudaEventElapsedTime(&ms, a, b) == cudaSuccess) halma_internal_set_kernel_ms(ms);
    }
    ~KernelTimer()
    {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};
}  // namespace halma
