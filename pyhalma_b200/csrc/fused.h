// Host-visible interface of fused.cu: the whole unbinding loop as one persistent cooperative kernel.
#pragma once
#include <cuda_runtime.h>

#include "halma_common.cuh"
#include "loop_kernels.h"

namespace halma {

// Which persistent kernel serves a plan, or -1 (tuning shapes and split mode use the multi-launch drivers).
int fused_kernel_index(int mode, int variant, bool np, int sym_rows);      // sym_rows: 0, 4 or 8
// Shared-memory attribute and resident blocks per SM of every persistent kernel (once per device).
cudaError_t fused_configure();
int fused_blocks_per_sm(int index);
// One cooperative launch = the complete loop (+ the pack when do_pack != 0).
cudaError_t fused_launch(int index, const PotParams &pp, const LoopParams &lp, int do_pack, int sm_count,
                         cudaStream_t stream);

}  // namespace halma
