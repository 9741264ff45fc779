// Host-visible interface of potential.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/halma_unbind.h"
#include "halma_common.cuh"

namespace halma {

constexpr int kPotentialBlock = 128;     // 4 warps; each warp schedules itself

// Targets per work item (warp): 32 (EXACT) or 32 * T (FAST, T = targets per lane of the variant).
int potential_group_size(int mode);
// Sets the dynamic shared-memory attribute and returns resident blocks per SM.
cudaError_t potential_configure(int mode, int *blocks_per_sm);
cudaError_t potential_launch(const PotParams &p, int mode, int grid_blocks, cudaStream_t stream);

}  // namespace halma
