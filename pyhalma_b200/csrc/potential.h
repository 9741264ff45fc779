// Host-visible interface of potential.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/halma_unbind.h"
#include "halma_common.cuh"

namespace halma {

constexpr int kPotentialBlock = 128;     // 4 warps; each warp schedules itself

// FAST kernel shapes (potential.cu): index 0 = throughput shape (128 targets per ticket),
// 1 = small-halo shape (32 targets per ticket).
constexpr int kMaxVariants = 8;
int potential_num_variants();
// groups_of_128: sum over haloes of ceil(n_targets / 128); max_sources: largest source count.
int potential_pick_variant(int64_t groups_of_128, int64_t max_sources, int resident_warps, int64_t max_members,
                           bool symmetric, int max_split, int min_split_sources);
// Targets per work item (warp): 32 (EXACT) or 32 * T (FAST, T = targets per lane of the shape).
int potential_group_size(int mode, int variant);
// Sets the dynamic shared-memory attribute and returns resident blocks per SM.
cudaError_t potential_configure(int mode, int variant, int *blocks_per_sm);
cudaError_t potential_launch(const PotParams &p, int mode, int variant, int grid_blocks, cudaStream_t stream);
// Device self-test of the EXACT kernel's branch-free sqrt / divide against the IEEE library
// routines; *d_mismatch (zeroed by the caller) receives the number of differing results.
cudaError_t potential_selftest_exact(int64_t n_random, uint64_t seed, unsigned long long *d_mismatch, int sm_count,
                                     cudaStream_t stream);

}  // namespace halma
