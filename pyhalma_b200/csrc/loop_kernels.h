// Host-visible interface of loop_kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/halma_unbind.h"
#include "halma_common.cuh"

namespace halma {

constexpr int kChunk = 256;               // members per bookkeeping chunk (one warp each) ...
constexpr int kChunkSmall = 64;           // ... and in plans with few members, where latency counts, not bandwidth
constexpr int kChunkSmallMaxMembers = 1 << 18;
constexpr int kChunkSums = 10;            // float64 sums kept per chunk
constexpr int kMinSplitSources = 2048;    // never split a halo's sources into pieces below this
constexpr int kMaxSplit = 8;              // planes of the partial-potential buffer
constexpr int kMinSplitSourcesSmall = 512;    // ... both in small plans (<= kChunkSmallMaxMembers members)
constexpr int kMaxSplitSmall = 32;
constexpr int kNominalTickets = 32768;    // j-split aims at this many tickets per pass (machine-independent)

struct LoopParams {
    // static description
    const HaloDesc *halo;
    const int32_t *chunk_halo, *chunk_p0, *order;
    int32_t n_halo, n_chunks;
    int32_t chunk;                        // members per chunk: kChunk or kChunkSmall
    int64_t n_pad, n_user;
    // pristine float64 members in the user's layout
    const double *x64, *y64, *z64, *vx, *vy, *vz, *m64;
    // float32 working sets (ping-pong) + global user index of each slot
    float *wx[2], *wy[2], *wz[2], *wm[2];
    int32_t *widx[2];
    // per-halo dynamic state
    int32_t *cnt, *cnt_next, *iter, *active, *active_next, *halo_buf, *item_base, *converged;
    int32_t *nsplit;                      // [n_halo] j-splits of the coming pass, in `order` space like item_base
    const int32_t *rank_of;               // [n_halo] position of a halo in `order`
    int4 *sched;                          // [n_halo] scheduling record per halo, in `order` space (decide_halo)
    int32_t *halo_done;                   // [n_halo] chunks of the halo finished by the energy step of this pass
    int32_t *halo_stamp;                  // [n_halo] pass + 1 once the halo's decision of that pass is published
    int32_t *halo_rmin, *halo_rmax;       // [3 n_halo] min / max of x, y, z as order-preserving ints (pack);
                                          // set to 0x7f7f7f7f / 0x80808080 before the pack
    double *hM, *hvb, *hvb_next, *hcom;
    unsigned long long *pairs;
    unsigned long long *evals;            // per halo: 1/r evaluations actually made (< pairs in symmetric mode)
    // per-chunk scratch
    int32_t *chunk_cnt, *chunk_off;
    double *chunk_sum;                    // kChunkSums per chunk: m, m*v[3], m*x[3] of the survivors,
                                          // cold survivors' m, removed cold m, removed hot m
    float *chunk_best;                    // per chunk: largest potential of the pass ...
    int32_t *chunk_best_q;                // ... and the member position that has it
    const double *temp;                   // optional member temperatures (user layout)
    double cold_T;
    double *hrps;                         // 4 per halo: initial M, cold bound, unbound cold, unbound hot
    int32_t *hbest;                       // per halo: user-local index of the most bound member
    // per-particle
    uint8_t *flag;                        // [n_pad] bound flag of the current pass
    uint8_t *out_mask;                    // [n_user]
    float *out_be;                        // [n_user]
    double *out_E;                        // [n_user]
    int32_t *out_idx;                     // [n_user]
    double *phi_part;                     // [kMaxSplit][n_pad]
    LoopState *st;
    float G32, kappa32;
    int32_t vb_fixed, max_iter, mode, group_size, rank, n_ranks, target_items, max_split, min_split_sources;
    // predicate-free FAST path (potential.cu): sorted copies + per-halo fallback flags
    SortedAxis ax[3];
    const float *ax_m0[3];                // pristine masses of the sorted copies
    int64_t n_spad;                       // length of one sorted copy
    int32_t *halo_redo;
    int32_t np_enabled;
    int32_t redo_enabled;                 // some pass may hand a halo to the predicated re-evaluation (halo_redo)
    // symmetric self-term (potential.cu::sym_ticket): off-diagonal member x member sums
    double *phi_sym;                      // [n_pad]; the energy step clears what it read
    int32_t sym_enabled;
    int32_t sym_rows;                     // row members per lane of the symmetric tickets (4 or 8), as in PotParams
    double *sym_ext;                      // [n_halo] largest coordinate extent of the halo's members
    double *sym_q;                        // [n_halo] quantum of the symmetric sums for the coming pass
    // external-sum cache (potential.cu, main tickets): first-pass sums over the external sources
    double *phi_ext;                      // [max_split][n_pad] by original slot; plane 0 holds the folded sum
    int32_t *ext_ok;                      // [n_halo] 1: phi_ext of this halo is valid
    int32_t cache_ext;
    // incremental passes: the complete float64 potential of the previous pass, members it removed
    double *phi_keep;                     // [n_pad] by original slot
    double *phi_full;                     // [n_pad] by original slot: the potential of the last FULL pass
    float *rx, *ry, *rz, *rm;             // [n_pad] removed members of the last pass, per halo at poff
    int32_t *rem_cnt;                     // [n_halo]
    int32_t *incr;                        // [n_halo] the coming / current pass of the halo is incremental
    int32_t incr_enabled;
    int32_t targets_only;                 // members are targets only (the f2py-level cross call): counters follow
    // CUDA-graph loop driver: conditional handle of the WHILE node (0 = not in a graph)
    unsigned long long cond_handle;
};

cudaError_t launch_pack_members(const LoopParams &p, int sm_count, cudaStream_t s);
cudaError_t launch_pack_group(const HaloDesc *halo, int n_halo, int seg_index, int max_count,
                              const int64_t *ext_off, const double *m, const double *x, const double *y,
                              const double *z, float *em, float *ex, float *ey, float *ez, cudaStream_t s);
cudaError_t launch_energy_flag(const LoopParams &p, int sm_count, cudaStream_t s);
cudaError_t launch_halo_decide_init(const LoopParams &p, int sm_count, cudaStream_t s);
cudaError_t launch_compact(const LoopParams &p, int sm_count, cudaStream_t s);
cudaError_t launch_schedule(const LoopParams &p, int init, int sm_count, cudaStream_t s);
cudaError_t launch_fold_partials(const LoopParams &p, int sm_count, cudaStream_t s);
cudaError_t launch_sync_redo(const LoopParams &p, cudaStream_t s);
cudaError_t launch_set_nsplit_one(const LoopParams &p, cudaStream_t s);
cudaError_t launch_finalize(const LoopParams &p, int sm_count, cudaStream_t s);

}  // namespace halma
