// Host-visible interface of sortprep.cu (sorted source copies for the correction tickets).
#pragma once
#include <algorithm>
#include <cuda_runtime.h>

#include "halma_common.cuh"

namespace halma {

struct SortedAxisMut {           // owning view of one axis (device pointers)
    float *x, *y, *z, *m, *m0;   // m0: pristine masses, copied into m at the start of every run
    uint32_t *key;
    int32_t *slot, *tgt, *inv;
    double *corr;
};

struct SortedBuild {
    const HaloDesc *halo;
    const int32_t *src_halo;     // [n_tot] halo of every source id, -1 for padding
    F32Set mem, ext;             // working buffer 0 after the pack, external sources
    int64_t n_pad, n_tot;        // member slots, member slots + external slots
    int64_t n_valid;             // real sources (members + externals)
    int64_t n_spad;              // length of the padded sorted copies
    int32_t n_halo;
    uint64_t *keys_in, *keys_out;
    uint32_t *ids_in, *ids_out;
    uint8_t *is_member;          // [n_spad]
    int32_t *n_selected;         // device scalar (cub::DeviceSelect output count)
    void *temp;
    size_t temp_bytes;
    SortedAxisMut *out;
};

cudaError_t sorted_fill_src_halo(const HaloDesc *halo, int n_halo, const int32_t *chunk_halo, const int32_t *chunk_p0,
                                 int n_chunks, int max_ext, int64_t n_pad, int64_t n_tot, int32_t *src_halo,
                                 cudaStream_t s);
size_t sorted_temp_bytes(int64_t n_tot, int64_t n_spad);
cudaError_t sorted_build_axis(const SortedBuild &b, int axis, cudaStream_t s);

}  // namespace halma
