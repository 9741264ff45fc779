// Sorted source copies for the correction tickets of the predicate-free FAST path
// (potential.cu).  Built once per upload, on the plan's stream:
//   for each axis a in {x, y, z}:
//     key64 = (halo << 32) | canonical bits of coordinate a   for every source of the plan
//     radix sort (CUB) of (key64, source id)
//     scatter into per-halo segments that start on 16-byte boundaries:
//       sorted x, y, z, m, key32, member slot (or -1), and slot -> sorted position
//     list of the sorted positions of the members (stream compaction, CUB)
// Canonical bits: the float's bit pattern, with -0 mapped to +0 so that coordinates the
// reference's `/=` treats as equal land in the same run.  NaNs keep their bits; the
// correction body re-checks equality with a float compare, so a NaN never matches.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "loop_kernels.h"
#include "sortprep.h"

namespace halma {

namespace {

__device__ __forceinline__ uint32_t canon_bits(float v)
{
    const uint32_t b = __float_as_uint(v);
    return (b << 1) == 0u ? 0u : b;
}

// source ids: [0, n_pad) member slots (working buffer 0), [n_pad, n_pad + n_ext_pad) externals
__global__ void __launch_bounds__(256) k_member_halo(const HaloDesc *halo, const int32_t *chunk_halo,
                                                     const int32_t *chunk_p0, int n_chunks, int32_t *src_halo)
{
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int h = chunk_halo[c];
        const int q = chunk_p0[c] + threadIdx.x;
        if (q < halo[h].n0) src_halo[halo[h].poff + q] = h;
    }
}

__global__ void __launch_bounds__(256) k_ext_halo(const HaloDesc *halo, int n_halo, int64_t n_pad, int32_t *src_halo)
{
    for (int h = blockIdx.y; h < n_halo; h += gridDim.y) {
        const HaloDesc &hd = halo[h];
        for (int s = 0; s < hd.nseg; ++s) {
            const SegDesc sd = hd.seg[s];
            if (sd.flags & kSegMembers) continue;
            for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < sd.count; k += gridDim.x * blockDim.x)
                src_halo[n_pad + sd.begin + k] = h;
        }
    }
}

__global__ void __launch_bounds__(256) k_make_keys(const int32_t *src_halo, const float *mem_c, const float *ext_c,
                                                   int64_t n_pad, int64_t n_tot, uint32_t n_halo, uint64_t *keys, uint32_t *ids)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_tot;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int32_t h = src_halo[i];
        // padding slots sort to the end: halo id one past the last (the sort looks at the bits that can differ only)
        uint64_t key = (static_cast<uint64_t>(n_halo) << 32) | 0xffffffffull;
        if (h >= 0) {
            const float v = i < n_pad ? mem_c[i] : ext_c[i - n_pad];
            key = (static_cast<uint64_t>(static_cast<uint32_t>(h)) << 32) | canon_bits(v);
        }
        keys[i] = key;
        ids[i] = static_cast<uint32_t>(i);
    }
}

__global__ void __launch_bounds__(256) k_scatter_sorted(const HaloDesc *halo, const uint64_t *keys, const uint32_t *ids,
                                                        int64_t n_valid, int64_t n_pad, F32Set mem, F32Set ext,
                                                        float *sx, float *sy, float *sz, float *sm, float *sm0,
                                                        uint32_t *key32, int32_t *slot, int32_t *inv, uint8_t *is_member)
{
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n_valid;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint64_t key = keys[k];
        const int h = static_cast<int>(key >> 32);
        const int64_t id = ids[k];
        const int64_t d = halo[h].sbegin + (k - halo[h].dbegin);
        float x, y, z, m;
        if (id < n_pad) {
            x = mem.x[id]; y = mem.y[id]; z = mem.z[id]; m = mem.m[id];
            slot[d] = static_cast<int32_t>(id);
            inv[id] = static_cast<int32_t>(d);
            is_member[d] = 1;
        } else {
            const int64_t e = id - n_pad;
            x = ext.x[e]; y = ext.y[e]; z = ext.z[e]; m = ext.m[e];
            slot[d] = -1;
        }
        sx[d] = x; sy[d] = y; sz[d] = z; sm[d] = m; sm0[d] = m;
        key32[d] = static_cast<uint32_t>(key);
    }
}

}  // namespace

cudaError_t sorted_fill_src_halo(const HaloDesc *halo, int n_halo, const int32_t *chunk_halo, const int32_t *chunk_p0,
                                 int n_chunks, int max_ext, int64_t n_pad, int64_t n_tot, int32_t *src_halo,
                                 cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(src_halo, 0xFF, n_tot * sizeof(int32_t), s);
    if (e != cudaSuccess) return e;
    if (n_chunks > 0) k_member_halo<<<n_chunks < 4096 ? n_chunks : 4096, 256, 0, s>>>(halo, chunk_halo, chunk_p0, n_chunks, src_halo);
    if (n_halo > 0 && max_ext > 0) {
        int gx = (max_ext + 255) / 256;
        gx = gx > 256 ? 256 : gx;
        const int gy = n_halo > 4096 ? 4096 : n_halo;
        k_ext_halo<<<dim3(gx, gy), 256, 0, s>>>(halo, n_halo, n_pad, src_halo);
    }
    return cudaGetLastError();
}

size_t sorted_temp_bytes(int64_t n_tot, int64_t n_spad)
{
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, static_cast<const uint64_t *>(nullptr), static_cast<uint64_t *>(nullptr),
                                    static_cast<const uint32_t *>(nullptr), static_cast<uint32_t *>(nullptr),
                                    static_cast<int>(n_tot), 0, 64);
    cub::DeviceSelect::Flagged(nullptr, b, thrust::counting_iterator<int32_t>(0), static_cast<const uint8_t *>(nullptr),
                               static_cast<int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                               static_cast<int>(n_spad));
    return (a > b ? a : b) + 256;
}

cudaError_t sorted_build_axis(const SortedBuild &b, int axis, cudaStream_t s)
{
    const float *mem_c = axis == 0 ? b.mem.x : axis == 1 ? b.mem.y : b.mem.z;
    const float *ext_c = axis == 0 ? b.ext.x : axis == 1 ? b.ext.y : b.ext.z;
    const int blocks = static_cast<int>(std::min<int64_t>((b.n_tot + 255) / 256, 148 * 16));
    if (b.n_tot == 0) return cudaSuccess;
    k_make_keys<<<blocks, 256, 0, s>>>(b.src_halo, mem_c, ext_c, b.n_pad, b.n_tot, static_cast<uint32_t>(b.n_halo), b.keys_in,
                                       b.ids_in);
    size_t tb = b.temp_bytes;
    int end_bit = 32;
    while (end_bit < 64 && (static_cast<uint64_t>(b.n_halo) >> (end_bit - 32)) != 0) ++end_bit;
    // 32 coordinate bits + the bits of the halo id (0 .. n_halo, n_halo = the padding): fewer radix passes than 64
    cudaError_t e = cub::DeviceRadixSort::SortPairs(b.temp, tb, b.keys_in, b.keys_out, b.ids_in, b.ids_out,
                                                    static_cast<int>(b.n_tot), 0, end_bit, s);
    if (e != cudaSuccess) return e;
    // padding of the sorted copies: coordinates NaN (never equal to anything), mass 0, key ~0, slot -1
    const size_t NS = static_cast<size_t>(b.n_spad);
    SortedAxisMut &A = *b.out;
    if ((e = cudaMemsetAsync(A.x, 0xFF, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(A.y, 0xFF, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(A.z, 0xFF, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(A.m, 0, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(A.m0, 0, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(A.key, 0xFF, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(A.slot, 0xFF, NS * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(b.is_member, 0, NS, s)) != cudaSuccess) return e;
    if (b.n_valid > 0) {
        const int sb = static_cast<int>(std::min<int64_t>((b.n_valid + 255) / 256, 148 * 16));
        k_scatter_sorted<<<sb, 256, 0, s>>>(b.halo, b.keys_out, b.ids_out, b.n_valid, b.n_pad, b.mem, b.ext, A.x, A.y,
                                            A.z, A.m, A.m0, A.key, A.slot, A.inv, b.is_member);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    tb = b.temp_bytes;
    e = cub::DeviceSelect::Flagged(b.temp, tb, thrust::counting_iterator<int32_t>(0), b.is_member, A.tgt, b.n_selected,
                                   static_cast<int>(b.n_spad), s);
    return e;
}

}  // namespace halma
