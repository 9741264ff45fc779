// Stand-alone kernels of the O(N) phases of the unbinding loop (kernels 2 and 3 and the device-side
// scheduler), sm_100a: thin wrappers around the phase functions of loop_device.cuh, used by the
// multi-launch drivers (enqueue-ahead, CUDA graph, split mode).  The default single-GPU driver runs
// the same functions inside ONE persistent kernel (fused.cu).
//
// One pass of the multi-launch loop is five launches on one stream, none of which
// needs the host:
//   k_potential_*   (potential.cu)  Phi for every current member of every active halo (+ the
//                   predicated re-launch for haloes the predicate-free pass flagged)
//   k_energy_flag   energy step + bound flag + per-chunk survivor count and mass sums; the block that
//                   finishes a halo's last chunk takes the halo's decision (new count, M, CoM, bulk
//                   velocity, converged / active, kind of the coming pass)
//   k_compact       stable warp-aggregated stream compaction into the other buffer
//   k_schedule      commits the per-halo state, builds the ticket table of the next pass
#include "loop_device.cuh"

namespace halma {

__global__ void __launch_bounds__(kLT) k_pack_members(const LoopParams p) { pack_phase(p); }

// ---------------------------------------------------------------------------------------
// One external group: user CSR layout (float64) -> padded float32 segment of each halo.
__global__ void __launch_bounds__(256) k_pack_group(const HaloDesc *halo, int n_halo, int seg_index,
                                                    const int64_t *ext_off, const double *m, const double *x,
                                                    const double *y, const double *z, float *em, float *ex,
                                                    float *ey, float *ez)
{
    for (int h = blockIdx.y; h < n_halo; h += gridDim.y) {
        const SegDesc sd = halo[h].seg[seg_index];
        const int64_t u0 = ext_off[h];
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < sd.count; k += gridDim.x * blockDim.x) {
            em[sd.begin + k] = __double2float_rn(m[u0 + k]);
            ex[sd.begin + k] = __double2float_rn(x[u0 + k]);
            ey[sd.begin + k] = __double2float_rn(y[u0 + k]);
            ez[sd.begin + k] = __double2float_rn(z[u0 + k]);
        }
    }
}


__global__ void __launch_bounds__(kLT) k_energy_flag(const LoopParams p)
{
    if (!p.st->any_active) return;
    energy_phase(p, p.st->parity, p.st->pass);
}

// Right after k_pack_members (no pass made yet); the passes decide inside k_energy_flag.
__global__ void __launch_bounds__(kLT) k_halo_decide_init(const LoopParams p) { decide_init_phase(p); }

__global__ void __launch_bounds__(kLT) k_compact(const LoopParams p)
{
    if (!p.st->any_active) return;
    compact_phase(p, p.st->parity, p.st->pass, false);
}

// Every block commits its share of the haloes; block 0 also builds the ticket table.
__global__ void __launch_bounds__(kLT) k_schedule(const LoopParams p, int init)
{
    __shared__ __align__(16) unsigned char smem[kSchedSmemBytes];
    if (!init && !p.st->any_active) {
        // graph driver: nothing left to do, leave the WHILE node
        if (blockIdx.x == 0 && threadIdx.x == 0 && p.cond_handle) cudaGraphSetConditional(p.cond_handle, 0u);
        return;
    }
    commit_phase(p, init);
    if (blockIdx.x == 0) schedule_block(p, smem, init);
}

// ---------------------------------------------------------------------------------------
// Split mode: fold the j-split partials into plane 0 before the all-reduce, and clear the
// entries this rank does not own (they are filled by the reduction with exact zeros added).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fold_partials(const LoopParams p)
{
    if (!p.st->any_active) return;
    for (int c = blockIdx.x; c < p.n_chunks; c += gridDim.x) {
        const int h = p.chunk_halo[c];
        if (!p.active[h]) continue;
        const int n = p.cnt[h];
        const int q = p.chunk_p0[c] + threadIdx.x;
        if (q >= n || threadIdx.x >= p.chunk) continue;      // (a block covers the largest chunk size)
        const int64_t i = p.halo[h].poff + q;
        const int owner = (q / p.group_size) % p.n_ranks;
        double phi = 0.0;
        if (owner == p.rank) {
            const int S = p.nsplit[p.rank_of[h]];
            phi = p.phi_part[i];
            for (int k = 1; k < S; ++k) phi += p.phi_part[static_cast<int64_t>(k) * p.n_pad + i];
        }
        p.phi_part[i] = phi;
        if (p.cache_ext && p.st->pass == 0 && p.halo[h].n_ext > 0) {
            // external-sum cache: the first pass's sums over the external sources are exchanged once
            // (in the first pass slot i is the member's original slot)
            double e = 0.0;
            if (owner == p.rank) {
                const int S = p.nsplit[p.rank_of[h]];
                e = p.phi_ext[i];
                for (int k = 1; k < S; ++k) e += p.phi_ext[static_cast<int64_t>(k) * p.n_pad + i];
            }
            p.phi_ext[i] = e;
        }
    }
}

// Split mode, predicate-free path: after the max-all-reduce of the per-halo fallback flags
// every rank derives the same redo_any.
__global__ void k_sync_redo(const LoopParams p)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int any = 0;
        for (int h = 0; h < p.n_halo; ++h) any |= p.halo_redo[h];
        p.st->redo_any = any;
    }
}

// After the all-reduce plane 0 holds the complete Phi: tell k_energy_flag to read only it.
__global__ void k_set_nsplit_one(const LoopParams p)
{
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < p.n_halo; h += gridDim.x * blockDim.x)
        p.nsplit[h] = 1;
}


__global__ void __launch_bounds__(kLT) k_finalize(const LoopParams p) { finalize_phase(p); }

// ---------------------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------------------
// one warp per chunk, grid-stride
static inline int chunk_grid(const LoopParams &p, int sm_count)
{
    const int want = sm_count * 16, need = (p.n_chunks + kLW - 1) / kLW;
    return need < want ? (need > 0 ? need : 1) : want;
}

cudaError_t launch_pack_members(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_pack_members<<<chunk_grid(p, sm_count), kLT, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_pack_group(const HaloDesc *halo, int n_halo, int seg_index, int max_count, const int64_t *ext_off,
                              const double *m, const double *x, const double *y, const double *z, float *em,
                              float *ex, float *ey, float *ez, cudaStream_t s)
{
    int gx = (max_count + 255) / 256;
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    int gy = n_halo < 1 ? 1 : (n_halo > 4096 ? 4096 : n_halo);
    if (gx * gy > 65536) gx = 65536 / gy > 0 ? 65536 / gy : 1;
    k_pack_group<<<dim3(gx, gy), 256, 0, s>>>(halo, n_halo, seg_index, ext_off, m, x, y, z, em, ex, ey, ez);
    return cudaGetLastError();
}

cudaError_t launch_energy_flag(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_energy_flag<<<chunk_grid(p, sm_count), kLT, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_halo_decide_init(const LoopParams &p, int sm_count, cudaStream_t s)
{
    const int want = sm_count * 16, need = (p.n_halo + kLW - 1) / kLW;
    const int g = need < want ? (need > 0 ? need : 1) : want;
    k_halo_decide_init<<<g, kLT, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_compact(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_compact<<<chunk_grid(p, sm_count), kLT, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_schedule(const LoopParams &p, int init, int sm_count, cudaStream_t s)
{
    // the commit is one thread per halo; block 0 alone builds the ticket table
    const int want = (p.n_halo + kLT - 1) / kLT;
    const int g = want < 1 ? 1 : (want > sm_count * 8 ? sm_count * 8 : want);
    k_schedule<<<g, kLT, 0, s>>>(p, init);
    return cudaGetLastError();
}

cudaError_t launch_fold_partials(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_fold_partials<<<chunk_grid(p, sm_count), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_sync_redo(const LoopParams &p, cudaStream_t s)
{
    k_sync_redo<<<1, 32, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_set_nsplit_one(const LoopParams &p, cudaStream_t s)
{
    const int g = (p.n_halo + 255) / 256;
    k_set_nsplit_one<<<g < 1 ? 1 : g, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const LoopParams &p, int sm_count, cudaStream_t s)
{
    k_finalize<<<chunk_grid(p, sm_count), kLT, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace halma
