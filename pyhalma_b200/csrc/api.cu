// C-ABI of libhalma_unbind.so (include/halma_unbind.h): error handling, device contexts,
// the f2py-level potential call and the unbinding plan with its device-resident loop.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "fused.h"
#include "loop_kernels.h"
#include "potential.h"
#include "sortprep.h"

using namespace halma;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            char b__[512];                                                                             \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                                        \
            return fail(HALMA_ERR_CUDA, b__);                                                          \
        }                                                                                              \
    } while (0)

extern "C" const char *halma_last_error(void) { return g_err.c_str(); }
extern "C" int halma_abi_version(void) { return HALMA_ABI_VERSION; }

// ---------------------------------------------------------------------------------------
// device contexts
// ---------------------------------------------------------------------------------------
struct DeviceCtx {
    bool ready = false;
    int sm_count = 0;
    int bps_exact = 0;            // resident potential blocks per SM, EXACT kernel
    int bps_fast[kMaxVariants] = {0};   // ... and per FAST kernel shape
    int blocks(int mode, int variant) const { return mode == HALMA_MODE_EXACT ? bps_exact : bps_fast[variant]; }
    cudaStream_t stream = nullptr;
    // grow-only scratch of the host-pointer potential call
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    std::mutex mu;
    // grow-only free list of small pinned buffers (the loop state a run reads back): cudaMallocHost synchronises the
    // device, so a plan borrows one for the duration of a run instead of owning one
    std::vector<void *> pinned_free;
    std::mutex pinned_mu;
};
constexpr size_t kPinnedSlotBytes = 4096;
static cudaError_t pinned_borrow(DeviceCtx *c, void **out)
{
    {
        std::lock_guard<std::mutex> lock(c->pinned_mu);
        if (!c->pinned_free.empty()) {
            *out = c->pinned_free.back();
            c->pinned_free.pop_back();
            return cudaSuccess;
        }
    }
    return cudaMallocHost(out, kPinnedSlotBytes);
}
static void pinned_return(DeviceCtx *c, void *p)
{
    std::lock_guard<std::mutex> lock(c->pinned_mu);
    c->pinned_free.push_back(p);
}

static DeviceCtx g_ctx[64];
static std::mutex g_ctx_mu;

static int get_ctx(int device, DeviceCtx **out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(HALMA_ERR_NO_DEVICE, std::string("no usable CUDA device: ") +
                                             (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                             " (libhalma_unbind has no CPU fallback)");
    }
    if (device < 0 || device >= n || device >= 64) return fail(HALMA_ERR_NO_DEVICE, "device index out of range");
    CU_TRY(cudaSetDevice(device));
    DeviceCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    if (!c.ready) {
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(HALMA_ERR_NO_DEVICE, "device is not sm_100 or newer; this library is built for sm_100a only");
        c.sm_count = prop.multiProcessorCount;
        CU_TRY(potential_configure(HALMA_MODE_EXACT, 0, &c.bps_exact));
        for (int v = 0; v < potential_num_variants() && v < kMaxVariants; ++v)
            CU_TRY(potential_configure(HALMA_MODE_FAST, v, &c.bps_fast[v]));
        CU_TRY(fused_configure());
        CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        CU_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;      // keep freed blocks in the pool for the next plan
        CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        c.ready = true;
    }
    *out = &c;
    return HALMA_OK;
}

// Used by shape.cu: device context (lazy init), its stream and SM count.
int halma_internal_ctx(int device, int *sm_count, cudaStream_t *stream)
{
    DeviceCtx *c;
    if (int rc = get_ctx(device, &c)) return rc;
    if (sm_count) *sm_count = c->sm_count;
    if (stream) *stream = c->stream;
    return HALMA_OK;
}
int halma_internal_fail(int code, const char *msg) { return fail(code, msg); }

// CUDA-event time of the kernels of the last halo_shape / sigma_projections / snapshot call on
// this thread (their entry points take host buffers, so wall time would mostly measure PCIe).
static thread_local double g_last_kernel_ms = 0.0;
void halma_internal_set_kernel_ms(double ms) { g_last_kernel_ms = ms; }
extern "C" double halma_last_kernel_ms(void) { return g_last_kernel_ms; }

extern "C" int halma_selftest_exact_arith(int device, int64_t n_random, uint64_t seed, int64_t *mismatches)
{
    if (!mismatches || n_random < 0) return fail(HALMA_ERR_INVALID, "bad argument");
    DeviceCtx *c;
    if (int rc = get_ctx(device, &c)) return rc;
    unsigned long long *d = nullptr, h = 0;
    CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d), sizeof h, c->stream));
    CU_TRY(cudaMemsetAsync(d, 0, sizeof h, c->stream));
    CU_TRY(potential_selftest_exact(n_random, seed, d, c->sm_count, c->stream));
    CU_TRY(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaFreeAsync(d, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    *mismatches = static_cast<int64_t>(h);
    return HALMA_OK;
}

extern "C" int halma_device_count(int *count)
{
    if (!count) return fail(HALMA_ERR_INVALID, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(HALMA_ERR_NO_DEVICE, cudaGetErrorString(e));
    }
    *count = n;
    return HALMA_OK;
}

extern "C" int halma_device_info(int device, char *name, int len, int *sm_count, int *clock_khz, int64_t *mem_bytes)
{
    DeviceCtx *c;
    int rc = get_ctx(device, &c);
    if (rc) return rc;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (name && len > 0) {
        strncpy(name, prop.name, len - 1);
        name[len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (clock_khz) CU_TRY(cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, device));
    if (mem_bytes) *mem_bytes = static_cast<int64_t>(prop.totalGlobalMem);
    return HALMA_OK;
}

extern "C" int halma_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail(HALMA_ERR_INVALID, "bad arguments");
    CU_TRY(cudaMallocHost(ptr, static_cast<size_t>(bytes > 0 ? bytes : 1)));
    return HALMA_OK;
}

extern "C" int halma_host_free(void *ptr)
{
    if (ptr) CU_TRY(cudaFreeHost(ptr));
    return HALMA_OK;
}

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
static inline int64_t up4(int64_t n) { return (n + 3) & ~int64_t(3); }
static inline size_t up256(size_t n) { return (n + 255) & ~size_t(255); }

// Device buffer from the device's default memory pool (stream-ordered).  The pool's release
// threshold is raised in get_ctx(), so memory freed by one plan is reused by the next
// without going back to the driver: one-shot calls do not pay cudaMalloc/cudaFree.
static thread_local cudaStream_t g_alloc_stream = nullptr;

template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n), s(o.s)
    {
        o.p = nullptr;
        o.n = 0;
    }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        s = g_alloc_stream;
        return cudaMallocAsync(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T), s);
    }
    void release()
    {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    ~DBuf() { release(); }
};

static int choose_split(int mode, int64_t n_tgt, int64_t n_src, int group, int target_items)
{
    if (mode != HALMA_MODE_FAST) return 1;
    const int64_t groups = (n_tgt + group - 1) / group;
    if (groups <= 0 || groups >= target_items) return 1;
    const int64_t want = (target_items + groups - 1) / groups;
    const int64_t cap = std::max<int64_t>(1, n_src / kMinSplitSources);
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, std::min<int64_t>(cap, kMaxSplit))));
}

// ---------------------------------------------------------------------------------------
// brute_force_binding_energy on device pointers
// ---------------------------------------------------------------------------------------
namespace {

struct PotWorkspaceHeader {      // mirrored at the start of the device workspace
    LoopState st;
    HaloDesc halo;
    int32_t order;
    int32_t nsplit;
    int32_t item_base[2];
};

constexpr size_t kPotHeaderBytes = 1024;
static_assert(sizeof(PotWorkspaceHeader) <= kPotHeaderBytes, "header too large");

__global__ void k_fold_to_f32(const double *__restrict__ phi_part, int64_t stride, int S, int64_t n,
                              float *__restrict__ out)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double phi = phi_part[i];
        for (int k = 1; k < S; ++k) phi += phi_part[k * stride + i];
        out[i] = __double2float_rn(phi);
    }
}

}  // namespace

extern "C" int64_t halma_potential_workspace_bytes(int64_t n_src, int64_t n_tgt)
{
    (void)n_src;
    if (n_tgt < 0) return 0;
    return static_cast<int64_t>(kPotHeaderBytes) + kMaxSplit * up4(n_tgt) * 8 + 256;
}

static int potential_dev(DeviceCtx *c, int mode, const float *sm, const float *sx, const float *sy, const float *sz,
                         int64_t n_src, const float *tx, const float *ty, const float *tz, int64_t n_tgt,
                         float *out_be, void *workspace, cudaStream_t stream)
{
    if (n_tgt == 0) return HALMA_OK;
    if (n_src == 0) {
        CU_TRY(cudaMemsetAsync(out_be, 0, n_tgt * sizeof(float), stream));
        return HALMA_OK;
    }
    const int variant = potential_pick_variant((n_tgt + 127) / 128, n_src, c->sm_count * c->bps_fast[0] * (kPotentialBlock / 32),
                                               0, false, kMaxSplit, kMinSplitSources);
    const int group = potential_group_size(mode, variant);
    const int grid = c->sm_count * c->blocks(mode, variant);
    const int S = choose_split(mode, n_tgt, n_src, group, kNominalTickets);
    const int64_t groups = (n_tgt + group - 1) / group;

    PotWorkspaceHeader hdr;
    memset(&hdr, 0, sizeof hdr);
    hdr.st.n_items = static_cast<int32_t>(groups * S);
    hdr.st.any_active = 1;
    hdr.halo.poff = 0;
    hdr.halo.uoff = 0;
    hdr.halo.n0 = static_cast<int32_t>(n_tgt);
    hdr.halo.nseg = 1;
    hdr.halo.n_ext = static_cast<int32_t>(n_src);
    hdr.halo.seg[0].begin = 0;
    hdr.halo.seg[0].count = static_cast<int32_t>(n_src);
    hdr.halo.seg[0].flags = kSegNewClass;
    hdr.order = 0;
    hdr.nsplit = S;
    hdr.item_base[0] = 0;
    hdr.item_base[1] = hdr.st.n_items;
    char *ws = static_cast<char *>(workspace);
    CU_TRY(cudaMemcpyAsync(ws, &hdr, sizeof hdr, cudaMemcpyHostToDevice, stream));
    PotWorkspaceHeader *d = reinterpret_cast<PotWorkspaceHeader *>(ws);

    PotParams p;
    memset(&p, 0, sizeof p);
    p.tx[0] = p.tx[1] = tx;
    p.ty[0] = p.ty[1] = ty;
    p.tz[0] = p.tz[1] = tz;
    p.src[2] = F32Set{sx, sy, sz, sm};
    p.halo = &d->halo;
    p.cnt = nullptr;
    p.order = &d->order;
    p.item_base = d->item_base;
    p.nsplit = &d->nsplit;
    p.st = &d->st;
    p.phi_part = reinterpret_cast<double *>(ws + kPotHeaderBytes);
    p.phi_stride = up4(n_tgt);
    p.n_halo = 1;
    p.tgt_members = 0;
    p.rank = 0;
    p.n_ranks = 1;
    CU_TRY(potential_launch(p, mode, variant, grid, stream));
    const int fb = static_cast<int>(std::min<int64_t>((n_tgt + 255) / 256, c->sm_count * 8));
    k_fold_to_f32<<<fb, 256, 0, stream>>>(p.phi_part, p.phi_stride, S, n_tgt, out_be);
    CU_TRY(cudaGetLastError());
    return HALMA_OK;
}

// halma_potential_f32 for large FAST calls, defined after the plan API below.
constexpr int kPlanPathDeclined = 1;
static bool potential_plan_worthwhile(int64_t n_src, int64_t n_tgt);
static int potential_via_plan(int device, const float *sm, const float *sx, const float *sy, const float *sz,
                              int64_t n_src, const float *tx, const float *ty, const float *tz, int64_t n_tgt,
                              float *out_be);

static int check_mode(int mode)
{
    if (mode != HALMA_MODE_FAST && mode != HALMA_MODE_EXACT) return fail(HALMA_ERR_INVALID, "unknown mode");
    return HALMA_OK;
}

extern "C" int halma_potential_f32_dev(int device, int mode, const float *src_m, const float *src_x,
                                       const float *src_y, const float *src_z, int64_t n_src, const float *tgt_x,
                                       const float *tgt_y, const float *tgt_z, int64_t n_tgt, float *out_be,
                                       void *workspace, void *stream)
{
    if (int rc = check_mode(mode)) return rc;
    if (n_src < 0 || n_tgt < 0) return fail(HALMA_ERR_INVALID, "negative size");
    if (n_src > 0x7ffffff0ll || n_tgt > 0x7ffffff0ll) return fail(HALMA_ERR_TOO_LARGE, "more than 2^31-16 particles");
    if (n_tgt == 0) return HALMA_OK;
    if (!out_be || !tgt_x || !tgt_y || !tgt_z || !workspace) return fail(HALMA_ERR_INVALID, "null pointer");
    if (n_src > 0 && (!src_m || !src_x || !src_y || !src_z)) return fail(HALMA_ERR_INVALID, "null source pointer");
    const void *al[] = {src_m, src_x, src_y, src_z, workspace};
    for (const void *q : al)
        if (reinterpret_cast<uintptr_t>(q) & 15) return fail(HALMA_ERR_ALIGN, "device pointer not 16-byte aligned");
    DeviceCtx *c;
    if (int rc = get_ctx(device, &c)) return rc;
    return potential_dev(c, mode, src_m, src_x, src_y, src_z, n_src, tgt_x, tgt_y, tgt_z, n_tgt, out_be, workspace,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int halma_potential_f32(int device, int mode, const float *src_m, const float *src_x, const float *src_y,
                                   const float *src_z, int64_t n_src, const float *tgt_x, const float *tgt_y,
                                   const float *tgt_z, int64_t n_tgt, float *out_be)
{
    if (int rc = check_mode(mode)) return rc;
    if (n_src < 0 || n_tgt < 0) return fail(HALMA_ERR_INVALID, "negative size");
    if (n_src > 0x7ffffff0ll || n_tgt > 0x7ffffff0ll) return fail(HALMA_ERR_TOO_LARGE, "more than 2^31-16 particles");
    if (n_tgt == 0) return HALMA_OK;
    if (!out_be || !tgt_x || !tgt_y || !tgt_z) return fail(HALMA_ERR_INVALID, "null pointer");
    if (n_src > 0 && (!src_m || !src_x || !src_y || !src_z)) return fail(HALMA_ERR_INVALID, "null source pointer");
    DeviceCtx *c;
    if (int rc = get_ctx(device, &c)) return rc;
    if (mode == HALMA_MODE_FAST && potential_plan_worthwhile(n_src, n_tgt)) {
        // large call: the predicate-free kernel (and, for a self call, the symmetric self-term) of the plans
        const int rc = potential_via_plan(device, src_m, src_x, src_y, src_z, n_src, tgt_x, tgt_y, tgt_z, n_tgt, out_be);
        if (rc != kPlanPathDeclined) return rc;
    }
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t sb = up256((n_src + 4) * sizeof(float)), tb = up256((n_tgt + 4) * sizeof(float));
    const size_t wb = up256(static_cast<size_t>(halma_potential_workspace_bytes(n_src, n_tgt)));
    const size_t need = 4 * sb + 4 * tb + wb;
    if (need > c->scratch_bytes) {
        if (c->scratch) CU_TRY(cudaFree(c->scratch));
        c->scratch = nullptr;
        c->scratch_bytes = 0;
        CU_TRY(cudaMalloc(&c->scratch, need + need / 4));
        c->scratch_bytes = need + need / 4;
    }
    char *base = static_cast<char *>(c->scratch);
    float *d_s[4], *d_t[4];
    for (int k = 0; k < 4; ++k) d_s[k] = reinterpret_cast<float *>(base + k * sb);
    for (int k = 0; k < 4; ++k) d_t[k] = reinterpret_cast<float *>(base + 4 * sb + k * tb);
    void *d_ws = base + 4 * sb + 4 * tb;
    cudaStream_t s = c->stream;
    const float *hs[4] = {src_m, src_x, src_y, src_z};
    for (int k = 0; k < 4 && n_src > 0; ++k)
        CU_TRY(cudaMemcpyAsync(d_s[k], hs[k], n_src * sizeof(float), cudaMemcpyHostToDevice, s));
    const float *ht[3] = {tgt_x, tgt_y, tgt_z};
    for (int k = 0; k < 3; ++k) CU_TRY(cudaMemcpyAsync(d_t[k], ht[k], n_tgt * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = potential_dev(c, mode, d_s[0], d_s[1], d_s[2], d_s[3], n_src, d_t[0], d_t[1], d_t[2], n_tgt, d_t[3],
                           d_ws, s);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(out_be, d_t[3], n_tgt * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return HALMA_OK;
}

// ---------------------------------------------------------------------------------------
// NCCL, loaded lazily (split mode only)
// ---------------------------------------------------------------------------------------
struct UID128 {      // ncclUniqueId: 128 opaque bytes, passed by value
    char internal[128];
};
namespace {
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, UID128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl()
{
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.handle) return HALMA_OK;
    // HALMA_NCCL_LIB lets the host side point at the NCCL build it already uses (the Python
    // package sets it to torch's bundled copy: loading an older system libnccl first would
    // break a later `import torch`).  A bare soname resolves to an already loaded copy.
    void *h = nullptr;
    if (const char *path = getenv("HALMA_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail(HALMA_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(h, "ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(HALMA_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl.handle = h;
    return HALMA_OK;
}

static int nccl_fail(int code, const char *what)
{
    return fail(HALMA_ERR_NCCL, std::string(what) + ": " +
                                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "nccl error"));
}

extern "C" int halma_nccl_unique_id(void *unique_id_128)
{
    if (!unique_id_128) return fail(HALMA_ERR_INVALID, "null pointer");
    if (int rc = load_nccl()) return rc;
    int e = g_nccl.GetUniqueId(unique_id_128);
    return e ? nccl_fail(e, "ncclGetUniqueId") : HALMA_OK;
}

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
struct halma_plan {
    halma_unbind_config cfg;
    DeviceCtx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    int64_t n_halo = 0, n_user = 0, n_pad = 0, n_ext_pad = 0;
    int n_chunks = 0;
    std::vector<int64_t> offsets;
    std::vector<std::vector<int64_t>> ext_offsets;
    std::vector<int> group_seg;           // segment index of each ext group inside HaloDesc::seg
    std::vector<int> group_max;           // largest per-halo count of each group
    std::vector<HaloDesc> host_halo;      // host copies of the static tables (sources of asynchronous uploads)
    std::vector<int32_t> host_order, host_rank_of, host_chunk_halo, host_chunk_p0;
    bool members_up = false, vb_up = false;
    std::vector<bool> group_up;

    // static tables
    DBuf<HaloDesc> d_halo;
    DBuf<int32_t> d_chunk_halo, d_chunk_p0, d_order;
    std::vector<DBuf<int64_t>> d_ext_off;
    // inputs
    DBuf<double> d_in;                    // 7 * n_user: x y z vx vy vz m
    DBuf<float> d_ext;                    // 4 * n_ext_pad: x y z m
    DBuf<double> d_stage;                 // staging for one ext group upload (4 * max group size)
    DBuf<double> d_vb_user;
    // working sets
    DBuf<float> d_work;                   // 2 * 4 * n_pad
    DBuf<int32_t> d_widx;                 // 2 * n_pad
    DBuf<int32_t> d_hint;                 // per-halo int arrays: 9 * n_halo + (n_halo + 1) + halo_done, halo_stamp
    DBuf<int32_t> d_rank_of, d_range;     // position of a halo in `order`; min | max of x, y, z per halo (3 + 3 ints)
    DBuf<int4> d_sched;                   // scheduling records, in `order` space
    DBuf<double> d_phi_full;
    int fused_index = -1;                 // persistent loop kernel serving this plan (fused.cu), -1: none
    int min_split_sources = kMinSplitSources;
    int driver_ran = HALMA_DRIVER_ENQUEUE;
    std::vector<cudaEvent_t> cev;         // split mode: 2 per pass around the potential collectives, 2 around the rest
    DBuf<double> d_hdbl;                  // per-halo doubles: M(1) vb(3) vb_next(3) com(3)
    DBuf<unsigned long long> d_pairs, d_evals;
    DBuf<double> d_phi_sym, d_symq;
    bool sym = false;
    int sym_rows = 4;
    // external-sum cache and incremental passes (FAST predicate-free path, one GPU)
    bool cache_ext = false, incr = false;
    DBuf<double> d_phi_ext, d_phi_keep;
    DBuf<float> d_rem;                    // 4 * n_pad: members removed by the last pass (x y z m)
    DBuf<int32_t> d_reuse_int;            // ext_ok, incr, rem_cnt: 3 * n_halo
    DBuf<int32_t> d_cint;                 // per-chunk ints: cnt, off
    DBuf<double> d_csum;                  // kChunkSums per chunk
    DBuf<float> d_cbest;
    DBuf<int32_t> d_cbestq, d_hbest;
    DBuf<double> d_hrps, d_temp;
    bool temp_up = false;
    double cold_T = 5e4;
    // CUDA-graph loop driver (cfg.use_graph): one graph launch runs every pass
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    bool graph_dirty = true;
    DBuf<uint8_t> d_flag, d_mask;
    DBuf<float> d_be;
    DBuf<double> d_E;
    DBuf<int32_t> d_idx;
    DBuf<double> d_phi;
    DBuf<LoopState> d_st;
    int32_t *h_flags = nullptr;           // pinned: any_active after each pass
    std::vector<cudaEvent_t> ev;          // per pass: before K1, after K1, end of pass
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    void *comm = nullptr;                 // ncclComm_t (split mode)
    bool comm_owned = false;
    bool ran = false;
    // predicate-free FAST path: sorted source copies (sortprep.cu)
    int variant = 0;                      // FAST kernel shape (potential.cu)
    bool np = false, sorted_dirty = true;
    int64_t n_spad = 0, n_valid = 0, n_tot = 0;
    int max_ext = 0;
    size_t sort_temp_bytes = 0;
    DBuf<int32_t> d_src_halo, d_sslot, d_stgt, d_sinv, d_redo, d_nsel;
    DBuf<uint64_t> d_keys;
    DBuf<uint32_t> d_ids, d_skey;
    DBuf<uint8_t> d_ismem, d_sorttemp;
    DBuf<float> d_sf;
    DBuf<double> d_corr;
    SortedAxisMut sax[3];
    LoopParams lp;
    PotParams pp;

    void free_buffers()
    {
        d_halo.release(); d_chunk_halo.release(); d_chunk_p0.release(); d_order.release();
        d_ext_off.clear();
        d_in.release(); d_ext.release(); d_stage.release(); d_vb_user.release(); d_work.release();
        d_widx.release(); d_hint.release(); d_hdbl.release(); d_pairs.release(); d_evals.release(); d_cint.release(); d_phi_sym.release(); d_symq.release();
        d_csum.release(); d_flag.release(); d_mask.release(); d_be.release(); d_E.release();
        d_idx.release(); d_phi.release(); d_st.release();
        d_cbest.release(); d_cbestq.release(); d_hbest.release(); d_hrps.release(); d_temp.release();
        d_src_halo.release(); d_sslot.release(); d_stgt.release(); d_sinv.release(); d_redo.release();
        d_nsel.release(); d_keys.release(); d_ids.release(); d_skey.release(); d_ismem.release();
        d_sorttemp.release(); d_sf.release(); d_corr.release();
        d_phi_ext.release(); d_phi_keep.release(); d_rem.release(); d_reuse_int.release();
        d_rank_of.release(); d_range.release(); d_sched.release(); d_phi_full.release();
    }
    ~halma_plan()
    {
        if (gexec) cudaGraphExecDestroy(gexec);
        if (graph) cudaGraphDestroy(graph);
        free_buffers();                       // stream-ordered frees, before the stream goes away
        if (stream) cudaStreamSynchronize(stream);
        if (comm && comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        for (auto e : ev) cudaEventDestroy(e);
        for (auto e : cev) cudaEventDestroy(e);
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_stop) cudaEventDestroy(ev_stop);
        if (h_flags) cudaFreeHost(h_flags);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int plan_build(halma_plan *P, const int64_t *offsets, const int64_t *const *ext_offsets)
{
    const halma_unbind_config &cfg = P->cfg;
    const int64_t nh = P->n_halo;
    g_alloc_stream = P->stream;
    P->offsets.assign(offsets, offsets + nh + 1);
    if (P->offsets[0] != 0) return fail(HALMA_ERR_INVALID, "offsets[0] must be 0");
    for (int64_t h = 0; h < nh; ++h)
        if (P->offsets[h + 1] < P->offsets[h]) return fail(HALMA_ERR_INVALID, "offsets must be non-decreasing");
    P->n_user = P->offsets[nh];
    if (P->n_user > 0x7ffffff0ll) return fail(HALMA_ERR_TOO_LARGE, "more than 2^31-16 members");
    P->ext_offsets.resize(cfg.n_groups);
    P->group_up.assign(cfg.n_groups, false);
    P->group_max.assign(cfg.n_groups, 0);
    for (int g = 0; g < cfg.n_groups; ++g) {
        if (!ext_offsets || !ext_offsets[g]) return fail(HALMA_ERR_INVALID, "ext_offsets missing for a group");
        P->ext_offsets[g].assign(ext_offsets[g], ext_offsets[g] + nh + 1);
        if (P->ext_offsets[g][0] != 0) return fail(HALMA_ERR_INVALID, "ext_offsets[g][0] must be 0");
        for (int64_t h = 0; h < nh; ++h) {
            const int64_t c = P->ext_offsets[g][h + 1] - P->ext_offsets[g][h];
            if (c < 0) return fail(HALMA_ERR_INVALID, "ext_offsets must be non-decreasing");
            if (c > 0x7ffffff0ll) return fail(HALMA_ERR_TOO_LARGE, "external group too large");
            P->group_max[g] = std::max<int>(P->group_max[g], static_cast<int>(c));
        }
    }

    // segment order (see include/halma_unbind.h): which HaloDesc::seg slot each group lands in
    P->group_seg.assign(cfg.n_groups, 0);
    const int members_slot = cfg.n_pre;
    for (int g = 0; g < cfg.n_groups; ++g) P->group_seg[g] = g < cfg.n_pre ? g : g + 1;

    // host tables live in the plan: their asynchronous uploads need no synchronisation here
    std::vector<HaloDesc> &halo = P->host_halo;
    halo.assign(nh, HaloDesc());
    std::vector<int32_t> &chunk_halo = P->host_chunk_halo, &chunk_p0 = P->host_chunk_p0;
    chunk_halo.clear();
    chunk_p0.clear();
    int64_t poff = 0, eoff = 0, soff = 0, doff = 0;
    const int chunk = P->n_user <= kChunkSmallMaxMembers ? kChunkSmall : kChunk;
    {
        const char *e = getenv("HALMA_NP");
        P->np = cfg.mode == HALMA_MODE_FAST && !(e && atoi(e) == 0);
    }
    for (int64_t h = 0; h < nh; ++h) {
        HaloDesc &d = halo[h];
        memset(&d, 0, sizeof d);
        const int64_t n0 = P->offsets[h + 1] - P->offsets[h];
        d.poff = poff;
        d.uoff = P->offsets[h];
        d.n0 = static_cast<int32_t>(n0);
        d.nseg = 1 + cfg.n_groups;
        d.chunk_begin = static_cast<int32_t>(chunk_halo.size());
        for (int64_t q = 0; q < n0; q += chunk) {
            chunk_halo.push_back(static_cast<int32_t>(h));
            chunk_p0.push_back(static_cast<int32_t>(q));
        }
        d.seg[members_slot].begin = poff;
        d.seg[members_slot].count = d.n0;
        d.seg[members_slot].flags = kSegMembers | (cfg.split_classes ? kSegNewClass : 0);
        int64_t next = 0;
        for (int g = 0; g < cfg.n_groups; ++g) {
            const int64_t c = P->ext_offsets[g][h + 1] - P->ext_offsets[g][h];
            SegDesc &s = d.seg[P->group_seg[g]];
            s.begin = eoff;
            s.count = static_cast<int32_t>(c);
            s.flags = cfg.split_classes ? kSegNewClass : 0;
            eoff += up4(c);
            next += c;
        }
        if (next > 0x7ffffff0ll) return fail(HALMA_ERR_TOO_LARGE, "too many external sources for one halo");
        d.n_ext = static_cast<int32_t>(next);
        d.sbegin = soff;
        d.dbegin = doff;
        soff += up4(n0 + next);
        doff += n0 + next;
        P->max_ext = std::max<int>(P->max_ext, static_cast<int>(std::min<int64_t>(next, 0x7fffffff)));
        poff += up4(n0);
    }
    P->n_spad = soff + 16;
    P->n_valid = doff;
    P->n_pad = poff + 16;
    P->n_ext_pad = eoff + 16;
    P->n_chunks = static_cast<int>(chunk_halo.size());
    P->n_tot = P->n_pad + P->n_ext_pad;
    if (P->n_tot > 0x7ffffff0ll || P->n_spad > 0x7ffffff0ll || P->n_user == 0) P->np = false;
    {
        // The predicate-free path (and with it the symmetric self-term) pays off from ~1e9 pairs per
        // pass; below that its one-off sort (48 small launches), extra tickets and launch cost more
        // than they save on a one-shot plan (profiles/midsize_probe_r01.txt).
        double pairs = 0.0;
        for (int64_t h = 0; h < nh; ++h) pairs += static_cast<double>(halo[h].n0) * (halo[h].n0 + halo[h].n_ext);
        const char *e = getenv("HALMA_NP_MIN_PAIRS");
        const double min_pairs = e ? atof(e) : 1e9;
        if (pairs < min_pairs) P->np = false;
    }

    std::vector<int32_t> &order = P->host_order;
    order.assign(nh, 0);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return halo[a].n0 > halo[b].n0; });

    cudaStream_t s = P->stream;
    CU_TRY(P->d_halo.alloc(nh));
    CU_TRY(P->d_chunk_halo.alloc(chunk_halo.size()));
    CU_TRY(P->d_chunk_p0.alloc(chunk_p0.size()));
    CU_TRY(P->d_order.alloc(nh));
    CU_TRY(P->d_rank_of.alloc(nh));
    CU_TRY(P->d_sched.alloc(nh));
    CU_TRY(P->d_range.alloc(6 * nh));
    std::vector<int32_t> &rank_of = P->host_rank_of;
    rank_of.assign(nh, 0);
    for (int64_t k = 0; k < nh; ++k) rank_of[order[k]] = static_cast<int32_t>(k);
    if (nh) {
        CU_TRY(cudaMemcpyAsync(P->d_rank_of.p, rank_of.data(), nh * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(P->d_halo.p, halo.data(), nh * sizeof(HaloDesc), cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(P->d_order.p, order.data(), nh * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
    if (P->n_chunks) {
        CU_TRY(cudaMemcpyAsync(P->d_chunk_halo.p, chunk_halo.data(), chunk_halo.size() * 4, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(P->d_chunk_p0.p, chunk_p0.data(), chunk_p0.size() * 4, cudaMemcpyHostToDevice, s));
    }
    P->d_ext_off.resize(cfg.n_groups);
    int64_t stage = 0;
    for (int g = 0; g < cfg.n_groups; ++g) {
        CU_TRY(P->d_ext_off[g].alloc(nh + 1));
        CU_TRY(cudaMemcpyAsync(P->d_ext_off[g].p, P->ext_offsets[g].data(), (nh + 1) * 8, cudaMemcpyHostToDevice, s));
        stage = std::max<int64_t>(stage, P->ext_offsets[g][nh]);
    }
    // the host tables stay alive in the plan until the copies have certainly been made (first run)

    const size_t NU = static_cast<size_t>(P->n_user), NP = static_cast<size_t>(P->n_pad);
    CU_TRY(P->d_in.alloc(7 * NU));
    CU_TRY(P->d_ext.alloc(4 * static_cast<size_t>(P->n_ext_pad)));
    CU_TRY(P->d_stage.alloc(4 * static_cast<size_t>(stage)));
    CU_TRY(P->d_vb_user.alloc(3 * nh));
    CU_TRY(P->d_work.alloc(8 * NP));
    CU_TRY(P->d_widx.alloc(2 * NP));
    CU_TRY(P->d_hint.alloc(12 * nh + 1));
    CU_TRY(P->d_hdbl.alloc(10 * nh));
    CU_TRY(P->d_pairs.alloc(nh));
    CU_TRY(P->d_evals.alloc(nh));
    CU_TRY(P->d_cint.alloc(2 * static_cast<size_t>(P->n_chunks)));
    CU_TRY(P->d_csum.alloc(kChunkSums * static_cast<size_t>(P->n_chunks)));
    CU_TRY(P->d_cbest.alloc(static_cast<size_t>(P->n_chunks)));
    CU_TRY(P->d_cbestq.alloc(static_cast<size_t>(P->n_chunks)));
    CU_TRY(P->d_hrps.alloc(4 * nh));
    CU_TRY(P->d_hbest.alloc(nh));
    CU_TRY(P->d_temp.alloc(NU));
    CU_TRY(P->d_flag.alloc(NP));
    CU_TRY(P->d_mask.alloc(NU));
    CU_TRY(P->d_be.alloc(NU));
    CU_TRY(P->d_E.alloc(NU));
    CU_TRY(P->d_idx.alloc(NU));
    // j-split planes: small plans (where one pass cannot fill the machine with 128-target tickets otherwise) may
    // cut a halo's sources into up to 32 pieces of >= 512 sources; large ones into 8 pieces of >= 2048
    const bool small_plan = P->n_user <= kChunkSmallMaxMembers;
    const int planes = cfg.mode == HALMA_MODE_FAST ? (small_plan ? kMaxSplitSmall : kMaxSplit) : 1;
    P->min_split_sources = small_plan ? kMinSplitSourcesSmall : kMinSplitSources;
    CU_TRY(P->d_phi.alloc(planes * NP));
    CU_TRY(P->d_st.alloc(1));
    CU_TRY(P->d_redo.alloc(nh));
    if (P->np) {
        const size_t NT = static_cast<size_t>(P->n_tot), NS = static_cast<size_t>(P->n_spad);
        CU_TRY(P->d_src_halo.alloc(NT));
        CU_TRY(P->d_keys.alloc(2 * NT));
        CU_TRY(P->d_ids.alloc(2 * NT));
        CU_TRY(P->d_ismem.alloc(NS));
        CU_TRY(P->d_nsel.alloc(4));
        P->sort_temp_bytes = sorted_temp_bytes(P->n_tot, P->n_spad);
        CU_TRY(P->d_sorttemp.alloc(P->sort_temp_bytes));
        CU_TRY(P->d_sf.alloc(15 * NS));
        CU_TRY(P->d_skey.alloc(3 * NS));
        CU_TRY(P->d_sslot.alloc(3 * NS));
        CU_TRY(P->d_stgt.alloc(3 * std::max<size_t>(NU, 1)));
        CU_TRY(P->d_sinv.alloc(3 * NP));
        CU_TRY(P->d_corr.alloc(3 * NP));
        CU_TRY(cudaMemsetAsync(P->d_corr.p, 0, 3 * NP * sizeof(double), s));
        CU_TRY(cudaMemsetAsync(P->d_sinv.p, 0, 3 * NP * sizeof(int32_t), s));
        for (int a = 0; a < 3; ++a) {
            SortedAxisMut &A = P->sax[a];
            float *f = P->d_sf.p + static_cast<size_t>(a) * 5 * NS;
            A.x = f;
            A.y = f + NS;
            A.z = f + 2 * NS;
            A.m = f + 3 * NS;
            A.m0 = f + 4 * NS;
            A.key = P->d_skey.p + a * NS;
            A.slot = P->d_sslot.p + a * NS;
            A.tgt = P->d_stgt.p + a * std::max<size_t>(NU, 1);
            A.inv = P->d_sinv.p + a * NP;
            A.corr = P->d_corr.p + a * NP;
        }
    }
    CU_TRY(cudaMemsetAsync(P->d_work.p, 0, 8 * NP * sizeof(float), s));
    CU_TRY(cudaMemsetAsync(P->d_ext.p, 0, 4 * static_cast<size_t>(P->n_ext_pad) * sizeof(float), s));
    CU_TRY(cudaMemsetAsync(P->d_vb_user.p, 0, 3 * nh * sizeof(double), s));
    // (the pinned convergence flags of the enqueue-ahead driver are allocated on first use: cudaMallocHost and
    // cudaFreeHost synchronise the device, which would serialise plans that run side by side)
    // events of the multi-launch drivers are created on first use (plan_run)
    CU_TRY(cudaEventCreate(&P->ev_start));
    CU_TRY(cudaEventCreate(&P->ev_stop));

    // parameter blocks
    LoopParams &L = P->lp;
    memset(&L, 0, sizeof L);
    L.halo = P->d_halo.p;
    L.chunk_halo = P->d_chunk_halo.p;
    L.chunk_p0 = P->d_chunk_p0.p;
    L.order = P->d_order.p;
    L.n_halo = static_cast<int32_t>(nh);
    L.n_chunks = P->n_chunks;
    L.chunk = P->n_user <= kChunkSmallMaxMembers ? kChunkSmall : kChunk;
    L.n_pad = P->n_pad;
    L.n_user = P->n_user;
    const double *in = P->d_in.p;
    L.x64 = in;
    L.y64 = in + NU;
    L.z64 = in + 2 * NU;
    L.vx = in + 3 * NU;
    L.vy = in + 4 * NU;
    L.vz = in + 5 * NU;
    L.m64 = in + 6 * NU;
    for (int b = 0; b < 2; ++b) {
        float *w = P->d_work.p + b * 4 * NP;
        L.wx[b] = w;
        L.wy[b] = w + NP;
        L.wz[b] = w + 2 * NP;
        L.wm[b] = w + 3 * NP;
        L.widx[b] = P->d_widx.p + b * NP;
    }
    int32_t *hi = P->d_hint.p;
    L.cnt = hi;
    L.cnt_next = hi + nh;
    L.iter = hi + 2 * nh;
    L.active = hi + 3 * nh;
    L.active_next = hi + 4 * nh;
    L.halo_buf = hi + 5 * nh;
    L.nsplit = hi + 6 * nh;
    L.converged = hi + 7 * nh;
    L.item_base = hi + 9 * nh;       // n_halo + 1 entries (slot 8 is spare)
    L.halo_done = hi + 10 * nh + 1;
    L.halo_stamp = hi + 11 * nh + 1;
    L.rank_of = P->d_rank_of.p;
    L.sched = P->d_sched.p;
    L.halo_rmin = P->d_range.p;
    L.halo_rmax = P->d_range.p + 3 * nh;
    double *hd = P->d_hdbl.p;
    L.hM = hd;
    L.hvb = hd + nh;
    L.hvb_next = hd + 4 * nh;
    L.hcom = hd + 7 * nh;
    L.pairs = P->d_pairs.p;
    L.evals = P->d_evals.p;
    L.chunk_cnt = P->d_cint.p;
    L.chunk_off = P->d_cint.p + P->n_chunks;
    L.chunk_sum = P->d_csum.p;
    L.chunk_best = P->d_cbest.p;
    L.chunk_best_q = P->d_cbestq.p;
    L.temp = nullptr;
    L.cold_T = 5e4;
    L.hrps = P->d_hrps.p;
    L.hbest = P->d_hbest.p;
    L.flag = P->d_flag.p;
    L.out_mask = P->d_mask.p;
    L.out_be = P->d_be.p;
    L.out_E = P->d_E.p;
    L.out_idx = P->d_idx.p;
    L.phi_part = P->d_phi.p;
    L.st = P->d_st.p;
    L.G32 = static_cast<float>(cfg.G);           // numpy: float32 array *= python float
    L.kappa32 = static_cast<float>(cfg.kappa);
    L.vb_fixed = cfg.vb_fixed;
    L.max_iter = cfg.max_iter;
    L.mode = cfg.mode;
    {
        int64_t g128 = 0, max_src = 0, max_n0 = 0;
        for (int64_t h = 0; h < nh; ++h) {
            g128 += (halo[h].n0 + 127) / 128;
            max_src = std::max<int64_t>(max_src, static_cast<int64_t>(halo[h].n0) + halo[h].n_ext);
            max_n0 = std::max<int64_t>(max_n0, halo[h].n0);
        }
        P->variant = potential_pick_variant(g128, max_src, P->ctx->sm_count * P->ctx->bps_fast[0] * (kPotentialBlock / 32),
                                            max_n0, cfg.symmetric != 0 && P->np && cfg.mode == HALMA_MODE_FAST, planes,
                                            P->min_split_sources);
    }
    L.group_size = potential_group_size(cfg.mode, P->variant);
    L.rank = cfg.rank;
    L.n_ranks = cfg.n_ranks;
    L.target_items = kNominalTickets;
    L.max_split = planes;
    L.min_split_sources = P->min_split_sources;
    L.halo_redo = P->d_redo.p;
    L.np_enabled = P->np ? 1 : 0;
    L.redo_enabled = 0;      // set below, once the reuse options are known
    for (int a = 0; a < 3 && P->np; ++a) {
        const SortedAxisMut &A = P->sax[a];
        L.ax[a] = SortedAxis{A.x, A.y, A.z, A.m, A.key, A.slot, A.tgt, A.inv, A.corr};
        L.ax_m0[a] = A.m0;
    }
    L.n_spad = P->np ? P->n_spad : 0;
    // symmetric self-term: rides on the predicate-free kernel's throughput shape (128-member tiles)
    P->sym = cfg.symmetric && P->np && P->variant == 0 && potential_group_size(cfg.mode, P->variant) == 128;
    if (P->sym) {
        CU_TRY(P->d_phi_sym.alloc(NP));
        CU_TRY(P->d_symq.alloc(2 * nh));
    }
    L.phi_sym = P->d_phi_sym.p;
    L.sym_enabled = P->sym ? 1 : 0;
    {
        // Row members per lane of the symmetric tickets: 4 (one row tile) or 8 (a pair of row tiles: half the
        // shared-memory loads per evaluation and the next step's sources prefetched across the warp barrier, 5.9
        // instead of 6.75 issue slots per rsqrt, at 96 registers = 5 blocks per SM instead of 6).  8 wins wherever
        // the member x member pairs are a good part of the work (one 1e6-star pass 163.6 -> 156.2 ms, the cfg3
        // catalogue 33.7 -> 33.3 ms, cfg2's gas 84.0 -> 82.4 ms) and loses a little where the one-sided code over
        // external sources dominates, which runs on the same, fewer, blocks (cfg2's stars, 14 % member pairs: +0.8 %;
        // profiles/sym_rows_ab_r02.txt): 8 from ~30 % of the evaluations on.  The rule depends on the problem only,
        // so a split run picks what the one-GPU run picks.  HALMA_SYM_ROWS=4|8 overrides it.
        double self_pairs = 0.0, ext_pairs = 0.0;
        for (int64_t h = 0; h < nh; ++h) {
            const double n0 = static_cast<double>(halo[h].n0);
            self_pairs += n0 * n0;
            ext_pairs += n0 * halo[h].n_ext;
        }
        P->sym_rows = self_pairs >= 0.85 * ext_pairs ? 8 : 4;      // member pairs are evaluated once: n0^2 / 2 each
        if (const char *e = getenv("HALMA_SYM_ROWS")) P->sym_rows = atoi(e) == 8 ? 8 : (atoi(e) == 4 ? 4 : P->sym_rows);
    }
    L.sym_rows = P->sym_rows;
    L.sym_ext = P->d_symq.p;
    L.sym_q = P->sym ? P->d_symq.p + nh : nullptr;
    // External-sum cache and incremental passes ride on the predicate-free path (whose sums are kept per
    // member in float64 and whose excluded pairs are handled by the correction tickets every pass).
    // Split mode: incremental passes need nothing extra -- their main tickets are dealt to the ranks and
    // all-reduced like any others and kernels 2-3, which keep the sums, run replicated; the external-sum
    // planes would need an all-reduce of their own in the first pass, so the cache stays off there.
    P->cache_ext = cfg.cache_external && cfg.mode == HALMA_MODE_FAST && P->max_ext > 0;
    P->incr = cfg.incremental && cfg.mode == HALMA_MODE_FAST;
    if (P->cache_ext || P->incr) CU_TRY(P->d_reuse_int.alloc(3 * std::max<size_t>(nh, 1)));
    if (P->cache_ext) CU_TRY(P->d_phi_ext.alloc(planes * NP));
    if (P->incr) {
        CU_TRY(P->d_phi_keep.alloc(NP));
        CU_TRY(P->d_phi_full.alloc(NP));
        CU_TRY(P->d_rem.alloc(4 * NP));
    }
    L.phi_ext = P->d_phi_ext.p;
    L.ext_ok = P->d_reuse_int.p;
    L.cache_ext = P->cache_ext ? 1 : 0;
    L.phi_keep = P->d_phi_keep.p;
    L.phi_full = P->d_phi_full.p;
    L.rx = P->d_rem.p;
    L.ry = P->incr ? P->d_rem.p + NP : nullptr;
    L.rz = P->incr ? P->d_rem.p + 2 * NP : nullptr;
    L.rm = P->incr ? P->d_rem.p + 3 * NP : nullptr;
    L.incr = P->d_reuse_int.p ? P->d_reuse_int.p + nh : nullptr;
    L.rem_cnt = P->d_reuse_int.p ? P->d_reuse_int.p + 2 * nh : nullptr;
    L.incr_enabled = P->incr ? 1 : 0;
    L.redo_enabled = (P->np || P->incr) ? 1 : 0;

    PotParams &Q = P->pp;
    memset(&Q, 0, sizeof Q);
    for (int b = 0; b < 2; ++b) {
        Q.tx[b] = L.wx[b];
        Q.ty[b] = L.wy[b];
        Q.tz[b] = L.wz[b];
        Q.src[b] = F32Set{L.wx[b], L.wy[b], L.wz[b], L.wm[b]};
    }
    const size_t NE = static_cast<size_t>(P->n_ext_pad);
    Q.src[2] = F32Set{P->d_ext.p, P->d_ext.p + NE, P->d_ext.p + 2 * NE, P->d_ext.p + 3 * NE};
    Q.halo = L.halo;
    Q.cnt = L.cnt;
    Q.order = L.order;
    Q.item_base = L.item_base;
    Q.nsplit = L.nsplit;
    Q.st = L.st;
    Q.phi_part = L.phi_part;
    Q.phi_stride = P->n_pad;
    Q.n_halo = L.n_halo;
    Q.tgt_members = 1;
    Q.rank = cfg.rank;
    Q.n_ranks = cfg.n_ranks;
    Q.halo_redo = P->d_redo.p;
    Q.np_enabled = P->np ? 1 : 0;
    Q.redo_only = 0;
    Q.phi_sym = P->d_phi_sym.p;
    Q.sym_q = L.sym_q;
    Q.sym_enabled = P->sym ? 1 : 0;
    Q.sym_rows = P->sym_rows;
    Q.phi_ext = L.phi_ext;
    Q.ext_ok = L.ext_ok;
    Q.cache_ext = L.cache_ext;
    Q.incr = L.incr;
    Q.rem_cnt = L.rem_cnt;
    Q.incr_enabled = L.incr_enabled;
    Q.widx[0] = L.widx[0];
    Q.widx[1] = L.widx[1];
    Q.phi_keep = L.phi_keep;
    Q.phi_full = L.phi_full;
    Q.src[6] = F32Set{L.rx, L.ry, L.rz, L.rm};
    for (int a = 0; a < 3 && P->np; ++a) {
        Q.ax[a] = L.ax[a];
        Q.src[3 + a] = F32Set{P->sax[a].x, P->sax[a].y, P->sax[a].z, P->sax[a].m};
    }
    // the persistent loop kernel, where one exists for this configuration (fused.cu)
    P->fused_index = cfg.n_ranks == 1 ? fused_kernel_index(cfg.mode, P->variant, P->np, P->sym ? P->sym_rows : 0) : -1;
    if (cfg.use_graph == HALMA_DRIVER_FUSED && P->fused_index < 0)
        return fail(HALMA_ERR_INVALID, "no persistent loop kernel for this plan (split mode or a tuning kernel shape)");
    return HALMA_OK;
}

// Sorted source copies of the predicate-free path: (re)built after an upload, masses
// restored before every run (the loop zeroes the masses of removed members).
static int plan_prepare_sorted(halma_plan *P)
{
    if (!P->np) return HALMA_OK;
    cudaStream_t s = P->stream;
    const size_t NS = static_cast<size_t>(P->n_spad), NT = static_cast<size_t>(P->n_tot);
    if (P->sorted_dirty) {
        CU_TRY(sorted_fill_src_halo(P->d_halo.p, static_cast<int>(P->n_halo), P->d_chunk_halo.p, P->d_chunk_p0.p,
                                    P->n_chunks, P->max_ext, P->n_pad, P->n_tot, P->d_src_halo.p, s));
        const size_t NE = static_cast<size_t>(P->n_ext_pad);
        SortedBuild b;
        b.halo = P->d_halo.p;
        b.src_halo = P->d_src_halo.p;
        b.mem = F32Set{P->lp.wx[0], P->lp.wy[0], P->lp.wz[0], P->lp.wm[0]};
        b.ext = F32Set{P->d_ext.p, P->d_ext.p + NE, P->d_ext.p + 2 * NE, P->d_ext.p + 3 * NE};
        b.n_pad = P->n_pad;
        b.n_tot = P->n_tot;
        b.n_valid = P->n_valid;
        b.n_spad = P->n_spad;
        b.n_halo = static_cast<int32_t>(P->n_halo);
        b.keys_in = P->d_keys.p;
        b.keys_out = P->d_keys.p + NT;
        b.ids_in = P->d_ids.p;
        b.ids_out = P->d_ids.p + NT;
        b.is_member = P->d_ismem.p;
        b.n_selected = P->d_nsel.p;
        b.temp = P->d_sorttemp.p;
        b.temp_bytes = P->sort_temp_bytes;
        for (int a = 0; a < 3; ++a) {
            b.out = &P->sax[a];
            CU_TRY(sorted_build_axis(b, a, s));
        }
        P->sorted_dirty = false;
    } else {
        for (int a = 0; a < 3; ++a)
            CU_TRY(cudaMemcpyAsync(P->sax[a].m, P->sax[a].m0, NS * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return HALMA_OK;
}

extern "C" int halma_plan_create(const halma_unbind_config *cfg, int64_t n_halo, const int64_t *offsets,
                                 const int64_t *const *ext_offsets, halma_plan **out)
{
    if (!cfg || !offsets || !out) return fail(HALMA_ERR_INVALID, "null pointer");
    if (cfg->struct_size != static_cast<int32_t>(sizeof(halma_unbind_config)))
        return fail(HALMA_ERR_INVALID, "halma_unbind_config size mismatch (ABI)");
    if (int rc = check_mode(cfg->mode)) return rc;
    if (n_halo < 0 || n_halo > 0x7ffffff0ll) return fail(HALMA_ERR_INVALID, "bad n_halo");
    if (cfg->n_groups < 0 || cfg->n_groups > HALMA_MAX_GROUPS) return fail(HALMA_ERR_INVALID, "bad n_groups");
    if (cfg->n_pre < 0 || cfg->n_pre > cfg->n_groups) return fail(HALMA_ERR_INVALID, "bad n_pre");
    if (cfg->max_iter < 1 || cfg->max_iter > 4096) return fail(HALMA_ERR_INVALID, "max_iter must be in 1..4096");
    if (cfg->n_ranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->n_ranks) return fail(HALMA_ERR_INVALID, "bad rank");
    if (cfg->n_ranks > 1 && n_halo != 1) return fail(HALMA_ERR_INVALID, "split mode shares exactly one halo");
    DeviceCtx *c;
    if (int rc = get_ctx(cfg->device, &c)) return rc;
    halma_plan *P = new halma_plan();
    P->cfg = *cfg;
    P->ctx = c;
    P->n_halo = n_halo;
    cudaError_t e = cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete P;
        return fail(HALMA_ERR_CUDA, cudaGetErrorString(e));
    }
    int rc = plan_build(P, offsets, ext_offsets);
    if (rc) {
        delete P;
        return rc;
    }
    *out = P;
    return HALMA_OK;
}

extern "C" void halma_plan_destroy(halma_plan *plan)
{
    if (!plan) return;
    cudaSetDevice(plan->cfg.device);
    cudaStreamSynchronize(plan->stream);
    delete plan;
}

extern "C" int halma_plan_upload_members(halma_plan *P, const double *x, const double *y, const double *z,
                                         const double *vx, const double *vy, const double *vz, const double *mass)
{
    if (!P) return fail(HALMA_ERR_INVALID, "null plan");
    const double *src[7] = {x, y, z, vx, vy, vz, mass};
    CU_TRY(cudaSetDevice(P->cfg.device));
    const size_t NU = static_cast<size_t>(P->n_user);
    for (int k = 0; k < 7; ++k) {
        if (NU && !src[k]) return fail(HALMA_ERR_INVALID, "null member array");
        if (NU) CU_TRY(cudaMemcpyAsync(P->d_in.p + k * NU, src[k], NU * 8, cudaMemcpyDefault, P->stream));
    }
    P->members_up = true;
    P->sorted_dirty = true;
    return HALMA_OK;
}

extern "C" int halma_plan_upload_group(halma_plan *P, int group, const double *mass, const double *x, const double *y,
                                       const double *z)
{
    if (!P) return fail(HALMA_ERR_INVALID, "null plan");
    if (group < 0 || group >= P->cfg.n_groups) return fail(HALMA_ERR_INVALID, "bad group index");
    CU_TRY(cudaSetDevice(P->cfg.device));
    const size_t n = static_cast<size_t>(P->ext_offsets[group][P->n_halo]);
    if (n) {
        if (!mass || !x || !y || !z) return fail(HALMA_ERR_INVALID, "null group array");
        const double *src[4] = {mass, x, y, z};
        // the staging buffer is reused by the next group: stream order keeps this safe
        for (int k = 0; k < 4; ++k)
            CU_TRY(cudaMemcpyAsync(P->d_stage.p + k * n, src[k], n * 8, cudaMemcpyDefault, P->stream));
        const size_t NE = static_cast<size_t>(P->n_ext_pad);
        float *e = P->d_ext.p;
        CU_TRY(launch_pack_group(P->d_halo.p, static_cast<int>(P->n_halo), P->group_seg[group], P->group_max[group],
                                 P->d_ext_off[group].p, P->d_stage.p, P->d_stage.p + n, P->d_stage.p + 2 * n,
                                 P->d_stage.p + 3 * n, e + 3 * NE, e, e + NE, e + 2 * NE, P->stream));
    }
    P->group_up[group] = true;
    P->sorted_dirty = true;
    return HALMA_OK;
}

extern "C" int halma_plan_upload_temp(halma_plan *P, const double *temp, double cold_T)
{
    if (!P) return fail(HALMA_ERR_INVALID, "null plan");
    CU_TRY(cudaSetDevice(P->cfg.device));
    if (P->n_user) {
        if (!temp) return fail(HALMA_ERR_INVALID, "null temperature array");
        CU_TRY(cudaMemcpyAsync(P->d_temp.p, temp, P->n_user * sizeof(double), cudaMemcpyDefault, P->stream));
    }
    P->temp_up = true;
    P->cold_T = cold_T;
    P->lp.temp = P->d_temp.p;
    P->lp.cold_T = cold_T;
    P->graph_dirty = true;          // kernel parameters are baked into the graph
    return HALMA_OK;
}

extern "C" int halma_plan_set_vb(halma_plan *P, const double *vb)
{
    if (!P || !vb) return fail(HALMA_ERR_INVALID, "null pointer");
    CU_TRY(cudaSetDevice(P->cfg.device));
    if (P->n_halo)
        CU_TRY(cudaMemcpyAsync(P->d_vb_user.p, vb, 3 * P->n_halo * 8, cudaMemcpyHostToDevice, P->stream));
    P->vb_up = true;
    return HALMA_OK;
}

// Blocks until everything enqueued on the plan's stream (uploads, a run) has completed.
extern "C" int halma_plan_sync(halma_plan *P)
{
    if (!P) return fail(HALMA_ERR_INVALID, "null plan");
    CU_TRY(cudaSetDevice(P->cfg.device));
    CU_TRY(cudaStreamSynchronize(P->stream));
    return HALMA_OK;
}

extern "C" int halma_plan_join(halma_plan *P, const void *unique_id_128)
{
    if (!P || !unique_id_128) return fail(HALMA_ERR_INVALID, "null pointer");
    if (P->cfg.n_ranks < 2) return fail(HALMA_ERR_STATE, "plan was not created in split mode");
    if (int rc = load_nccl()) return rc;
    CU_TRY(cudaSetDevice(P->cfg.device));
    UID128 id;
    memcpy(&id, unique_id_128, sizeof id);
    int e = g_nccl.CommInitRank(&P->comm, P->cfg.n_ranks, id, P->cfg.rank);
    if (e) return nccl_fail(e, "ncclCommInitRank");
    P->comm_owned = true;
    return HALMA_OK;
}

struct halma_comm {
    void *nccl = nullptr;
    int device = 0, rank = 0, n_ranks = 1;
};

extern "C" int halma_comm_create(int device, int rank, int n_ranks, const void *unique_id_128, halma_comm **out)
{
    if (!unique_id_128 || !out) return fail(HALMA_ERR_INVALID, "null pointer");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(HALMA_ERR_INVALID, "bad rank");
    if (int rc = load_nccl()) return rc;
    DeviceCtx *c;
    if (int rc = get_ctx(device, &c)) return rc;
    UID128 id;
    memcpy(&id, unique_id_128, sizeof id);
    halma_comm *C = new halma_comm();
    C->device = device;
    C->rank = rank;
    C->n_ranks = n_ranks;
    int e = g_nccl.CommInitRank(&C->nccl, n_ranks, id, rank);
    if (e) {
        delete C;
        return nccl_fail(e, "ncclCommInitRank");
    }
    *out = C;
    return HALMA_OK;
}

extern "C" void halma_comm_destroy(halma_comm *comm)
{
    if (!comm) return;
    if (comm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->nccl);
    delete comm;
}

extern "C" int halma_plan_use_comm(halma_plan *P, halma_comm *comm)
{
    if (!P || !comm) return fail(HALMA_ERR_INVALID, "null pointer");
    if (P->cfg.n_ranks < 2) return fail(HALMA_ERR_STATE, "plan was not created in split mode");
    if (comm->n_ranks != P->cfg.n_ranks || comm->rank != P->cfg.rank || comm->device != P->cfg.device)
        return fail(HALMA_ERR_INVALID, "communicator does not match the plan's rank / n_ranks / device");
    if (P->comm && P->comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(P->comm);
    P->comm = comm->nccl;
    P->comm_owned = false;
    return HALMA_OK;
}

// Kernels of one pass, without events or host copies (graph driver).
static int enqueue_pass_kernels(halma_plan *P, const LoopParams &lp)
{
    cudaStream_t s = P->stream;
    const int sm = P->ctx->sm_count;
    const int grid = sm * P->ctx->blocks(P->cfg.mode, P->variant);
    CU_TRY(potential_launch(P->pp, P->cfg.mode, P->variant, grid, s));
    if (P->lp.redo_enabled) {
        PotParams redo = P->pp;
        redo.redo_only = 1;
        CU_TRY(potential_launch(redo, P->cfg.mode, P->variant, grid, s));
    }
    CU_TRY(launch_energy_flag(lp, sm, s));
    CU_TRY(launch_compact(lp, sm, s));
    CU_TRY(launch_schedule(lp, 0, sm, s));
    return HALMA_OK;
}

// Device-resident loop as a CUDA graph: a WHILE conditional node whose body is one pass;
// k_schedule sets the condition on the device (any halo still active), so one graph launch
// runs the whole unbinding without the host.
static int plan_build_graph(halma_plan *P)
{
    if (P->gexec) {
        cudaGraphExecDestroy(P->gexec);
        P->gexec = nullptr;
    }
    if (P->graph) {
        cudaGraphDestroy(P->graph);
        P->graph = nullptr;
    }
    CU_TRY(cudaGraphCreate(&P->graph, 0));
    cudaGraphConditionalHandle handle;
    CU_TRY(cudaGraphConditionalHandleCreate(&handle, P->graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams cp = {};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t node;
    CU_TRY(cudaGraphAddNode(&node, P->graph, nullptr, 0, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    LoopParams lp = P->lp;
    lp.cond_handle = handle;
    CU_TRY(cudaStreamBeginCaptureToGraph(P->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_pass_kernels(P, lp);
    cudaGraph_t captured = nullptr;
    cudaError_t e = cudaStreamEndCapture(P->stream, &captured);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(HALMA_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
    CU_TRY(cudaGraphInstantiate(&P->gexec, P->graph, 0));
    P->graph_dirty = false;
    return HALMA_OK;
}

// One pass of the enqueue-ahead driver (the only driver of split mode).  Split mode: the ranks evaluate
// disjoint target groups; what the replicated kernels 2-3 need of the others' results is exchanged by
// all-reduces on the plan's stream, in which every element is written by one rank and the others add
// exact zeros (the two-sided sums: exact additions of quantised addends), so the outcome is bit-identical
// to the single-GPU run.  *comm_bytes counts what went through the collectives.
static int enqueue_pass(halma_plan *P, int pass, int64_t *comm_bytes)
{
    cudaStream_t s = P->stream;
    const int sm = P->ctx->sm_count;
    const int grid = sm * P->ctx->blocks(P->cfg.mode, P->variant);
    const bool split = P->cfg.n_ranks > 1;
    const size_t NP = static_cast<size_t>(P->n_pad);
    CU_TRY(cudaEventRecord(P->ev[3 * pass], s));
    if (P->np && split)      // split mode: every correction entry is written by one rank only
        CU_TRY(cudaMemsetAsync(P->d_corr.p, 0, 3 * NP * sizeof(double), s));
    CU_TRY(potential_launch(P->pp, P->cfg.mode, P->variant, grid, s));
    if (P->lp.redo_enabled) {
        // haloes whose predicate-free sums came out non-finite (or whose incremental pass removed too much of
        // a member's potential) are recomputed with the predicate
        if (split) {
            // split mode: every rank must take the same decision
            CU_TRY(cudaEventRecord(P->cev[4 * pass], s));
            int e = g_nccl.AllReduce(P->d_redo.p, P->d_redo.p, static_cast<size_t>(P->n_halo), /*ncclInt32*/ 2,
                                     /*ncclMax*/ 2, P->comm, s);
            if (e) return nccl_fail(e, "ncclAllReduce(flags)");
            CU_TRY(launch_sync_redo(P->lp, s));
            CU_TRY(cudaEventRecord(P->cev[4 * pass + 1], s));
            *comm_bytes += static_cast<int64_t>(P->n_halo) * 4;
        }
        PotParams redo = P->pp;
        redo.redo_only = 1;
        CU_TRY(potential_launch(redo, P->cfg.mode, P->variant, grid, s));
    }
    CU_TRY(cudaEventRecord(P->ev[3 * pass + 1], s));
    if (split) {
        CU_TRY(launch_fold_partials(P->lp, sm, s));
        CU_TRY(cudaEventRecord(P->cev[4 * pass + 2], s));
        // one group = one fused NCCL launch for everything kernels 2-3 need
        if (g_nccl.GroupStart) g_nccl.GroupStart();
        int e = g_nccl.AllReduce(P->d_phi.p, P->d_phi.p, NP, /*ncclFloat64*/ 8, /*ncclSum*/ 0, P->comm, s);
        *comm_bytes += NP * 8;
        if (!e && P->np) {
            e = g_nccl.AllReduce(P->d_corr.p, P->d_corr.p, 3 * NP, 8, 0, P->comm, s);
            *comm_bytes += 3 * NP * 8;
        }
        if (!e && P->sym) {
            // every rank summed the two-sided terms of its row tiles; the addends are multiples of the
            // halo's quantum and the totals stay inside the exact window, so this sum is exact too
            e = g_nccl.AllReduce(P->d_phi_sym.p, P->d_phi_sym.p, NP, 8, 0, P->comm, s);
            *comm_bytes += NP * 8;
        }
        if (!e && P->cache_ext && pass == 0) {
            // external-sum cache: the sums over the fixed external sources are exchanged once per run
            e = g_nccl.AllReduce(P->d_phi_ext.p, P->d_phi_ext.p, NP, 8, 0, P->comm, s);
            *comm_bytes += NP * 8;
        }
        if (g_nccl.GroupEnd) {
            const int e2 = g_nccl.GroupEnd();
            if (!e) e = e2;
        }
        if (e) return nccl_fail(e, "ncclAllReduce");
        CU_TRY(cudaEventRecord(P->cev[4 * pass + 3], s));
        CU_TRY(launch_set_nsplit_one(P->lp, s));
    }
    CU_TRY(launch_energy_flag(P->lp, sm, s));
    CU_TRY(launch_compact(P->lp, sm, s));
    CU_TRY(launch_schedule(P->lp, 0, sm, s));
    CU_TRY(cudaMemcpyAsync(&P->h_flags[pass], &P->d_st.p->any_active, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaEventRecord(P->ev[3 * pass + 2], s));
    return HALMA_OK;
}

static int pick_driver(const halma_plan *P)
{
    int d = P->cfg.use_graph;
    if (d == HALMA_DRIVER_AUTO) {
        if (const char *e = getenv("HALMA_DRIVER")) {      // fused | enqueue | graph: tuning and A/B tests
            if (!strcmp(e, "enqueue")) d = HALMA_DRIVER_ENQUEUE;
            else if (!strcmp(e, "graph")) d = HALMA_DRIVER_GRAPH;
        }
    }
    if (P->cfg.n_ranks > 1) return HALMA_DRIVER_ENQUEUE;
    if (d == HALMA_DRIVER_AUTO || d == HALMA_DRIVER_FUSED) return P->fused_index >= 0 ? HALMA_DRIVER_FUSED : HALMA_DRIVER_ENQUEUE;
    return d;
}

extern "C" int halma_plan_run(halma_plan *P, halma_run_stats *stats)
{
    if (!P) return fail(HALMA_ERR_INVALID, "null plan");
    if (!P->members_up) return fail(HALMA_ERR_STATE, "members were not uploaded");
    for (int g = 0; g < P->cfg.n_groups; ++g)
        if (!P->group_up[g]) return fail(HALMA_ERR_STATE, "an external group was not uploaded");
    if (P->cfg.vb_fixed && !P->vb_up) return fail(HALMA_ERR_STATE, "vb_fixed = 1 but halma_plan_set_vb was not called");
    if (P->cfg.n_ranks > 1 && !P->comm) return fail(HALMA_ERR_STATE, "split mode needs halma_plan_join first");
    CU_TRY(cudaSetDevice(P->cfg.device));
    cudaStream_t s = P->stream;
    const int sm = P->ctx->sm_count;
    const int nh = static_cast<int>(P->n_halo);
    const int driver = pick_driver(P);
    int launches = 0, pot_launches = 0, passes = 0;
    int64_t comm_bytes = 0;
    LoopState hst;
    memset(&hst, 0, sizeof hst);

    if (driver != HALMA_DRIVER_FUSED && P->ev.empty()) {
        if (!P->h_flags)
            CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&P->h_flags), sizeof(int32_t) * (P->cfg.max_iter + 1)));
        P->ev.resize(3 * static_cast<size_t>(P->cfg.max_iter));
        for (auto &e : P->ev) CU_TRY(cudaEventCreate(&e));
        if (P->cfg.n_ranks > 1) {
            P->cev.resize(4 * static_cast<size_t>(P->cfg.max_iter));
            for (auto &e : P->cev) CU_TRY(cudaEventCreate(&e));
        }
    }
    CU_TRY(cudaEventRecord(P->ev_start, s));
    // every member is evaluated by the first pass, which writes its mask / be / E: no need to clear them
    CU_TRY(cudaMemsetAsync(P->d_st.p, 0, sizeof(LoopState), s));
    CU_TRY(cudaMemsetAsync(P->d_hint.p, 0, P->d_hint.n * sizeof(int32_t), s));
    if (P->sym && nh) {
        CU_TRY(cudaMemsetAsync(P->d_phi_sym.p, 0, static_cast<size_t>(P->n_pad) * sizeof(double), s));
        CU_TRY(cudaMemsetAsync(P->d_range.p, 0x7f, 3 * nh * sizeof(int32_t), s));
        CU_TRY(cudaMemsetAsync(P->d_range.p + 3 * nh, 0x80, 3 * nh * sizeof(int32_t), s));
    }
    if (P->cfg.vb_fixed && nh)
        CU_TRY(cudaMemcpyAsync(P->lp.hvb, P->d_vb_user.p, 3 * nh * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (nh > 0 && driver == HALMA_DRIVER_FUSED) {
        // the whole loop in one persistent kernel; the pack runs inside it unless the sorted copies have to
        // be (re)built from the packed members first
        int do_pack = 1;
        if (P->np && P->sorted_dirty) {
            CU_TRY(launch_pack_members(P->lp, sm, s));
            ++launches;
            if (int rc = plan_prepare_sorted(P)) return rc;
            do_pack = 0;
            if (P->sym) {      // the pack's coordinate ranges are kept; nothing else to reset
            }
        }
        CU_TRY(fused_launch(P->fused_index, P->pp, P->lp, do_pack, sm, s));
        ++launches;
        // the loop state comes back through pinned memory, so that the copy and the stop event are queued behind the
        // kernel without the host in between
        static_assert(sizeof(LoopState) <= kPinnedSlotBytes, "pinned slot");
        void *pin = nullptr;
        CU_TRY(pinned_borrow(P->ctx, &pin));
        cudaError_t ce = cudaMemcpyAsync(pin, P->d_st.p, sizeof hst, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaEventRecord(P->ev_stop, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        if (ce == cudaSuccess) memcpy(&hst, pin, sizeof hst);
        pinned_return(P->ctx, pin);
        CU_TRY(ce);
        passes = pot_launches = hst.pass;
    } else if (nh > 0) {
        CU_TRY(launch_pack_members(P->lp, sm, s));
        if (int rc = plan_prepare_sorted(P)) return rc;
        CU_TRY(launch_halo_decide_init(P->lp, sm, s));
        CU_TRY(launch_schedule(P->lp, 1, sm, s));
        launches += 3;
        // Enqueue passes ahead of the device; a pinned flag written at the end of each pass
        // tells the host when the loop has converged.  The device never waits for the host:
        // pass k+kAhead is queued before the host looks at the flag of pass k, and a pass
        // queued after convergence returns at once (any_active == 0).
        constexpr int kAhead = 2;
        const int max_iter = P->cfg.max_iter;
        int queued = 0;
        bool done = false;
        if (driver == HALMA_DRIVER_GRAPH) {
            // graph driver: every pass inside one launch, convergence decided on the device
            if (P->graph_dirty)
                if (int rc = plan_build_graph(P)) return rc;
            CU_TRY(cudaGraphLaunch(P->gexec, s));
            CU_TRY(cudaMemcpyAsync(&hst, P->d_st.p, sizeof hst, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            passes = hst.pass;
            queued = passes;        // kernels of the early-out pass that ends the loop are not counted
            done = true;
        }
        while (!done) {
            while (queued < max_iter && queued < passes + kAhead) {
                if (int rc = enqueue_pass(P, queued, &comm_bytes)) return rc;
                ++queued;
            }
            if (passes >= queued) break;
            CU_TRY(cudaEventSynchronize(P->ev[3 * passes + 2]));
            const int still = P->h_flags[passes];
            ++passes;
            if (!still) done = true;
        }
        launches += queued * (4 + (P->cfg.n_ranks > 1 ? 2 : 0) + (P->lp.redo_enabled ? (P->cfg.n_ranks > 1 ? 2 : 1) : 0));
        pot_launches = queued;
        CU_TRY(launch_finalize(P->lp, sm, s));
        ++launches;
        CU_TRY(cudaEventRecord(P->ev_stop, s));
        CU_TRY(cudaStreamSynchronize(s));
    } else {
        CU_TRY(cudaEventRecord(P->ev_stop, s));
        CU_TRY(cudaStreamSynchronize(s));
    }
    P->ran = true;
    P->driver_ran = driver;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, P->ev_start, P->ev_stop));
        stats->total_ms = ms;
        double pot = 0.0, comm = 0.0;
        if (driver == HALMA_DRIVER_FUSED) {
            pot = hst.pot_ns * 1e-6;
            stats->loop_ms = hst.loop_ns * 1e-6;
            for (int k = 0; k < 5; ++k) stats->phase_ms[k] = hst.phase_ns[k] * 1e-6;
        } else if (driver == HALMA_DRIVER_ENQUEUE) {      // per-launch events exist only with the enqueue-ahead driver
            for (int k = 0; k < pot_launches; ++k) {
                CU_TRY(cudaEventElapsedTime(&ms, P->ev[3 * k], P->ev[3 * k + 1]));
                pot += ms;
                if (P->cfg.n_ranks > 1) {
                    if (P->lp.redo_enabled) {
                        CU_TRY(cudaEventElapsedTime(&ms, P->cev[4 * k], P->cev[4 * k + 1]));
                        comm += ms;
                        pot -= ms;      // the flag exchange sits between the two potential launches
                    }
                    CU_TRY(cudaEventElapsedTime(&ms, P->cev[4 * k + 2], P->cev[4 * k + 3]));
                    comm += ms;
                }
            }
        }
        stats->potential_ms = pot;
        stats->comm_ms = comm;
        stats->comm_bytes = comm_bytes;
        stats->potential_launches = pot_launches;
        stats->launches = launches;
        stats->passes = passes;
        stats->driver = driver;
        if (nh && driver == HALMA_DRIVER_FUSED) {
            stats->pairs = static_cast<int64_t>(hst.pairs_total);
            stats->evaluations = static_cast<int64_t>(hst.evals_total);
        } else if (nh) {
            std::vector<unsigned long long> pr(nh);
            CU_TRY(cudaMemcpy(pr.data(), P->d_pairs.p, nh * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            unsigned long long t = 0;
            for (auto v : pr) t += v;
            stats->pairs = static_cast<int64_t>(t);
            CU_TRY(cudaMemcpy(pr.data(), P->d_evals.p, nh * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            t = 0;
            for (auto v : pr) t += v;
            stats->evaluations = static_cast<int64_t>(t);
        }
    }
    return HALMA_OK;
}

// Tuning aid: per-pass phase times of the last FUSED run (potential, energy + compaction, tables; ns), first 16 passes.
extern "C" int halma_plan_debug_pass_ns(halma_plan *P, uint32_t *out48)
{
    if (!P || !out48) return fail(HALMA_ERR_INVALID, "null pointer");
    CU_TRY(cudaSetDevice(P->cfg.device));
    LoopState hst;
    CU_TRY(cudaMemcpy(&hst, P->d_st.p, sizeof hst, cudaMemcpyDeviceToHost));
    memcpy(out48, hst.pass_ns, sizeof hst.pass_ns);
    // tuning aid: how long the slowest warp spent in the energy phase (overwrites nothing the caller relies on: printed)
    if (getenv("HALMA_DEBUG_PHASES"))
        for (int k = 0; k < 16 && hst.dbg_e_end[k]; ++k)
            fprintf(stderr, "pass %d: energy phase (slowest warp) %.1f us of energy+compaction %.1f us\n", k,
                    (hst.dbg_e_end[k] - hst.dbg_phase_start[k]) * 1e-3, hst.pass_ns[k][1] * 1e-3);
    if (getenv("HALMA_DEBUG_PHASES") && P->fused_index >= 0) {
        const double warps = static_cast<double>(P->ctx->sm_count) * fused_blocks_per_sm(P->fused_index) * (kPotentialBlock / 32);
        for (int k = 0; k < 16 && hst.dbg_pot_busy[k]; ++k)
            fprintf(stderr, "pass %d: potential phase %.1f us, %u tickets, warps busy %.3f of it on average\n", k,
                    hst.pass_ns[k][0] * 1e-3, hst.dbg_items[k],
                    hst.dbg_pot_busy[k] / (warps * (hst.pass_ns[k][0] > 0 ? hst.pass_ns[k][0] : 1)));
    }
    return HALMA_OK;
}

extern "C" int halma_plan_download(halma_plan *P, uint8_t *mask, float *be, double *energy, int32_t *idx,
                                   halma_halo_result *halos)
{
    if (!P) return fail(HALMA_ERR_INVALID, "null plan");
    if (!P->ran) return fail(HALMA_ERR_STATE, "halma_plan_run has not been called");
    CU_TRY(cudaSetDevice(P->cfg.device));
    cudaStream_t s = P->stream;
    const size_t NU = static_cast<size_t>(P->n_user);
    const size_t nh = static_cast<size_t>(P->n_halo);
    if (NU) {
        if (mask) CU_TRY(cudaMemcpyAsync(mask, P->d_mask.p, NU, cudaMemcpyDeviceToHost, s));
        if (be) CU_TRY(cudaMemcpyAsync(be, P->d_be.p, NU * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (energy) CU_TRY(cudaMemcpyAsync(energy, P->d_E.p, NU * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (idx) CU_TRY(cudaMemcpyAsync(idx, P->d_idx.p, NU * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    if (halos && nh) {
        std::vector<int32_t> hi(P->d_hint.n);
        std::vector<double> hd(P->d_hdbl.n);
        std::vector<unsigned long long> pr(nh);
        std::vector<double> rps(4 * nh);
        std::vector<int32_t> best(nh);
        CU_TRY(cudaMemcpyAsync(rps.data(), P->d_hrps.p, rps.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(best.data(), P->d_hbest.p, nh * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(hi.data(), P->d_hint.p, hi.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(hd.data(), P->d_hdbl.p, hd.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(pr.data(), P->d_pairs.p, nh * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        for (size_t h = 0; h < nh; ++h) {
            halma_halo_result &r = halos[h];
            r.n_bound = hi[h];
            r.n_iter = hi[2 * nh + h];
            r.converged = hi[7 * nh + h];
            r.mass = hd[h];
            for (int k = 0; k < 3; ++k) {
                r.vb[k] = hd[nh + 3 * h + k];
                r.com[k] = hd[7 * nh + 3 * h + k];
            }
            r.pairs = static_cast<int64_t>(pr[h]);
            r.most_bound = best[h];
            r.mass_initial = rps[4 * h];
            r.cold_bound_mass = P->temp_up ? rps[4 * h + 1] : 0.0;
            r.unbound_cold_mass = P->temp_up ? rps[4 * h + 2] : 0.0;
            r.unbound_hot_mass = P->temp_up ? rps[4 * h + 3] : 0.0;
        }
    }
    CU_TRY(cudaStreamSynchronize(s));
    return HALMA_OK;
}

extern "C" int halma_unbind_catalogue(const halma_unbind_config *cfg, int64_t n_halo, const int64_t *offsets,
                                      const double *x, const double *y, const double *z, const double *vx,
                                      const double *vy, const double *vz, const double *mass,
                                      const int64_t *const *group_offsets, const double *const *group_mass,
                                      const double *const *group_x, const double *const *group_y,
                                      const double *const *group_z, const double *vb, const double *temp,
                                      double cold_T, uint8_t *mask, float *be, double *energy, int32_t *idx,
                                      halma_halo_result *halos, halma_run_stats *stats)
{
    if (!cfg) return fail(HALMA_ERR_INVALID, "null config");
    if (cfg->n_groups > 0 && (!group_offsets || !group_mass || !group_x || !group_y || !group_z))
        return fail(HALMA_ERR_INVALID, "null group table");
    halma_plan *P = nullptr;
    int rc = halma_plan_create(cfg, n_halo, offsets, group_offsets, &P);
    if (rc) return rc;
    rc = halma_plan_upload_members(P, x, y, z, vx, vy, vz, mass);
    for (int g = 0; !rc && g < cfg->n_groups; ++g)
        rc = halma_plan_upload_group(P, g, group_mass[g], group_x[g], group_y[g], group_z[g]);
    if (!rc && cfg->vb_fixed) rc = vb ? halma_plan_set_vb(P, vb) : fail(HALMA_ERR_INVALID, "vb_fixed = 1 needs vb");
    if (!rc && temp) rc = halma_plan_upload_temp(P, temp, cold_T);
    if (!rc) rc = halma_plan_run(P, stats);
    if (!rc) rc = halma_plan_download(P, mask, be, energy, idx, halos);
    halma_plan_destroy(P);
    return rc;
}

extern "C" int halma_unbind_halo(const halma_unbind_config *cfg, int64_t n, const double *x, const double *y,
                                 const double *z, const double *vx, const double *vy, const double *vz,
                                 const double *mass, const int64_t *group_n, const double *const *group_mass,
                                 const double *const *group_x, const double *const *group_y,
                                 const double *const *group_z, const double *vb, uint8_t *mask, float *be,
                                 double *energy, int32_t *idx, halma_halo_result *result, halma_run_stats *stats)
{
    if (!cfg) return fail(HALMA_ERR_INVALID, "null config");
    if (n < 0) return fail(HALMA_ERR_INVALID, "negative size");
    if (cfg->n_groups > 0 && !group_n) return fail(HALMA_ERR_INVALID, "null group table");
    if (cfg->n_groups < 0 || cfg->n_groups > HALMA_MAX_GROUPS) return fail(HALMA_ERR_INVALID, "bad n_groups");
    const int64_t offsets[2] = {0, n};
    int64_t eo[HALMA_MAX_GROUPS][2];
    const int64_t *eop[HALMA_MAX_GROUPS];
    for (int g = 0; g < cfg->n_groups; ++g) {
        eo[g][0] = 0;
        eo[g][1] = group_n[g];
        eop[g] = eo[g];
    }
    return halma_unbind_catalogue(cfg, 1, offsets, x, y, z, vx, vy, vz, mass, eop, group_mass, group_x, group_y, group_z,
                                  vb, nullptr, 0.0, mask, be, energy, idx, result, stats);
}

// ---------------------------------------------------------------------------------------
// brute_force_binding_energy through a one-pass plan: large FAST calls of halma_potential_f32 whose targets
// are also sources (from HALMA_POT_PLAN_MIN_PAIRS pairs on, default 1e10; 0 = never).
//
// The direct path (potential_dev) runs the predicated kernel.  For a large call it pays to build the sorted
// source copies once and use the predicate-free kernel + correction tickets of the plans, with the symmetric
// self-term on top.  The reference's callers come in three shapes:
//   * the targets ARE the sources (gas-gas of RPS, halo_gas.py:306-328; stars-stars of most_bound_particle,
//     :602-622): 84.8 -> 47.1 ms for 5e5 gas cells, host arrays in and out (profiles/bench_f2py_call_r01.jsonl);
//   * the targets are a block of the sources (escape_velocity_unbinding_fortran, halo_properties.py:333-339:
//     sources = concat(gas, stars, DM), targets = the stars): 49.9 -> 37.1 ms for 2e5 stars in 7e5 sources;
//   * the targets are other particles (DM -> gas, stars -> gas): these stay on the direct path -- as massless
//     members left out of the source list they gained 9 %, and targets that coincide with sources (a sampled
//     subset of the gas as sources, halo_gas.py:307-321) would send every such call through the fallback.
// The block becomes the plan's members, the sources before / after it external groups.  One pass, fixed zero
// bulk velocity: out_be is the plan's `be`.
// ---------------------------------------------------------------------------------------
namespace {
__global__ void k_widen_members(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                                const float *__restrict__ m, int64_t n, double *__restrict__ out7)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        out7[i] = static_cast<double>(x[i]);
        out7[n + i] = static_cast<double>(y[i]);
        out7[2 * n + i] = static_cast<double>(z[i]);
        out7[3 * n + i] = 0.0;
        out7[4 * n + i] = 0.0;
        out7[5 * n + i] = 0.0;
        out7[6 * n + i] = m ? static_cast<double>(m[i]) : 0.0;
    }
}

// Offset of the block of sources whose coordinates equal the targets' bit for bit, or -1.
int64_t find_target_block(const float *sx, const float *sy, const float *sz, int64_t n_src, const float *tx,
                          const float *ty, const float *tz, int64_t n_tgt)
{
    if (n_tgt < 1 || n_tgt > n_src) return -1;
    const size_t tb = static_cast<size_t>(n_tgt) * sizeof(float);
    uint32_t first;
    memcpy(&first, tx, sizeof first);
    const uint32_t *bits = reinterpret_cast<const uint32_t *>(sx);
    int tries = 0;
    for (int64_t off = 0; off + n_tgt <= n_src; ++off) {
        if (bits[off] != first) continue;
        if (memcmp(sx + off, tx, tb) == 0 && memcmp(sy + off, ty, tb) == 0 && memcmp(sz + off, tz, tb) == 0) return off;
        if (++tries >= 64) break;       // many sources share the first target's x (a lattice): give up
    }
    return -1;
}
}  // namespace

extern "C" int64_t halma_find_target_block(const float *src_x, const float *src_y, const float *src_z, int64_t n_src,
                                           const float *tgt_x, const float *tgt_y, const float *tgt_z, int64_t n_tgt)
{
    if (n_src < 0 || n_tgt < 0) return -1;
    if (n_tgt > 0 && (!src_x || !src_y || !src_z || !tgt_x || !tgt_y || !tgt_z)) return -1;
    return find_target_block(src_x, src_y, src_z, n_src, tgt_x, tgt_y, tgt_z, n_tgt);
}

static bool potential_plan_worthwhile(int64_t n_src, int64_t n_tgt)
{
    const char *e = getenv("HALMA_POT_PLAN_MIN_PAIRS");          // <= 0: never
    const double min_pairs = e ? atof(e) : 1e10;
    if (min_pairs <= 0.0 || n_src < 1 || n_tgt < 1) return false;
    return static_cast<double>(n_src) * static_cast<double>(n_tgt) >= min_pairs;
}

static int potential_via_plan(int device, const float *sm, const float *sx, const float *sy, const float *sz,
                              int64_t n_src, const float *tx, const float *ty, const float *tz, int64_t n_tgt,
                              float *out_be)
{
    const size_t tb = static_cast<size_t>(n_tgt) * sizeof(float);
    const int64_t off = find_target_block(sx, sy, sz, n_src, tx, ty, tz, n_tgt);
    // cross call (DM -> gas, stars -> gas, gas -> stars ...): the targets become massless members that are not
    // sources, all sources one external group; predicate-free kernel + correction tickets, no symmetric term.
    // A target that coincides with a source (subsampled self calls) sends the call to the predicated kernel
    // through the usual fallback, so the result is always the reference's.
    const bool cross = off < 0;
    const char *ce = getenv("HALMA_POT_PLAN_CROSS");
    if (cross && ce && atoi(ce) == 0) return kPlanPathDeclined;
    // external groups: the sources before and after the block (all of them for a cross call)
    int64_t g_begin[2], g_count[2];
    int n_groups = 0, n_pre = 0;
    if (cross) {
        g_begin[0] = 0;
        g_count[0] = n_src;
        n_groups = n_pre = 1;
    } else {
        if (off > 0) {
            g_begin[n_groups] = 0;
            g_count[n_groups++] = off;
            n_pre = 1;
        }
        if (off + n_tgt < n_src) {
            g_begin[n_groups] = off + n_tgt;
            g_count[n_groups++] = n_src - off - n_tgt;
        }
    }
    halma_unbind_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = static_cast<int32_t>(sizeof cfg);
    cfg.device = device;
    cfg.mode = HALMA_MODE_FAST;
    cfg.n_groups = n_groups;
    cfg.n_pre = n_pre;
    cfg.vb_fixed = 1;
    cfg.max_iter = 1;
    cfg.G = 1.0;
    cfg.kappa = 1.0;
    cfg.n_ranks = 1;
    {
        const char *e = getenv("HALMA_SYMMETRIC");      // the same switch as the plans of the Python layer
        cfg.symmetric = (!cross && !(e && atoi(e) == 0)) ? 1 : 0;
    }
    const int64_t offsets[2] = {0, n_tgt};
    const int64_t eo[2][2] = {{0, n_groups > 0 ? g_count[0] : 0}, {0, n_groups > 1 ? g_count[1] : 0}};
    const int64_t *eop[2] = {eo[0], eo[1]};
    halma_plan *P = nullptr;
    int rc = halma_plan_create(&cfg, 1, offsets, n_groups ? eop : nullptr, &P);
    if (rc == HALMA_ERR_TOO_LARGE) return kPlanPathDeclined;      // the direct path takes up to 2^31-16 of each
    if (rc) return rc;
    struct Guard {
        halma_plan *p;
        ~Guard() { halma_plan_destroy(p); }
    } guard{P};
    if (!P->np) return kPlanPathDeclined;          // e.g. too many particles for the sorted copies
    if (cross) {
        P->lp.targets_only = 1;
        P->pp.targets_only = 1;
    }
    cudaStream_t s = P->stream;
    g_alloc_stream = s;
    {
        // targets -> the plan's float64 member arrays (exact), velocities zero, masses of the block (cross: zero)
        DBuf<float> stage;
        CU_TRY(stage.alloc(4 * static_cast<size_t>(n_tgt)));
        const float *h[4] = {tx, ty, tz, cross ? nullptr : sm + off};
        for (int k = 0; k < 4; ++k)
            if (h[k]) CU_TRY(cudaMemcpyAsync(stage.p + k * n_tgt, h[k], tb, cudaMemcpyHostToDevice, s));
        const int grid = static_cast<int>(std::min<int64_t>((n_tgt + 255) / 256, P->ctx->sm_count * 8));
        k_widen_members<<<grid, 256, 0, s>>>(stage.p, stage.p + n_tgt, stage.p + 2 * n_tgt,
                                             cross ? nullptr : stage.p + 3 * n_tgt, n_tgt, P->d_in.p);
        CU_TRY(cudaGetLastError());
        P->members_up = true;
        P->sorted_dirty = true;
    }
    // the external groups are float32 already: straight into their packed segments (one halo: group g starts
    // where group g-1 ended, rounded up to 4 elements, plan_build)
    const size_t NE = static_cast<size_t>(P->n_ext_pad);
    int64_t seg_begin = 0;
    for (int g = 0; g < n_groups; ++g) {
        const size_t sb = static_cast<size_t>(g_count[g]) * sizeof(float);
        float *e = P->d_ext.p + seg_begin;
        CU_TRY(cudaMemcpyAsync(e, sx + g_begin[g], sb, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(e + NE, sy + g_begin[g], sb, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(e + 2 * NE, sz + g_begin[g], sb, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(e + 3 * NE, sm + g_begin[g], sb, cudaMemcpyHostToDevice, s));
        P->group_up[g] = true;
        seg_begin += up4(g_count[g]);
    }
    const double vb0[3] = {0.0, 0.0, 0.0};
    rc = halma_plan_set_vb(P, vb0);
    if (!rc) rc = halma_plan_run(P, nullptr);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(out_be, P->d_be.p, tb, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return HALMA_OK;
}
