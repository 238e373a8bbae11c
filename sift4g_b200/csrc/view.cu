// NVLink-striped database: one resident stripe of the residue array per GPU, every GPU maps ALL stripes into one
// contiguous virtual range and reads its peers' pages over NVLink (sm_100a boxes: NVLink 5 / NVSwitch).
//
// The reference spreads a database over its cards by handing each card thread a slice of the sequences and merging
// the per-card results on the host (vendor/swsharp/swsharp/src/database.c:497-532, scoreDatabasesGpu
// src/gpu_module.h:280-291).  Here the *memory* is sharded and the *queries* are: a rank keeps 1/N of the residues in
// its HBM, maps the other stripes through the CUDA virtual-memory API (cuMemCreate / cuMemExportToShareableHandle /
// cuMemMap / cuMemSetAccess) and runs the whole hot path for its own queries against the full database -- the kernels
// see one `codes` array and do not know which pages are remote.  No candidate or hit ever has to be merged, so the
// data path needs no collective at all; what crosses NVLink is the residue stream of the prefilter (coalesced 128-byte
// lines, ~120 GB/s per GPU at configs[1]) and the residues of the candidates the SW kernels stage (sector gathers).
//
// Driver entry points are resolved through cudaGetDriverEntryPoint, so the library keeps linking against cudart only
// (it must load on hosts without a driver: tests/test_abi.py).
#include <cuda.h>
#include <unistd.h>

#include "common.cuh"

struct s4g_stripe {
    s4g_ctx* ctx = nullptr;
    CUmemGenericAllocationHandle handle = 0;
    uint64_t bytes = 0;
};

struct s4g_view {
    s4g_ctx* ctx = nullptr;
    CUdeviceptr base = 0;
    uint64_t bytes = 0;
    std::vector<CUmemGenericAllocationHandle> imported;      // released on close (local stripes stay with their owner)
    std::vector<std::pair<uint64_t, uint64_t>> mapped;       // (offset, bytes) of every mapping
    uint64_t local_lo = 0, local_hi = 0;                     // byte range of the stripes handed in as local handles (resident on ctx's device)
};

namespace {

struct Drv {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
};

template <typename F>
bool resolve(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return false; }
    fn = reinterpret_cast<F>(p);
    return true;
}

Drv& drv() {
    static Drv d = [] {
        Drv x;
        x.ok = resolve("cuMemCreate", x.MemCreate) && resolve("cuMemRelease", x.MemRelease) &&
               resolve("cuMemGetAllocationGranularity", x.MemGetAllocationGranularity) &&
               resolve("cuMemExportToShareableHandle", x.MemExportToShareableHandle) &&
               resolve("cuMemImportFromShareableHandle", x.MemImportFromShareableHandle) &&
               resolve("cuMemAddressReserve", x.MemAddressReserve) && resolve("cuMemAddressFree", x.MemAddressFree) &&
               resolve("cuMemMap", x.MemMap) && resolve("cuMemUnmap", x.MemUnmap) && resolve("cuMemSetAccess", x.MemSetAccess) &&
               resolve("cuGetErrorString", x.GetErrorString);
        return x;
    }();
    return d;
}

const char* cu_text(CUresult r) {
    const char* s = nullptr;
    if (drv().GetErrorString && drv().GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "unknown driver error";
}

#define S4G_CU(ctx, call)                                                                                     \
    do {                                                                                                      \
        CUresult r_ = (call);                                                                                 \
        if (r_ != CUDA_SUCCESS) {                                                                             \
            s4g_set_error((ctx), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cu_text(r_));                  \
            return S4G_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)

CUmemAllocationProp stripe_prop(int device) {
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

int enter(s4g_ctx* ctx) {
    if (!ctx) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    S4G_CUDA(ctx, cudaFree(nullptr));                 // the runtime's primary context becomes current for the driver calls
    if (!drv().ok) { s4g_set_error(ctx, "the CUDA driver does not export the virtual-memory API (cuMemCreate ...)"); return S4G_ERR_CUDA; }
    return S4G_OK;
}

// plain grid-stride copies / fills through the view: stores to pages of a peer go out over NVLink
__global__ void view_copy_kernel(uint8_t* dst, const uint8_t* src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 16-byte body when both sides are aligned alike, bytes otherwise
    if ((((uintptr_t)dst ^ (uintptr_t)src) & 15u) == 0) {
        uint64_t head = (16 - ((uintptr_t)dst & 15u)) & 15u;
        if (head > n) head = n;
        if (i < head) dst[i] = src[i];
        const uint64_t nv = (n - head) / 16;
        uint4* d4 = reinterpret_cast<uint4*>(dst + head);
        const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
        for (uint64_t v = i; v < nv; v += stride) d4[v] = s4[v];
        const uint64_t done = head + nv * 16;
        if (i < n - done) dst[done + i] = src[done + i];
    } else {
        for (; i < n; i += stride) dst[i] = src[i];
    }
}

__global__ void view_fill_kernel(uint8_t* dst, uint8_t value, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

}  // namespace

extern "C" {

uint64_t s4g_stripe_granularity(s4g_ctx* ctx) {
    if (enter(ctx) != S4G_OK) return 0;
    const CUmemAllocationProp p = stripe_prop(ctx->device);
    size_t g = 0;
    if (drv().MemGetAllocationGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return 0;
    return (uint64_t)g;
}

int s4g_stripe_create(s4g_ctx* ctx, uint64_t bytes, s4g_stripe** out) {
    if (!out || bytes == 0) return S4G_ERR_ARG;
    *out = nullptr;
    int rc = enter(ctx);
    if (rc != S4G_OK) return rc;
    const uint64_t g = s4g_stripe_granularity(ctx);
    if (g == 0 || bytes % g != 0) { s4g_set_error(ctx, "s4g_stripe_create: %llu bytes is not a multiple of the allocation granularity %llu", (unsigned long long)bytes, (unsigned long long)g); return S4G_ERR_ARG; }
    const CUmemAllocationProp p = stripe_prop(ctx->device);
    CUmemGenericAllocationHandle h = 0;
    S4G_CU(ctx, drv().MemCreate(&h, (size_t)bytes, &p, 0));
    s4g_stripe* s = new s4g_stripe();
    s->ctx = ctx; s->handle = h; s->bytes = bytes;
    *out = s;
    return S4G_OK;
}

uint64_t s4g_stripe_bytes(const s4g_stripe* s) { return s ? s->bytes : 0; }

int s4g_stripe_export_fd(s4g_stripe* s, int* out_fd) {
    if (!s || !out_fd) return S4G_ERR_ARG;
    int rc = enter(s->ctx);
    if (rc != S4G_OK) return rc;
    int fd = -1;
    S4G_CU(s->ctx, drv().MemExportToShareableHandle(&fd, s->handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *out_fd = fd;
    return S4G_OK;
}

void s4g_stripe_free(s4g_stripe* s) {
    if (!s) return;
    if (enter(s->ctx) == S4G_OK && s->handle) drv().MemRelease(s->handle);     // the memory goes when the last mapping does
    delete s;
}

void s4g_view_close(s4g_view* v) {
    if (!v) return;
    if (enter(v->ctx) == S4G_OK) {
        cudaStreamSynchronize(v->ctx->stream);
        for (auto& m : v->mapped) drv().MemUnmap(v->base + m.first, (size_t)m.second);
        for (auto h : v->imported) drv().MemRelease(h);
        if (v->base) drv().MemAddressFree(v->base, (size_t)v->bytes);
    }
    delete v;
}

int s4g_view_open(s4g_ctx* ctx, int n_stripes, s4g_stripe* const* local, const int* fds, const uint64_t* bytes, s4g_view** out) {
    if (!out || n_stripes < 1 || !bytes || (!local && !fds)) return S4G_ERR_ARG;
    *out = nullptr;
    int rc = enter(ctx);
    if (rc != S4G_OK) return rc;
    uint64_t total = 0;
    for (int i = 0; i < n_stripes; ++i) {
        if (bytes[i] == 0) return S4G_ERR_ARG;
        if (local && local[i] && local[i]->bytes != bytes[i]) { s4g_set_error(ctx, "s4g_view_open: stripe %d holds %llu bytes, not %llu", i, (unsigned long long)local[i]->bytes, (unsigned long long)bytes[i]); return S4G_ERR_ARG; }
        if (!(local && local[i]) && (!fds || fds[i] < 0)) { s4g_set_error(ctx, "s4g_view_open: stripe %d has neither a local handle nor a file descriptor", i); return S4G_ERR_ARG; }
        total += bytes[i];
    }
    const uint64_t g = s4g_stripe_granularity(ctx);
    s4g_view* v = new s4g_view();
    v->ctx = ctx; v->bytes = total;
    rc = [&]() -> int {
        S4G_CU(ctx, drv().MemAddressReserve(&v->base, (size_t)total, (size_t)g, 0, 0));
        uint64_t at = 0;
        for (int i = 0; i < n_stripes; ++i) {
            CUmemGenericAllocationHandle h;
            if (local && local[i]) {
                h = local[i]->handle;
                if (local[i]->ctx->device == ctx->device) {         // contiguous local stripes form one resident range
                    if (v->local_hi == v->local_lo) { v->local_lo = at; v->local_hi = at + bytes[i]; }
                    else if (v->local_hi == at) v->local_hi = at + bytes[i];
                }
            } else {
                S4G_CU(ctx, drv().MemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[i], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
                v->imported.push_back(h);
            }
            S4G_CU(ctx, drv().MemMap(v->base + at, (size_t)bytes[i], 0, h, 0));
            v->mapped.emplace_back(at, bytes[i]);
            at += bytes[i];
        }
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = ctx->device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        S4G_CU(ctx, drv().MemSetAccess(v->base, (size_t)total, &acc, 1));
        return S4G_OK;
    }();
    if (rc != S4G_OK) { s4g_view_close(v); return rc; }
    *out = v;
    return S4G_OK;
}

void* s4g_view_ptr(const s4g_view* v) { return v ? (void*)(uintptr_t)v->base : nullptr; }
uint64_t s4g_view_bytes(const s4g_view* v) { return v ? v->bytes : 0; }
void s4g_view_local_range(const s4g_view* v, uint64_t* lo, uint64_t* hi) { *lo = v ? v->local_lo : 0; *hi = v ? v->local_hi : 0; }

int s4g_view_write(s4g_view* v, uint64_t at, const uint8_t* d_src, uint64_t bytes) {
    if (!v || (bytes > 0 && !d_src) || at > v->bytes || bytes > v->bytes - at) return S4G_ERR_ARG;
    s4g_ctx* ctx = v->ctx;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bytes == 0) return S4G_OK;
    view_copy_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint8_t*)(uintptr_t)v->base + at, d_src, bytes);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}

int s4g_view_fill(s4g_view* v, uint64_t at, int value, uint64_t bytes) {
    if (!v || at > v->bytes || bytes > v->bytes - at) return S4G_ERR_ARG;
    s4g_ctx* ctx = v->ctx;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bytes == 0) return S4G_OK;
    view_fill_kernel<<<(unsigned)std::min<uint64_t>((bytes + 255) / 256, (uint64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>((uint8_t*)(uintptr_t)v->base + at, (uint8_t)value, bytes);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}

}  // extern "C"
