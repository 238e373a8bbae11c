// Measurement helper: sustained issue rate of the DPX / integer-ALU pipe on this device.
// The SW roofline ("Algorithmic work definitions") is
//     roofline_GCUPS = lane_ops_per_s * 2 (s16x2 lanes) / 6 (instructions per cell update)
// and this kernel measures lane_ops_per_s with independent VIADDMNMX.S16x2 chains (ILP 8, 32 warps/SM).
#include "common.cuh"

namespace {
constexpr int kIlp = 8;
__global__ void __launch_bounds__(256) dpx_peak_kernel(unsigned* out, unsigned seed, int iters) {
    unsigned v[kIlp];
    unsigned b = seed * 0x9e3779b9u + threadIdx.x, c = seed ^ 0x7f4a7c15u;
#pragma unroll
    for (int i = 0; i < kIlp; ++i) v[i] = threadIdx.x * 7919u + i * 104729u + seed;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kIlp; ++i) v[i] = __viaddmax_s16x2(v[i], b, c);
#pragma unroll
        for (int i = 0; i < kIlp; ++i) v[i] = __viaddmax_s16x2(v[i], c, b);
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < kIlp; ++i) r ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
}  // namespace

extern "C" int s4g_measure_dpx_peak(s4g_ctx* ctx, int millis, double* lane_ops_per_s) {
    if (!ctx || !lane_ops_per_s) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 4, threads = 256;
    unsigned* d_out = (unsigned*)s4g_scratch(ctx, SLOT_IO_F, sizeof(unsigned) * blocks * threads);
    if (!d_out) return S4G_ERR_NOMEM;
    cudaEvent_t e0, e1;
    S4G_CUDA(ctx, cudaEventCreate(&e0));
    S4G_CUDA(ctx, cudaEventCreate(&e1));
    int iters = 4096;
    dpx_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, 1u, iters);   // warm-up
    S4G_CHECK_LAUNCH(ctx);
    double best = 0.0, total_ms = 0.0;
    if (millis < 1) millis = 1;
    // repeat back to back until `millis` of device time has been spent; report the rate over the whole
    // window (sustained), not the best burst
    double ops = 0.0;
    S4G_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    int launches = 0;
    while (true) {
        for (int r = 0; r < 8; ++r) {
            dpx_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, 2u + launches, iters);
            ++launches;
        }
        ctx->launches += 8;
        S4G_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        S4G_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        S4G_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        total_ms = ms;
        if (ms >= millis || launches > 100000) break;
    }
    ops = (double)launches * blocks * threads * (double)iters * kIlp * 2.0;
    best = ops / (total_ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lane_ops_per_s = best;
    return S4G_OK;
}

// Measurement helper: rate of random 8-byte lookups into an L2-resident table -- what bounds the prefilter scan, which probes
// the 8 MB presence/rank table once per k-mer position.  An SM completes one random 32-byte sector request per clock
// whatever the load flavour (tools/gather_microbench.cu: 290-320 G lookups/s on a B200), far below L2 bandwidth.
namespace {
__device__ __forceinline__ uint32_t gp_mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(1024, 1) gather_peak_kernel(const uint2* tab, uint32_t mask, int rounds, unsigned* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, seed = gp_mix(tid * 0x9E3779B1u + 12345u);
    for (int r = 0; r < rounds; ++r) {
        uint32_t idx[4];
        uint2 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { seed = seed * 1664525u + 1013904223u; idx[i] = gp_mix(seed + i); }
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldg(tab + (idx[i] & mask));
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += v[i].x ^ v[i].y;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
}  // namespace

extern "C" int s4g_measure_gather_peak(s4g_ctx* ctx, int millis, double* lookups_per_s) {
    if (!ctx || !lookups_per_s) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t entries = (size_t)1 << 20;                       // 8 MiB: the size of the scan's presence/rank table
    uint2* d_tab = (uint2*)s4g_scratch(ctx, SLOT_IO_E, sizeof(uint2) * entries + 64);
    if (!d_tab) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cudaMemsetAsync(d_tab, 0x5a, sizeof(uint2) * entries, ctx->stream));
    unsigned* d_out = (unsigned*)(d_tab + entries);
    cudaEvent_t e0, e1;
    S4G_CUDA(ctx, cudaEventCreate(&e0));
    S4G_CUDA(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->sm_count, threads = 1024, rounds = 2048;
    gather_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_tab, (uint32_t)entries - 1, rounds, d_out);
    S4G_CHECK_LAUNCH(ctx);
    if (millis < 1) millis = 1;
    int launches = 0;
    float ms = 0.f;
    S4G_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    while (true) {
        for (int r = 0; r < 4; ++r) { gather_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_tab, (uint32_t)entries - 1, rounds, d_out); ++launches; }
        ctx->launches += 4;
        S4G_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        S4G_CUDA(ctx, cudaEventSynchronize(e1));
        S4G_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (ms >= millis || launches > 100000) break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lookups_per_s = (double)launches * blocks * threads * (double)rounds * 4.0 / (ms * 1e-3);
    return S4G_OK;
}
