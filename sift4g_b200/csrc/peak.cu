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
