// Measurement tool (not product code): what bounds the prefilter's index probes on a B200?
// Whole-GPU rate of random lookups through the paths the scan kernel could use:
//   ldg64     random 8-byte __ldg from a table of `mb` MiB (the bitrank probe of pf_scan_kernel)
//   ldg32     random 4-byte __ldg (bits-only table)
//   ldcg64    the same through ld.global.cg (no L1 allocation)
//   tex64     tex1Dfetch<uint2> from a texture object over the same table (TEX pipe instead of the LSU pipe)
//   mix       two __ldg + two tex fetches per lane and round
//   lds       random 4-byte LDS from a 128 KiB table in shared memory (a Bloom filter in front of the probe)
//   lds+ldg   LDS filter, then the __ldg predicated on a filter bit that passes `pass`% of the lanes
//   atoms     random shared-memory atomicAdd on packed 16-bit counters (2 KiB per warp), result used
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_microbench tools/gather_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

constexpr int kIL = 4;       // independent lookups per lane and round

template <int MODE>
__global__ void __launch_bounds__(1024, 1) gather_kernel(const uint2* tab, uint32_t mask, cudaTextureObject_t tex, int rounds, int pass_pct,
                                                         unsigned long long* sink) {
    extern __shared__ uint32_t smem[];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE >= 5) {
        const int words = MODE == 7 ? (blockDim.x / 32) * 512 : 32768;
        for (int i = threadIdx.x; i < words; i += blockDim.x) smem[i] = MODE == 7 ? 0u : mix32(i * 2654435761u + 17u);
        __syncthreads();
    }
    uint32_t acc = 0;
    uint32_t seed = mix32(tid * 0x9E3779B1u + 12345u);
    const uint32_t thresh = pass_pct >= 100 ? 0xffffffffu : (uint32_t)((double)pass_pct * 42949672.96);
    uint32_t* my_cnt = smem + (threadIdx.x >> 5) * 512;
    for (int r = 0; r < rounds; ++r) {
        uint32_t idx[kIL];
#pragma unroll
        for (int i = 0; i < kIL; ++i) { seed = seed * 1664525u + 1013904223u; idx[i] = mix32(seed + i); }
        if (MODE == 0) {
            uint2 v[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = __ldg(tab + (idx[i] & mask));
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i].x ^ v[i].y;
        } else if (MODE == 1) {
            uint32_t v[kIL];
            const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tab);
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = __ldg(t32 + (idx[i] & mask));
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i];
        } else if (MODE == 2) {
            uint2 v[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = __ldcg(tab + (idx[i] & mask));
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i].x ^ v[i].y;
        } else if (MODE == 3) {
            uint2 v[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = tex1Dfetch<uint2>(tex, (int)(idx[i] & mask));
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i].x ^ v[i].y;
        } else if (MODE == 4) {
            uint2 v[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = (i & 1) ? tex1Dfetch<uint2>(tex, (int)(idx[i] & mask)) : __ldg(tab + (idx[i] & mask));
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i].x ^ v[i].y;
        } else if (MODE == 5) {
            uint32_t v[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = smem[idx[i] & 32767u];
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i];
        } else if (MODE == 6) {
            uint32_t f[kIL];
            uint2 v[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) f[i] = smem[idx[i] & 32767u];
#pragma unroll
            for (int i = 0; i < kIL; ++i) v[i] = (f[i] <= thresh) ? __ldg(tab + (idx[i] & mask)) : make_uint2(0u, 0u);
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += v[i].x ^ v[i].y;
        } else if (MODE == 7) {
            uint32_t old[kIL];
#pragma unroll
            for (int i = 0; i < kIL; ++i) { const uint32_t s = idx[i] & 1023u; old[i] = atomicAdd(my_cnt + (s >> 1), (s & 1u) ? 0x10000u : 1u); }
#pragma unroll
            for (int i = 0; i < kIL; ++i) acc += old[i];
        }
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int MODE>
static double run(const char* name, const uint2* tab, uint32_t mask, cudaTextureObject_t tex, int ctas, int threads, int rounds, int pass_pct,
                  unsigned long long* sink, size_t smem) {
    CK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather_kernel<MODE><<<ctas, threads, smem>>>(tab, mask, tex, rounds / 8, pass_pct, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        gather_kernel<MODE><<<ctas, threads, smem>>>(tab, mask, tex, rounds, pass_pct, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double n = (double)ctas * threads * (double)rounds * kIL;
    const double rate = n / (best * 1e-3) / 1e9;
    printf("{\"mode\": \"%s\", \"table_mib\": %.2f, \"ctas\": %d, \"threads\": %d, \"pass_pct\": %d, \"ms\": %.3f, \"glookups_per_s\": %.1f}\n", name,
           (double)(mask + 1) * (MODE == 1 ? 4 : 8) / 1048576.0, ctas, threads, pass_pct, best, rate);
    fflush(stdout);
    return rate;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t max_entries = (size_t)1 << 22;            // 32 MiB of uint2
    uint2* tab;
    CK(cudaMalloc(&tab, max_entries * sizeof(uint2)));
    CK(cudaMemset(tab, 0x5a, max_entries * sizeof(uint2)));
    unsigned long long* sink;
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(sink, 0, 8));
    const int rounds = 2048;
    for (int lg = 17; lg <= 22; ++lg) {                       // 1 MiB .. 32 MiB tables of 8-byte entries
        if (lg == 18 || lg == 21) continue;
        const uint32_t mask = (1u << lg) - 1u;
        cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = tab;
        rd.res.linear.desc = cudaCreateChannelDesc<uint2>();
        rd.res.linear.sizeInBytes = ((size_t)mask + 1) * sizeof(uint2);
        cudaTextureDesc td; memset(&td, 0, sizeof(td));
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        for (int threads : {1024, 512}) {
            run<0>("ldg64", tab, mask, tex, sms, threads, rounds, 100, sink, 0);
            run<2>("ldcg64", tab, mask, tex, sms, threads, rounds, 100, sink, 0);
            run<3>("tex64", tab, mask, tex, sms, threads, rounds, 100, sink, 0);
            run<4>("mix_ldg_tex", tab, mask, tex, sms, threads, rounds, 100, sink, 0);
        }
        run<1>("ldg32", tab, mask, tex, sms, 1024, rounds, 100, sink, 0);
        for (int pass : {25, 50, 75}) run<6>("lds_then_ldg64", tab, mask, tex, sms, 1024, rounds, pass, sink, 131072);
        CK(cudaDestroyTextureObject(tex));
    }
    run<5>("lds32_128k", tab, 0, 0, sms, 1024, rounds, 100, sink, 131072);
    run<5>("lds32_128k", tab, 0, 0, sms, 512, rounds, 100, sink, 131072);
    run<7>("atoms16", tab, 0, 0, sms, 1024, rounds, 100, sink, 65536);
    run<7>("atoms16", tab, 0, 0, sms, 512, rounds, 100, sink, 32768);
    return 0;
}
