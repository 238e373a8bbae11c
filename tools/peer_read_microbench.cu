// Remote (NVLink peer) read behaviour of access patterns the prefilter / SW kernels use.  Single process, 2 GPUs:
// memory on device 1 (cudaMalloc + peer access, and a VMM mapping like csrc/view.cu makes), kernels on device 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peer_read_microbench tools/peer_read_microbench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("%s:%d %s\n", __FILE__, __LINE__, s); exit(1); } } while (0)

__global__ void stream16(const uint4* p, size_t n, unsigned* out) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345) *out = acc;
}

__global__ void chase(const unsigned* p, int steps, unsigned* out, long long* cycles) {
    unsigned i = 0;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) i = p[i];
    long long t1 = clock64();
    *out = i; *cycles = t1 - t0;
}

// one warp per "sequence" of `len` bytes at a random offset; steps of 128 bytes; MODE 0: 3 x ld.nc 4B per lane (the scan's
// pattern) 1: same + prefetch.global.L2 of the sequence first  2: 1 x 4B load per lane per step, next step prefetched in a register
// 3: like 0 with plain ld.global (no .nc)  4: like 0 but 4 sequences in flight per warp (loads of all issued before use)
template <int MODE>
__global__ void seqwalk(const unsigned char* base, const unsigned* offs, int nseq, int len, unsigned* out, unsigned long long* cursor) {
    const int lane = threadIdx.x & 31;
    unsigned acc = 0;
    while (true) {
        unsigned long long s = 0;
        if (lane == 0) s = atomicAdd(cursor, MODE == 4 ? 4ull : 1ull);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= (unsigned long long)nseq) break;
        if (MODE == 4) {
            for (int base_j = 0; base_j < len; base_j += 128) {
                unsigned w[4][3];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const unsigned* q = reinterpret_cast<const unsigned*>(base + offs[s + b] + base_j) + lane;
                    w[b][0] = __ldg(q); w[b][1] = __ldg(q + 1); w[b][2] = __ldg(q + 2);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) acc += __funnelshift_r(w[b][0], w[b][1], 8) ^ w[b][2];
            }
            continue;
        }
        const unsigned char* seq = base + offs[s];
        if (MODE == 1) for (int a = 128 * lane; a < len; a += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(seq + a));
        if (MODE == 2) {
            const unsigned* q = reinterpret_cast<const unsigned*>(seq) + lane;
            unsigned nxt = __ldg(q);
            for (int j = 0; j < len; j += 128) {
                const unsigned cur = nxt;
                if (j + 128 < len) nxt = __ldg(q + (j + 128) / 4);
                const unsigned w1 = __shfl_down_sync(0xffffffffu, cur, 1);
                acc += __funnelshift_r(cur, w1, 8);
                // emulate ~200 cycles of dependent work per step
                for (int k = 0; k < 50; ++k) acc = acc * 1664525u + 1013904223u;
            }
        } else {
            for (int j = 0; j < len; j += 128) {
                const unsigned* q = reinterpret_cast<const unsigned*>(seq + j) + lane;
                unsigned w0, w1, w2;
                if (MODE == 3) { w0 = q[0]; w1 = q[1]; w2 = q[2]; }
                else { w0 = __ldg(q); w1 = __ldg(q + 1); w2 = __ldg(q + 2); }
                acc += __funnelshift_r(w0, w1, 8) ^ w2;
                for (int k = 0; k < 50; ++k) acc = acc * 1664525u + 1013904223u;
            }
        }
    }
    if (acc == 0x12345) *out = acc;
}

template <int MODE>
float run_walk(const unsigned char* base, const unsigned* offs, int nseq, int len, unsigned* out, unsigned long long* cursor, int warps_per_sm) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 3; ++it) {
        CK(cudaMemset(cursor, 0, 8));
        cudaEventRecord(a);
        seqwalk<MODE><<<148, warps_per_sm * 32>>>(base, offs, nseq, len, out, cursor);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t bytes = (size_t)1 << 30;
    CK(cudaSetDevice(0)); CK(cudaFree(0));
    cudaDeviceEnablePeerAccess(1, 0);
    unsigned char *loc, *peer;
    CK(cudaMalloc(&loc, bytes));
    CK(cudaSetDevice(1)); CK(cudaFree(0)); CK(cudaMalloc(&peer, bytes)); CK(cudaMemset(peer, 1, bytes));
    // VMM memory on device 1 mapped for device 0
    CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 1;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h; CU(cuMemCreate(&h, bytes, &prop, 0));
    CK(cudaSetDevice(0));
    CUdeviceptr va; CU(cuMemAddressReserve(&va, bytes, 2 << 20, 0, 0)); CU(cuMemMap(va, bytes, 0, h, 0));
    CUmemAccessDesc acc[2] = {}; acc[0].location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc[0].location.id = 0; acc[0].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    acc[1] = acc[0]; acc[1].location.id = 1;
    CU(cuMemSetAccess(va, bytes, acc, 2));
    unsigned char* vmm = (unsigned char*)va;
    CK(cudaMemset(loc, 1, bytes)); CK(cudaMemset(vmm, 1, bytes));
    unsigned* out; CK(cudaMalloc(&out, 64)); unsigned long long* cursor; CK(cudaMalloc(&cursor, 8));
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    const char* names[3] = {"local cudaMalloc", "peer cudaMalloc", "peer VMM map"};
    unsigned char* bufs[3] = {loc, peer, vmm};
    // random sequence offsets (4-byte aligned), 323-byte sequences
    const int nseq = 2000000, len = 323;
    std::vector<unsigned> h_offs(nseq + 8);
    unsigned x = 12345;
    for (auto& o : h_offs) { x = x * 1664525u + 1013904223u; o = (x % (unsigned)(bytes - 4096)) & ~3u; }
    unsigned* offs; CK(cudaMalloc(&offs, 4 * h_offs.size())); CK(cudaMemcpy(offs, h_offs.data(), 4 * h_offs.size(), cudaMemcpyHostToDevice));
    // pointer-chase table (stride 4 KiB + 4)
    {
        std::vector<unsigned> t(bytes / 4 / 64);   // 16 MiB worth of entries used
        for (size_t i = 0; i < t.size(); ++i) t[i] = (unsigned)((i * 1031 + 7) % t.size());
        for (int b = 0; b < 3; ++b) CK(cudaMemcpy(bufs[b], t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int bi = 0; bi < 3; ++bi) {
        chase<<<1, 1>>>((const unsigned*)bufs[bi], 2000, out, cyc); CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        cudaEventRecord(a); stream16<<<148 * 8, 256>>>((const uint4*)bufs[bi], bytes / 16, out); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        cudaEventRecord(a); stream16<<<148 * 8, 256>>>((const uint4*)bufs[bi], bytes / 16, out); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
        printf("{\"memory\": \"%s\", \"chase_cycles_per_load\": %.0f, \"stream16_GBps\": %.1f", names[bi], (double)c / 2000, bytes / ms / 1e6);
        for (int wps : {8, 32}) {
            printf(", \"walk_ldg3_w%d_ms\": %.2f", wps, run_walk<0>(bufs[bi], offs, nseq, len, out, cursor, wps));
            printf(", \"walk_ldg3_prefetchL2_w%d_ms\": %.2f", wps, run_walk<1>(bufs[bi], offs, nseq, len, out, cursor, wps));
            printf(", \"walk_1load_regprefetch_w%d_ms\": %.2f", wps, run_walk<2>(bufs[bi], offs, nseq, len, out, cursor, wps));
            printf(", \"walk_ld3_plain_w%d_ms\": %.2f", wps, run_walk<3>(bufs[bi], offs, nseq, len, out, cursor, wps));
            printf(", \"walk_ldg3_4seqs_w%d_ms\": %.2f", wps, run_walk<4>(bufs[bi], offs, nseq, len, out, cursor, wps));
        }
        printf(", \"walk_bytes\": %.0f}\n", (double)nseq * len);
        fflush(stdout);
    }
    return 0;
}
