// Integer / DPX issue-rate microbenchmark for sm_100a.
// Measures sustained warp-instructions per clock per SM for the instruction
// classes the Smith-Waterman kernels are built from.  The result is the
// denominator of the SW "integer/DPX roofline" (DESIGN.md, section Measurement).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_microbench dpx_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__device__ __forceinline__ unsigned apply(unsigned a, unsigned b, unsigned c) {
    if (OP == 0) return __viaddmax_s16x2(a, b, c);
    if (OP == 1) return __vimax3_s16x2(a, b, c);
    if (OP == 2) return __viaddmax_s16x2_relu(a, b, c);
    if (OP == 3) return __vimax3_s16x2_relu(a, b, c);
    if (OP == 4) return (unsigned)__viaddmax_s32((int)a, (int)b, (int)c);
    if (OP == 5) return (unsigned)__vimax3_s32((int)a, (int)b, (int)c);
    if (OP == 6) { unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; } // IADD
    if (OP == 7) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
    if (OP == 8) return __vmaxs2(a, b) ^ c;              // max.s16x2 (+xor)
    if (OP == 9) return (unsigned)max((int)a, (int)b) + c; // IMNMX + IADD
    if (OP == 10) return a * b + c;                       // IMAD
    if (OP == 11) return __vadd2(a, b) ^ c;
    if (OP == 12) return __vsub2(a, b) ^ c;
    if (OP == 13) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
    if (OP == 14) return __shfl_up_sync(0xffffffffu, a, 1) + c;
    if (OP == 15) return __vimax_s16x2_relu(a, b) + c;
    if (OP == 16) {   // SW cell mix: PRMT + 6 DPX/SIMD ops (one s16x2 cell pair)
        unsigned sc, m, h, hq, e, f;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(a), "r"(c), "r"(0xC480u));
        m = __viaddmax_s16x2_relu(a, sc, c);
        h = __vmaxs2(m, b);
        hq = __vadd2(h, 0xfff6fff6u);
        e = __viaddmax_s16x2(c, 0xffffffffu, hq);
        f = __viaddmax_s16x2(b, 0xffffffffu, hq);
        return __vimax3_s16x2(e, f, h);
    }
    return a;
}

template <int OP>
__global__ void __launch_bounds__(256) bench(unsigned* out, unsigned seed, long long* cycles) {
    unsigned v[ILP];
    unsigned b = seed * 0x9e3779b9u + threadIdx.x, c = seed ^ 0x7f4a7c15u;
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 7919u + i * 104729u + seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = apply<OP>(v[i], b, c);
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = apply<OP>(v[i], c, b);
    }
    long long t1 = clock64();
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name, int instr_per_apply, int nsm, unsigned* d_out, long long* d_cyc) {
    const int blocks = nsm * 4, threads = 256;   // 4 CTAs x 8 warps = 32 warps / SM
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<OP><<<blocks, threads>>>(d_out, 1u, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    bench<OP><<<blocks, threads>>>(d_out, 2u, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc[1]; CK(cudaMemcpy(cyc, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    double applies = (double)blocks * threads * ITERS * ILP * 2;
    double warp_instr = applies / 32.0;
    double per_s = applies / (ms * 1e-3);
    // per-SM per-clock using in-kernel cycle count of block 0 (all blocks co-resident)
    double wi_per_clk_sm = (warp_instr / nsm) / (double)cyc[0];
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"lane_ops_per_s\": %.4e, \"warp_instr_per_clk_per_sm\": %.3f, \"cycles\": %lld, \"sass_instr_per_op\": %d}\n",
           name, ms, per_s, wi_per_clk_sm, cyc[0], instr_per_apply);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, clk);
    unsigned* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, (size_t)nsm * 4 * 256 * 4));
    CK(cudaMalloc(&d_cyc, (size_t)nsm * 4 * 8));
    run<0>("viaddmax_s16x2", 1, nsm, d_out, d_cyc);
    run<1>("vimax3_s16x2", 1, nsm, d_out, d_cyc);
    run<2>("viaddmax_s16x2_relu", 1, nsm, d_out, d_cyc);
    run<3>("vimax3_s16x2_relu", 1, nsm, d_out, d_cyc);
    run<4>("viaddmax_s32", 1, nsm, d_out, d_cyc);
    run<5>("vimax3_s32", 1, nsm, d_out, d_cyc);
    run<6>("iadd3", 1, nsm, d_out, d_cyc);
    run<7>("prmt", 1, nsm, d_out, d_cyc);
    run<8>("vmaxs2_xor", 2, nsm, d_out, d_cyc);
    run<9>("imnmx_iadd", 2, nsm, d_out, d_cyc);
    run<10>("imad", 1, nsm, d_out, d_cyc);
    run<11>("vadd2_xor", 2, nsm, d_out, d_cyc);
    run<12>("vsub2_xor", 2, nsm, d_out, d_cyc);
    run<13>("lop3", 1, nsm, d_out, d_cyc);
    run<14>("shfl_up_add", 2, nsm, d_out, d_cyc);
    run<15>("vimax_s16x2_relu_add", 2, nsm, d_out, d_cyc);
    run<16>("sw_cell_mix7", 7, nsm, d_out, d_cyc);
    return 0;
}
