// The bare step loop of the packed score kernel (csrc/sw_score.cu) without pair boundaries, staging or tiles: what the loop alone
// sustains, and how many issue slots a cell costs in SASS, for the forms that were tried (profiles/r04d_loop_microbench.jsonl,
// profiles/r04h_kernels_digest.md).
//   -DVARIANT=0  one stream column per step (stream_pairs_packed): ptxas parks the new H in a temporary and moves it back, one
//                IMAD.MOV per cell
//   -DVARIANT=1  the same with H - Q carried instead of H (profile bytes hold S + Q): fewer registers, the moves stay
//   -DVARIANT=2  running maximum taken from the previous column's H: ptxas folds the add into VIADDMNMX (5.5 ALU ops per cell) -- worse
//   -DVARIANT=3  two stream columns per step (stream_pairs_packed2, the shipped form): no moves
//   -DNOFLAG     without the pair-boundary block
// Build and run (B200):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DVARIANT=3 -DNOFLAG -o sw_loop_microbench_2col
//                        tools/sw_loop_microbench.cu && ./sw_loop_microbench_2col        (one JSON line per row class and CTAs per SM)
// SASS counts:           nvcc ... -cubin, cuobjdump -sass, count the instructions of the loop that holds the VIADDMNMX.
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}
#ifndef VARIANT
#define VARIANT 0
#endif
template <int K>
__device__ __forceinline__ void body(const unsigned* prof_lane, const unsigned short* ring1, const unsigned short* ring2, int nsteps, unsigned negQ, unsigned negR, int* out, int lane) {
    constexpr int KW = (K + 3) / 4;
    const unsigned FULL = 0xffffffffu;
    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0, b_out = 0;
    int n31 = 0;
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
#pragma unroll 1
    for (int ss = 0; ss < nsteps; ++ss) {
        const int j = (ss - lane) & 127;
        int o1 = (short)ring1[j];
        const unsigned o2 = ring2[j];
        unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
        unsigned f = __shfl_up_sync(FULL, f_out, 1);
        unsigned b_in = __shfl_up_sync(FULL, b_out, 1);
        if (lane == 0) { h_up = 0; f = 0; b_in = 0; }
#ifndef NOFLAG
        if (o1 < 0) {
            o1 &= 0x7fff;
            b_out = __vmaxs2(best, b_in);
            if (lane == 31) { if (n31 > 0) out[n31] = b_out; ++n31; }
            best = 0; diag_in = 0;
#pragma unroll
            for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
        }
#endif
        unsigned w1[KW], w2[KW];
#pragma unroll
        for (int m = 0; m < KW; ++m) {
            w1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1 + m * 128);
            w2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2 + m * 128);
        }
#if VARIANT == 0
        unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;
        diag_in = h_up;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            unsigned t_next = 0;
            if (r + 1 < K) {
                const unsigned sel = ((r + 1) & 3) == 0 ? 0xC480u : ((r + 1) & 3) == 1 ? 0xD591u : ((r + 1) & 3) == 2 ? 0xE6A2u : 0xF7B3u;
                t_next = __vadd2(H[r], prmt(w1[(r + 1) >> 2], w2[(r + 1) >> 2], sel));
            }
            const unsigned h = __vimax3_s16x2_relu(t, E[r], f);
            H[r] = h;
            const unsigned hq = __vadd2(h, negQ);
            E[r] = __viaddmax_s16x2(E[r], negR, hq);
            f = __viaddmax_s16x2(f, negR, hq);
            if (r & 1) best = __vimax3_s16x2(best, t_prev, t);
            else if (r == K - 1) best = __vmaxs2(best, t);
            t_prev = t;
            t = t_next;
        }
        h_last = H[K - 1];
        f_out = f;
#elif VARIANT == 1
        // G formulation: H[] holds h - Q; profile bytes hold S + Q
        unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;
        diag_in = h_up;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            unsigned t_next = 0;
            if (r + 1 < K) {
                const unsigned sel = ((r + 1) & 3) == 0 ? 0xC480u : ((r + 1) & 3) == 1 ? 0xD591u : ((r + 1) & 3) == 2 ? 0xE6A2u : 0xF7B3u;
                t_next = __vadd2(H[r], prmt(w1[(r + 1) >> 2], w2[(r + 1) >> 2], sel));
            }
            const unsigned h = __vimax3_s16x2_relu(t, E[r], f);
            H[r] = __vadd2(h, negQ);
            E[r] = __viaddmax_s16x2(E[r], negR, H[r]);
            f = __viaddmax_s16x2(f, negR, H[r]);
            if (r & 1) best = __vimax3_s16x2(best, t_prev, t);
            else if (r == K - 1) best = __vmaxs2(best, t);
            t_prev = t;
            t = t_next;
        }
        h_last = H[K - 1];
        f_out = f;
#elif VARIANT == 2
        // best from the previous column's H (read before it is overwritten)
        unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u));
        diag_in = h_up;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            if (!(r & 1)) { if (r + 1 < K) best = __vimax3_s16x2(best, H[r], H[r + 1]); else best = __vmaxs2(best, H[r]); }
            unsigned t_next = 0;
            if (r + 1 < K) {
                const unsigned sel = ((r + 1) & 3) == 0 ? 0xC480u : ((r + 1) & 3) == 1 ? 0xD591u : ((r + 1) & 3) == 2 ? 0xE6A2u : 0xF7B3u;
                t_next = __vadd2(H[r], prmt(w1[(r + 1) >> 2], w2[(r + 1) >> 2], sel));
            }
            const unsigned h = __vimax3_s16x2_relu(t, E[r], f);
            H[r] = h;
            const unsigned hq = __vadd2(h, negQ);
            E[r] = __viaddmax_s16x2(E[r], negR, hq);
            f = __viaddmax_s16x2(f, negR, hq);
            t = t_next;
        }
        h_last = H[K - 1];
        f_out = f;
#endif
    }
    out[lane] = best + n31;
}

template <int K>
__device__ __forceinline__ void body2(const unsigned* prof_lane, const unsigned short* ring1, const unsigned short* ring2, int nsteps, unsigned negQ, unsigned negR, int* out, int lane) {
    constexpr int KW = (K + 3) / 4;
    const unsigned FULL = 0xffffffffu;
    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, hA_last = 0, hB_last = 0, fA_out = 0, fB_out = 0, diag_in = 0, b_out = 0;
    int n31 = 0;
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
    const unsigned* ring1w = reinterpret_cast<const unsigned*>(ring1);
    const unsigned* ring2w = reinterpret_cast<const unsigned*>(ring2);
#pragma unroll 1
    for (int ss = 0; ss < nsteps; ++ss) {
        const int j = (ss - lane) & 63;
        const unsigned p1 = ring1w[j], p2 = ring2w[j];      // two columns per word
        int o1a = (short)(p1 & 0xffffu);
        const unsigned o1b = p1 >> 16, o2a = p2 & 0xffffu, o2b = p2 >> 16;
        unsigned hA_up = __shfl_up_sync(FULL, hA_last, 1);
        unsigned hB_up = __shfl_up_sync(FULL, hB_last, 1);
        unsigned fA = __shfl_up_sync(FULL, fA_out, 1);
        unsigned fB = __shfl_up_sync(FULL, fB_out, 1);
        unsigned b_in = __shfl_up_sync(FULL, b_out, 1);
        if (lane == 0) { hA_up = 0; hB_up = 0; fA = 0; fB = 0; b_in = 0; }
#ifndef NOFLAG
        if (o1a < 0) {
            o1a &= 0x7fff;
            b_out = __vmaxs2(best, b_in);
            if (lane == 31) { if (n31 > 0) out[n31] = b_out; ++n31; }
            best = 0; diag_in = 0;
#pragma unroll
            for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
        }
#endif
        unsigned wa1[KW], wa2[KW], wb1[KW], wb2[KW];
#pragma unroll
        for (int m = 0; m < KW; ++m) {
            wa1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1a + m * 128);
            wa2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2a + m * 128);
            wb1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1b + m * 128);
            wb2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2b + m * 128);
        }
        unsigned tA = __vadd2(diag_in, prmt(wa1[0], wa2[0], 0xC480u));
        unsigned tB = __vadd2(hA_up, prmt(wb1[0], wb2[0], 0xC480u));
        diag_in = hB_up;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            unsigned tA_next = 0, tB_next = 0;
            const unsigned sel = ((r + 1) & 3) == 0 ? 0xC480u : ((r + 1) & 3) == 1 ? 0xD591u : ((r + 1) & 3) == 2 ? 0xE6A2u : 0xF7B3u;
            if (r + 1 < K) tA_next = __vadd2(H[r], prmt(wa1[(r + 1) >> 2], wa2[(r + 1) >> 2], sel));
            const unsigned hA = __vimax3_s16x2_relu(tA, E[r], fA);
            const unsigned hqA = __vadd2(hA, negQ);
            const unsigned eB = __viaddmax_s16x2(E[r], negR, hqA);
            fA = __viaddmax_s16x2(fA, negR, hqA);
            if (r + 1 < K) tB_next = __vadd2(hA, prmt(wb1[(r + 1) >> 2], wb2[(r + 1) >> 2], sel));
            const unsigned hB = __vimax3_s16x2_relu(tB, eB, fB);
            H[r] = hB;
            const unsigned hqB = __vadd2(hB, negQ);
            E[r] = __viaddmax_s16x2(eB, negR, hqB);
            fB = __viaddmax_s16x2(fB, negR, hqB);
            best = __vimax3_s16x2(best, tA, tB);
            if (r == K - 1) hA_last = hA;
            tA = tA_next; tB = tB_next;
        }
        hB_last = H[K - 1];
        fA_out = fA; fB_out = fB;
    }
    out[lane] = best + n31;
}
template <int K>
__global__ void __launch_bounds__(256, 2) kern(const unsigned* prof, const unsigned short* ring, int nsteps, unsigned negQ, unsigned negR, int* out) {
    extern __shared__ unsigned smem[];
    for (int i = threadIdx.x; i < 27 * 8 * 32; i += blockDim.x) smem[i] = prof[i];
    unsigned short* r = reinterpret_cast<unsigned short*>(smem + 27 * 8 * 32);
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) r[i] = ring[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if VARIANT == 3
    body2<K>(smem + lane, r + warp * 256, r + warp * 256 + 128, nsteps, negQ, negR, out + blockIdx.x * 256 + warp * 32, lane);
#else
    body<K>(smem + lane, r + warp * 256, r + warp * 256 + 128, nsteps, negQ, negR, out + blockIdx.x * 256 + warp * 32, lane);
#endif
}
template __global__ void kern<17>(const unsigned*, const unsigned short*, int, unsigned, unsigned, int*);
template __global__ void kern<32>(const unsigned*, const unsigned short*, int, unsigned, unsigned, int*);
template __global__ void kern<8>(const unsigned*, const unsigned short*, int, unsigned, unsigned, int*);
#include <cstdio>
#include <vector>
#include <cstdlib>
template <int K>
static void run(const char* name, int ctas_per_sm, int nsteps, int cols_per_step) {
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int KW = (K + 3) / 4;
    std::vector<unsigned> prof(27 * 8 * 32);
    for (auto& x : prof) x = (unsigned)rand() * 2654435761u;
    std::vector<unsigned short> ring(8 * 256);
    for (auto& x : ring) x = (unsigned short)((rand() % 26) * KW * 128);
    unsigned *d_prof; unsigned short* d_ring; int* d_out;
    cudaMalloc(&d_prof, prof.size() * 4); cudaMalloc(&d_ring, ring.size() * 2); cudaMalloc(&d_out, 4 * 256 * 148 * 8);
    cudaMemcpy(d_prof, prof.data(), prof.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ring, ring.data(), ring.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 27 * 8 * 32 * 4 + 8 * 256 * 2;
    cudaFuncSetAttribute(kern<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<K>, 256, smem);
    const int grid = pr.multiProcessorCount * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<K><<<grid, 256, smem>>>(d_prof, d_ring, nsteps, 0xfff6fff6u, 0xffffffffu, d_out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<K><<<grid, 256, smem>>>(d_prof, d_ring, nsteps, 0xfff6fff6u, 0xffffffffu, d_out);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cellpairs = (double)grid * 256 * (double)nsteps * K * cols_per_step;   // per lane
    const double laneops = cellpairs * 4.5;
    printf("{\"variant\": \"%s\", \"K\": %d, \"ctas_per_sm\": %d, \"occupancy_limit\": %d, \"ms\": %.3f, \"gcups\": %.1f, \"alu_lane_ops_per_s_at_4.5\": %.4e, \"err\": \"%s\"}\n",
           name, K, ctas_per_sm, occ, ms, cellpairs * 2 / ms / 1e6, laneops / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_prof); cudaFree(d_ring); cudaFree(d_out);
}
int main() {
#if VARIANT == 3
    const int cps = 2; const char* nm = "two_columns";
#else
    const int cps = 1; const char* nm = "one_column";
#endif
    run<32>(nm, 2, 20000 / cps, cps); run<17>(nm, 2, 40000 / cps, cps); run<8>(nm, 2, 80000 / cps, cps);
    run<17>(nm, 1, 40000 / cps, cps); run<8>(nm, 1, 80000 / cps, cps);
    return 0;
}
