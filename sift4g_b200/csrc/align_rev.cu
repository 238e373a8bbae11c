// Stage 3, step 2: BEGIN cells of the kept hits with the packed s16x2 systolic sweep, two hits per warp (sm_100a).
//
// SSW finds the begin cell of an alignment by sweeping the reversed query prefix q[q_end..0] against the reversed target
// prefix t[t_end..0] and taking the first column, then the first row in it, whose H equals the score (vendor/swsharp/
// swsharp/src/ssw/ssw.c:296,500,827-838).  The 32-bit sweep of align.cu does that one hit per warp; here two hits ride in the
// halves of s16x2 registers like in the score kernel (sw_score.cu) -- the cell update is the same five DPX/ALU instructions.
//
// What differs from the forward kernels is the query profile: two hits of a query end at different query rows, so their
// reversed rows are different rows of the query.  The CTA builds ONE profile of the reversed query from the row
// R = min(32 K, qlen) - 1 downwards (K rows per lane; hits are grouped by query and by the K class of their end row), and
// a hit that ends `off` rows above R reads it `off` rows further on: every lane's block of K rows is stored with the K + 4
// rows that follow it (block stride 2 K/4 + 1 words: odd, so the 32 lanes hit 32 banks whatever the shift), which turns the
// shift into a per-hit base pointer -- whole lanes and words go into the address, the last 0..3 bytes are funnelled with one
// PRMT per profile word.  Cost: K/4 + 1 loads and K/4 PRMT per half and step on top of the forward kernel's, no masks, no
// dead rows.
//
// Work: the hits are sorted by (query, K class, columns descending) and cut into tiles of 128 sorted hits; a CTA takes a
// tile group by group (profile build per group), its warps pull pairs of neighbouring hits.  Hits that need more than 1024
// rows, hits under the swAlign rules and hits without an end cell are left to al_sweep32_kernel (coords stay -1).
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kRing = 32;                         // columns per refill = granularity of the early stop (the wavefront needs 2 x 32 ring entries)
constexpr int kTileHits = 128;
constexpr unsigned kNotFound = 0xffffffffu;
constexpr int kBlocks = 33;                       // 32 lanes + one all-pad block for lanes shifted beyond the query start
constexpr size_t kProfBytes = (((size_t)(S4G_PAD_CODE + 1) * kBlocks * (2 * 8 + 1) * 4 + 15) / 16) * 16;      // profile of the K = 32 class
constexpr size_t kRevSmem = kProfBytes + kTileHits * (sizeof(unsigned long long) + sizeof(uint32_t)) + kWarps * 4 * kRing * sizeof(unsigned short) +
                            (S4G_PAD_CODE + 1) * 32;

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

struct RevParams {
    const uint8_t* db_codes;
    const int64_t* db_off;
    uint32_t id_base;
    const uint8_t* q_codes;
    const int64_t* q_off;
    const uint32_t* pair_q;
    const uint32_t* pair_t;
    const int32_t* pair_score;
    const int8_t* mat8;
    int32_t go, ge;
    int32_t* coords;
    const int32_t* q_rank;                // position of a query in the descending-length order: the key's query field (long queries' tiles first)
    const int32_t* q_order;               // and back
    const unsigned long long* keys;       // sorted
    const uint32_t* vals;                 // hit index of every sorted key
    const unsigned long long* n_valid;    // sorted keys that are real work (the rest carry ~0)
    unsigned long long* counters;         // [0] tile cursor
    unsigned long long* flags;            // bit 0: a score was not attained
};

struct RevHit {
    const uint8_t* t_end_ptr;     // &target[t_end]
    int ncols;                    // t_end + 1
    int q_end;
    unsigned score;
};

// profile words of one group: prof[letter][block][x], block b holds reversed rows b K .. b K + 2 K + 3 (4 per word)
template <int K>
__device__ void build_profile_rev(unsigned* prof, const int8_t* smat, const uint8_t* q, int R) {
    constexpr int KW = K / 4, BS = 2 * KW + 1, LS = kBlocks * BS;
    for (int w = threadIdx.x; w < (S4G_PAD_CODE + 1) * LS; w += blockDim.x) {
        const int letter = w / LS, rem = w - letter * LS, blk = rem / BS, x = rem - blk * BS;
        unsigned word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int rho = blk * K + 4 * x + b;
            const int v = rho <= R ? (int)smat[letter * 32 + q[R - rho]] : 0;
            word |= (unsigned)(v & 0xff) << (8 * b);
        }
        prof[w] = word;
    }
}

template <int K>
__device__ __forceinline__ void reverse_pair_packed(const unsigned* prof, int R, unsigned short* ring1, unsigned short* ring2, const RevHit& h1,
                                                    const RevHit& h2, unsigned negQ, unsigned negR, int lane, unsigned* found1, unsigned* found2) {
    constexpr int KW = K / 4, BS = 2 * KW + 1, LS = kBlocks * BS;
    constexpr unsigned kLetterBytes = LS * 4;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kLetterBytes;
    constexpr int kRingMask = 2 * kRing - 1;
    static_assert(kPadOff < 65536, "profile offsets must fit the 16-bit ring");
    const unsigned FULL = 0xffffffffu;

    // per hit: block and byte shift of this lane's rows inside the group's profile
    const int off1 = R - h1.q_end, off2 = R - h2.q_end;
    const int blk1 = min(lane + off1 / K, kBlocks - 1), blk2 = min(lane + off2 / K, kBlocks - 1);
    const int b1 = off1 % K, b2 = off2 % K;
    const char* p1 = reinterpret_cast<const char*>(prof) + (blk1 * BS + (b1 >> 2)) * 4;
    const char* p2 = reinterpret_cast<const char*>(prof) + (blk2 * BS + (b2 >> 2)) * 4;
    const unsigned sel1 = 0x3210u + 0x1111u * (unsigned)(b1 & 3), sel2 = 0x3210u + 0x1111u * (unsigned)(b2 & 3);
    const unsigned score2 = (h1.score & 0xffffu) | (h2.score << 16);

    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0;
    unsigned fnd1 = kNotFound, fnd2 = kNotFound;
    const int len1 = h1.ncols, len2 = h2.ncols;
    const int maxlen = len1 > len2 ? len1 : len2;
    const int nsteps = maxlen + 31;
    for (int c = lane; c < kRing; c += 32) { ring1[kRing + c] = kPadOff; ring2[kRing + c] = kPadOff; }

    for (int s0 = 0; s0 < nsteps; s0 += kRing) {
        if (s0 > 0) {
            // a hit is settled once every lane has passed the column of its first cell with the score (or its last column)
            unsigned m1 = fnd1, m2 = fnd2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { m1 = min(m1, __shfl_xor_sync(FULL, m1, o)); m2 = min(m2, __shfl_xor_sync(FULL, m2, o)); }
            const bool d1 = (m1 != kNotFound && (int)(m1 >> 10) + 32 <= s0) || len1 + 31 <= s0;
            const bool d2 = (m2 != kNotFound && (int)(m2 >> 10) + 32 <= s0) || len2 + 31 <= s0;
            if (d1 && d2) break;
        }
        {
            const int base = s0 & kRingMask;
#pragma unroll
            for (int c = 0; c < kRing; c += 32) {
                const int j = s0 + c + lane;                    // reversed column j = target position t_end - j
                const unsigned o1 = j < len1 ? (unsigned)h1.t_end_ptr[-j] * kLetterBytes : kPadOff;
                const unsigned o2 = j < len2 ? (unsigned)h2.t_end_ptr[-j] * kLetterBytes : kPadOff;
                ring1[base + c + lane] = (unsigned short)o1;
                ring2[base + c + lane] = (unsigned short)o2;
            }
        }
        __syncwarp();
        const int send = (nsteps - s0) < kRing ? (nsteps - s0) : kRing;
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int j = (s0 + ss - lane) & kRingMask;
            const unsigned o1 = ring1[j];
            const unsigned o2 = ring2[j];
            unsigned w1[KW], w2[KW];
            {
                unsigned x1[KW + 1], x2[KW + 1];
#pragma unroll
                for (int m = 0; m <= KW; ++m) {
                    x1[m] = *reinterpret_cast<const unsigned*>(p1 + o1 + m * 4);
                    x2[m] = *reinterpret_cast<const unsigned*>(p2 + o2 + m * 4);
                }
#pragma unroll
                for (int m = 0; m < KW; ++m) { w1[m] = prmt(x1[m], x1[m + 1], sel1); w2[m] = prmt(x2[m], x2[m + 1], sel2); }
            }
            unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
            unsigned f = __shfl_up_sync(FULL, f_out, 1);
            if (lane == 0) { h_up = 0; f = 0; }
            unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;       // cell update: see score_pair_packed (sw_score.cu)
            diag_in = h_up;
            best = 0;
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
            const unsigned eq = __vcmpeq2(best, score2);
            if (eq) {
                const int col = s0 + ss - lane;
                if ((eq & 0xffffu) && col >= 0 && col < len1) {
                    int rr = -1;
#pragma unroll
                    for (int r = K - 1; r >= 0; --r) if ((H[r] & 0xffffu) == (score2 & 0xffffu) && lane * K + r <= h1.q_end) rr = r;
                    if (rr >= 0) fnd1 = min(fnd1, ((unsigned)col << 10) | (unsigned)(lane * K + rr));
                }
                if ((eq >> 16) && col >= 0 && col < len2) {
                    int rr = -1;
#pragma unroll
                    for (int r = K - 1; r >= 0; --r) if ((H[r] >> 16) == (score2 >> 16) && lane * K + r <= h2.q_end) rr = r;
                    if (rr >= 0) fnd2 = min(fnd2, ((unsigned)col << 10) | (unsigned)(lane * K + rr));
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { fnd1 = min(fnd1, __shfl_xor_sync(FULL, fnd1, o)); fnd2 = min(fnd2, __shfl_xor_sync(FULL, fnd2, o)); }
    *found1 = fnd1; *found2 = fnd2;
}

// sort key of a hit: query, as its position in the descending-length order (bits 63..35; the long queries' tiles go first) | K class - 1 (34..32) | ~columns (31..0); hits this kernel does not take: ~0
__device__ __forceinline__ int key_kclass(unsigned long long key) { return (int)((key >> 32) & 7u) + 1; }     // rows per lane / 4

template <int K>
__device__ void run_group(const RevParams& P, unsigned* prof, const int8_t* smat, unsigned short* rings, int* s_next, const unsigned long long* s_keys,
                          const uint32_t* s_vals, int g0, int g1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = (uint32_t)P.q_order[(uint32_t)(s_keys[g0] >> 35)];
    const int64_t qo = P.q_off[q];
    const int qlen = (int)(P.q_off[q + 1] - qo);
    const int R = min(32 * K, qlen) - 1;
    __syncthreads();                                   // the previous group's profile is no longer read
    build_profile_rev<K>(prof, smat, P.q_codes + qo, R);
    if (threadIdx.x == 0) *s_next = 0;
    __syncthreads();
    unsigned short* ring1 = rings + warp * (4 * kRing);
    unsigned short* ring2 = ring1 + 2 * kRing;
    const unsigned negQ = ((unsigned)(-P.go) & 0xffffu) * 0x10001u;
    const unsigned negR = ((unsigned)(-P.ge) & 0xffffu) * 0x10001u;
    const int n_pairs = (g1 - g0 + 1) >> 1;
    while (true) {
        int p = 0;
        if (lane == 0) p = atomicAdd(s_next, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= n_pairs) break;
        const int i1 = g0 + 2 * p, i2 = i1 + 1;
        const bool has2 = i2 < g1;
        const uint32_t c1 = s_vals[i1], c2 = has2 ? s_vals[i2] : c1;
        RevHit h1, h2;
        {
            const uint32_t t1 = P.pair_t[c1] - P.id_base, t2 = P.pair_t[c2] - P.id_base;
            const int te1 = P.coords[4 * (int64_t)c1 + 3], te2 = P.coords[4 * (int64_t)c2 + 3];
            h1.t_end_ptr = P.db_codes + P.db_off[t1] + te1; h1.ncols = te1 + 1; h1.q_end = P.coords[4 * (int64_t)c1 + 1]; h1.score = (unsigned)P.pair_score[c1];
            h2.t_end_ptr = P.db_codes + P.db_off[t2] + te2; h2.ncols = has2 ? te2 + 1 : 0; h2.q_end = P.coords[4 * (int64_t)c2 + 1];
            h2.score = has2 ? (unsigned)P.pair_score[c2] : 0x7fffu;        // missing second hit: a score no cell can reach
        }
        unsigned f1, f2;
        reverse_pair_packed<K>(prof, R, ring1, ring2, h1, h2, negQ, negR, lane, &f1, &f2);
        if (lane == 0) {
            if (f1 == kNotFound) atomicOr(P.flags, 1ull);
            else { P.coords[4 * (int64_t)c1 + 0] = h1.q_end - (int)(f1 & 1023u); P.coords[4 * (int64_t)c1 + 2] = h1.ncols - 1 - (int)(f1 >> 10); }
            if (has2) {
                if (f2 == kNotFound) atomicOr(P.flags, 1ull);
                else { P.coords[4 * (int64_t)c2 + 0] = h2.q_end - (int)(f2 & 1023u); P.coords[4 * (int64_t)c2 + 2] = h2.ncols - 1 - (int)(f2 >> 10); }
            }
        }
    }
}

__global__ void __launch_bounds__(kWarps * 32, 2) al_reverse_packed_kernel(RevParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* prof = reinterpret_cast<unsigned*>(smem);
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem + kProfBytes);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + kTileHits);
    unsigned short* rings = reinterpret_cast<unsigned short*>(s_vals + kTileHits);
    int8_t* smat = reinterpret_cast<int8_t*>(rings + kWarps * 4 * kRing);
    __shared__ int s_next;
    __shared__ long long s_tile;
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    const long long n_valid = (long long)*P.n_valid;
    const long long n_tiles = (n_valid + kTileHits - 1) / kTileHits;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(&P.counters[0], 1ull);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= n_tiles) break;
        const long long h0 = tile * kTileHits;
        const int n = (int)min((long long)kTileHits, n_valid - h0);
        for (int i = threadIdx.x; i < n; i += blockDim.x) { s_keys[i] = P.keys[h0 + i]; s_vals[i] = P.vals[h0 + i]; }
        __syncthreads();
        // groups of equal (query, K class), in order
        int g0 = 0;
        while (g0 < n) {
            const unsigned long long gk = s_keys[g0] >> 32;
            int g1 = g0 + 1;
            while (g1 < n && (s_keys[g1] >> 32) == gk) ++g1;
            switch (key_kclass(s_keys[g0])) {
                case 1: run_group<4>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                case 2: run_group<8>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                case 3: run_group<12>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                case 4: run_group<16>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                case 5: run_group<20>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                case 6: run_group<24>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                case 7: run_group<28>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
                default: run_group<32>(P, prof, smat, rings, &s_next, s_keys, s_vals, g0, g1); break;
            }
            g0 = g1;
        }
    }
}

__global__ void rev_keys_kernel(RevParams P, int64_t n, int swalign_all, unsigned long long* keys, uint32_t* vals, unsigned long long* n_valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q_end = P.coords[4 * i + 1], t_end = P.coords[4 * i + 3], score = P.pair_score[i];
    unsigned long long key = ~0ull;
    if (!swalign_all && score <= 32767 && score > 0 && q_end >= 0 && t_end >= 0 && q_end < 1024) {
        const unsigned kc = (unsigned)(q_end / 128);                   // rows per lane = 4 (kc + 1) covers q_end + 1 rows
        key = ((unsigned long long)(uint32_t)P.q_rank[P.pair_q[i]] << 35) | ((unsigned long long)kc << 32) | (unsigned long long)(0xffffffffu - (unsigned)(t_end + 1));
        atomicAdd(n_valid, 1ull);
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

}  // namespace

// Begin cells (coords[4i+0], coords[4i+2]) of the hits whose end cells are in coords[4i+1], coords[4i+3]; hits it does not
// take keep -1 there.  d_flags: bit 0 set when a score is not attained.
int s4g_sw_reverse_begins_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n, const uint32_t* d_pair_q, const uint32_t* d_pair_t,
                                 const int32_t* d_pair_score, const int8_t* d_mat8, int gap_open, int gap_extend, int swalign_all,
                                 int32_t* d_coords, unsigned long long* d_flags) {
    cudaStream_t st = ctx->stream;
    if (n == 0) return S4G_OK;
    unsigned long long* d_keys = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_KEYS, sizeof(unsigned long long) * n);
    unsigned long long* d_keys2 = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_KEYS2, sizeof(unsigned long long) * n);
    uint32_t* d_vals = (uint32_t*)s4g_scratch(ctx, SLOT_SW_VALS, sizeof(uint32_t) * n);
    uint32_t* d_vals2 = (uint32_t*)s4g_scratch(ctx, SLOT_SW_VALS2, sizeof(uint32_t) * n);
    unsigned long long* d_counters = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_MISC, 64);
    if (!d_keys || !d_keys2 || !d_vals || !d_vals2 || !d_counters) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 64, st));
    RevParams P;
    P.db_codes = db->d_codes; P.db_off = db->d_off; P.id_base = db->id_base;
    P.q_codes = q->d_codes; P.q_off = q->d_off;
    P.pair_q = d_pair_q; P.pair_t = d_pair_t; P.pair_score = d_pair_score; P.mat8 = d_mat8;
    P.go = gap_open; P.ge = gap_extend; P.coords = d_coords;
    P.q_rank = q->d_len_rank; P.q_order = q->d_len_order;
    P.keys = d_keys2; P.vals = d_vals2; P.n_valid = d_counters + 1; P.counters = d_counters; P.flags = d_flags;
    rev_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, n, swalign_all, d_keys, d_vals, d_counters + 1);
    S4G_CHECK_LAUNCH(ctx);
    {
        int qbits = 1;
        while ((1ll << qbits) < q->n) ++qbits;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 64, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp_bytes);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 64, st));
        (void)qbits;
        ctx->launches += 4;
    }
    const size_t smem = kRevSmem;
    S4G_CUDA(ctx, cudaFuncSetAttribute(al_reverse_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, al_reverse_packed_kernel, kWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    al_reverse_packed_kernel<<<ctx->sm_count * per_sm, kWarps * 32, smem, st>>>(P);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}
