// Stage 3: traceback (alignment paths) of the kept hits (sm_100a).
//
// Replaces alignScoredPair -> alignScoredPairCpu -> SSW (vendor/swsharp/swsharp/src/align.c:235-255,
// cpu_module.c:111-142, sse_module.c:62-122,178-265, ssw/ssw.c:771-856) for score <= 32767.  The result
// must be the reference's, not just "an" optimal alignment, so the three SSW steps are reproduced rule
// for rule:
//   1. end cell   : among cells with H == score the smallest target column, then the smallest query row
//                   (ssw.c:283-308,491-512);
//   2. begin cell : the same sweep over the reversed query prefix / target prefix scanned right to left;
//                   first column holding `score`, smallest reversed row (ssw.c:296,500,827-838);
//   3. path       : banded_sw on the sub-rectangle, band |dt - dq| + 1 doubled until the banded maximum
//                   reaches the score; direction rules and band-edge behaviour of ssw.c:549-727.
// Step 1 (end cells): packed s16x2 systolic sweep, two hits of a query per warp (al_forward_packed_kernel, sw_score.cu);
// the 32-bit sweep below (al_sweep32_kernel, one hit per warp, lanes own 8 query rows, 256-row passes) takes the hits of
// queries beyond the packed kernel's reach and the swAlign-rule hits.  Step 2 (begin cells): packed reverse sweep, two
// hits of a query per warp sharing one profile (al_reverse_packed_kernel); al_sweep32_kernel for what it cannot take.
// Step 3 (path): al_band_persistent_kernel -- one warp takes a hit through all its band-doubling attempts, rows
// warp-parallel (max-plus prefix scan for F), direction bytes in a per-warp scratch region; bands that outgrow its
// shared-memory rows go to the host-sequenced rounds (al_band_warp_kernel; al_band_kernel, one thread per hit, for
// gap_extend > gap_open, where the scan does not apply).  Score > 32767 or gap penalties beyond 8 bits: al_swalign_kernel.
#include <algorithm>
#include <cstring>
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;

struct AlParams {
    const uint8_t* db_codes;
    const int64_t* db_off;
    uint32_t id_base;
    const uint8_t* q_codes;
    const int64_t* q_off;
    const uint32_t* pair_q;
    const uint32_t* pair_t;
    const int32_t* pair_score;
    int64_t n_pairs;
    const int8_t* mat8;          // [target letter 27][query letter 32]
    int32_t go, ge;
    int32_t* coords;             // 4 per pair
    unsigned long long* counters;   // [0] work cursor, [1] error flags
    int32_t* bound;              // per-warp boundary rows for multi-pass sweeps
    int64_t bound_stride;
    int32_t swalign_all;         // gap penalties outside SSW's range: every hit takes the swAlign rules (sse_module.c:215-236)
};

// which engine defines the path of a hit: SSW (score <= 32767, sse_module.c:181-185) or swsharp's swAlign
__device__ __forceinline__ bool takes_swalign(const AlParams& P, int score) { return P.swalign_all || score > 32767; }

// 32-bit systolic sweep, one warp per hit: Gotoh local alignment of rows i -> q[i * qstep] (i < rows) against
// columns j -> t[j * tstep] (j < cols); finds the lexicographically smallest (column, row) whose H equals `score`
// and stops once every lane has passed that column.  Lanes own 8 consecutive rows (256 rows per pass; the boundary
// row travels between passes through a per-warp global scratch, staged through shared memory 64 columns at a time).
// The lane's rows x 27 letters substitution scores sit in a per-warp shared-memory profile (4 int8 per word,
// conflict free), rebuilt per pass.  Per cell: IADD (Hdiag + S), VIMNMX3.RELU, IADD (H - Q), 2 x VIADDMNMX.
constexpr int kSwWarps = 8;
constexpr int kSwRows = 8;
constexpr int kSwRing = 64;
constexpr int kSwProfWords = (S4G_PAD_CODE + 1) * 2 * 32;
constexpr int kSwWarpBytes = kSwProfWords * 4 + 2 * kSwRing * 2 + 2 * kSwRing * 4 + 4 * kSwRing * 4;
constexpr int kSwSmatBytes = 2 * 896;      // substitution matrix and its transpose

struct SweepSmem {
    unsigned* prof;          // [27][2][32]
    unsigned short* ring;    // [128] byte offset of the column letter's profile row
    int *inH, *inF;          // [64] boundary row of the previous pass for the current block of columns
    int *outH, *outF;        // [128] boundary row produced by lane 31
};

__device__ unsigned long long sweep32(const SweepSmem& S, const int8_t* smat, const uint8_t* q, int qstep, int rows, const uint8_t* t,
                                      int tstep, int cols, int score, int Q, int R, int32_t* bH, int32_t* bF, int lane) {
    const unsigned FULL = 0xffffffffu;
    unsigned long long found = ~0ull;
    int jlim = cols;                                   // columns [0, jlim) can still hold the answer
    const int npass = (rows + 32 * kSwRows - 1) / (32 * kSwRows);
    const char* prof_b = reinterpret_cast<const char*>(S.prof) + lane * 4;
    for (int pass = 0; pass < npass; ++pass) {
        const int row0 = pass * 32 * kSwRows + lane * kSwRows;
        const bool first = pass == 0, last = pass == npass - 1;
        // profile of this lane's rows
        {
            int ql[kSwRows];
#pragma unroll
            for (int r = 0; r < kSwRows; ++r) ql[r] = row0 + r < rows ? (int)q[(long long)(row0 + r) * qstep] : -1;
            for (int letter = 0; letter <= S4G_PAD_CODE; ++letter) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    unsigned word = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int v = ql[4 * m + b] >= 0 ? (int)smat[letter * 32 + ql[4 * m + b]] : -128;
                        word |= (unsigned)(v & 0xff) << (8 * b);
                    }
                    S.prof[(letter * 2 + m) * 32 + lane] = word;
                }
            }
        }
        for (int c = lane; c < kSwRing; c += 32) S.ring[kSwRing + c] = (unsigned short)(S4G_PAD_CODE * 256);   // columns -64..-1
        int H[kSwRows], E[kSwRows];
#pragma unroll
        for (int r = 0; r < kSwRows; ++r) { H[r] = 0; E[r] = 0; }
        int h_last = 0, f_out = 0, diag_in = 0;
        const int nsteps = jlim + 31;
        int flushed = 0;
        for (int s0 = 0; s0 < nsteps; s0 += kSwRing) {
            // boundary row of the columns every lane has finished -> global (for the next pass)
            if (!last) {
                const int upto = min(jlim, s0 - 31);
                for (int c = flushed + lane; c < upto; c += 32) { bH[c] = S.outH[c & (2 * kSwRing - 1)]; bF[c] = S.outF[c & (2 * kSwRing - 1)]; }
                if (upto > flushed) flushed = upto;
            }
            if (s0 > 0) {
                unsigned long long fm = found;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(FULL, fm, o); if (x < fm) fm = x; }
                if (fm != ~0ull && (int)(fm >> 32) + 32 <= s0) break;
            }
            {
                const int base = s0 & (2 * kSwRing - 1);
#pragma unroll
                for (int c = 0; c < kSwRing; c += 32) {
                    const int j = s0 + c + lane;
                    S.ring[base + c + lane] = (unsigned short)((j < jlim ? (unsigned)t[(long long)j * tstep] : (unsigned)S4G_PAD_CODE) * 256u);
                    if (!first) { S.inH[c + lane] = j < jlim ? bH[j] : 0; S.inF[c + lane] = j < jlim ? bF[j] : 0; }
                }
            }
            __syncwarp();
            const int send = (nsteps - s0) < kSwRing ? (nsteps - s0) : kSwRing;
#pragma unroll 1
            for (int ss = 0; ss < send; ++ss) {
                const int j = s0 + ss - lane;
                const unsigned o = S.ring[j & (2 * kSwRing - 1)];
                const unsigned w0 = *reinterpret_cast<const unsigned*>(prof_b + o);
                const unsigned w1 = *reinterpret_cast<const unsigned*>(prof_b + o + 128);
                int h_up = __shfl_up_sync(FULL, h_last, 1);
                int f = __shfl_up_sync(FULL, f_out, 1);
                if (lane == 0) { h_up = first ? 0 : S.inH[ss]; f = first ? 0 : S.inF[ss]; }
                int tt = diag_in + (int)(int8_t)(w0 & 0xffu), t_prev = 0, cm = 0;
                diag_in = h_up;
#pragma unroll
                for (int r = 0; r < kSwRows; ++r) {
                    int t_next = 0;
                    if (r + 1 < kSwRows) {
                        const unsigned w = (r + 1) < 4 ? w0 : w1;
                        t_next = H[r] + (int)(int8_t)((w >> (8 * ((r + 1) & 3))) & 0xffu);
                    }
                    const int h = __vimax3_s32_relu(tt, E[r], f);
                    H[r] = h;
                    const int hq = h - Q;
                    E[r] = __viaddmax_s32(E[r], -R, hq);
                    f = __viaddmax_s32(f, -R, hq);
                    if (r & 1) cm = __vimax3_s32(cm, t_prev, tt);        // the maximum is reached by a diagonal step
                    t_prev = tt;
                    tt = t_next;
                }
                h_last = H[kSwRows - 1];
                f_out = f;
                if (lane == 31 && !last && j >= 0 && j < jlim) { S.outH[j & (2 * kSwRing - 1)] = h_last; S.outF[j & (2 * kSwRing - 1)] = f_out; }
                if (cm == score && j >= 0 && j < jlim) {
                    int rr = -1;
#pragma unroll
                    for (int r = kSwRows - 1; r >= 0; --r) if (H[r] == score && row0 + r < rows) rr = r;
                    if (rr >= 0) {
                        const unsigned long long cand = ((unsigned long long)(unsigned)j << 32) | (unsigned)(row0 + rr);
                        if (cand < found) found = cand;
                    }
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(FULL, found, o); if (x < found) found = x; }
        if (found != ~0ull) jlim = (int)(found >> 32) + 1;
        if (!last) {
            for (int c = flushed + lane; c < jlim; c += 32) { bH[c] = S.outH[c & (2 * kSwRing - 1)]; bF[c] = S.outF[c & (2 * kSwRing - 1)]; }
        }
        __syncwarp();
    }
    return found;
}

// mode 0: begin cells of all hits (reverse sweep from the end cell; coords[1], coords[3] must be set)
// mode 1: end cells of the hits whose query is longer than `long_rows` (forward sweep; the packed kernel did the rest)
__global__ void __launch_bounds__(kSwWarps * 32) al_sweep32_kernel(AlParams P, int mode, int long_rows, unsigned long long* cursor) {
    extern __shared__ __align__(16) unsigned char ssm[];
    int8_t* smat = reinterpret_cast<int8_t*>(ssm);            // [target letter][query letter]
    int8_t* smatT = smat + 896;                               // [query letter][target letter], pad row copied
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) {
        smat[i] = P.mat8[i];
        const int a = i >> 5, b = i & 31;
        smatT[i] = a == S4G_PAD_CODE ? P.mat8[i] : (b <= S4G_PAD_CODE ? P.mat8[b * 32 + a] : 0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wb = ssm + kSwSmatBytes + warp * kSwWarpBytes;
    SweepSmem S;
    S.prof = reinterpret_cast<unsigned*>(wb);
    S.ring = reinterpret_cast<unsigned short*>(S.prof + kSwProfWords);
    S.inH = reinterpret_cast<int*>(S.ring + 2 * kSwRing);
    S.inF = S.inH + kSwRing;
    S.outH = S.inF + kSwRing;
    S.outF = S.outH + 2 * kSwRing;
    const int gwarp = blockIdx.x * kSwWarps + warp;
    int32_t* bH = P.bound + (int64_t)gwarp * P.bound_stride;
    int32_t* bF = bH + P.bound_stride / 2;
    const unsigned FULL = 0xffffffffu;
    while (true) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(cursor, 1ull);
        w = __shfl_sync(FULL, w, 0);
        if ((long long)w >= P.n_pairs) break;
        const uint32_t qi = P.pair_q[w], ti = P.pair_t[w] - P.id_base;
        const uint8_t* q = P.q_codes + P.q_off[qi];
        const int qlen = (int)(P.q_off[qi + 1] - P.q_off[qi]);
        const uint8_t* t = P.db_codes + P.db_off[ti];
        const int tlen = (int)(P.db_off[ti + 1] - P.db_off[ti]);
        const int score = P.pair_score[w];
        if (mode == 1) {
            const bool swa = takes_swalign(P, score);
            if (qlen <= long_rows && !swa) continue;
            if (!swa && P.coords[4 * w + 1] >= 0) continue;      // settled by the striped end-cell sweep (sw_score.cu)
            if (swa) {
                // swAlign's end cell: the first cell in query-major order that reaches the maximum (cpu_module.c:1320-1325)
                // = the SSW rule on the transposed problem (rows = target, columns = query)
                const unsigned long long e = score > 0 ? sweep32(S, smatT, t, 1, tlen, q, 1, qlen, score, P.go, P.ge, bH, bF, lane) : ~0ull;
                if (lane == 0) {
                    if (e == ~0ull) atomicOr(P.counters + 1, 1ull);
                    P.coords[4 * w + 1] = e == ~0ull ? -1 : (int)(e >> 32);
                    P.coords[4 * w + 3] = e == ~0ull ? -1 : (int)(e & 0xffffffffu);
                }
                __syncwarp();
                continue;
            }
            const unsigned long long e = score > 0 ? sweep32(S, smat, q, 1, qlen, t, 1, tlen, score, P.go, P.ge, bH, bF, lane) : ~0ull;
            if (lane == 0) {
                if (e == ~0ull) atomicOr(P.counters + 1, 1ull);
                P.coords[4 * w + 1] = e == ~0ull ? -1 : (int)(e & 0xffffffffu);
                P.coords[4 * w + 3] = e == ~0ull ? -1 : (int)(e >> 32);
            }
        } else {
            if (takes_swalign(P, score)) continue;
            if (P.coords[4 * w + 0] >= 0) continue;              // the packed reverse sweep has settled this hit
            const int q_end = P.coords[4 * w + 1], t_end = P.coords[4 * w + 3];
            unsigned long long b = ~0ull;
            if (q_end >= 0 && t_end >= 0) b = sweep32(S, smat, q + q_end, -1, q_end + 1, t + t_end, -1, t_end + 1, score, P.go, P.ge, bH, bF, lane);
            if (lane == 0) {
                if (b == ~0ull) atomicOr(P.counters + 1, 1ull);
                P.coords[4 * w + 0] = b == ~0ull ? -1 : q_end - (int)(b & 0xffffffffu);
                P.coords[4 * w + 2] = b == ~0ull ? -1 : t_end - (int)(b >> 32);
            }
        }
        __syncwarp();
    }
}

// ---- banded traceback ---------------------------------------------------------------------------------------

struct BandWork {
    uint32_t pair;       // index into the pair arrays
    int32_t w;           // band half-width of this round
    int64_t dir_off;     // offset of this hit's direction bytes
};

struct BandParams {
    const BandWork* work;
    int32_t n_work;
    int32_t stride;              // ints per band row in shared memory (odd, >= 2*w_max + 4)
    uint8_t* dir;
    uint8_t* rev_paths;          // per pair slot: reversed path ops
    const int64_t* slot_off;     // n_pairs + 1
    int32_t* path_len;           // per pair, -1 while pending
    int32_t* status;             // per work item: 1 done, 0 needs a wider band, <0 error
};

__device__ __forceinline__ int slot_of(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x + 1; }

__global__ void al_band_kernel(AlParams P, BandParams B) {
    extern __shared__ int32_t rows[];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= B.n_work) return;
    const BandWork wk = B.work[tid];
    int32_t* hb = rows + (size_t)threadIdx.x * 3 * B.stride;
    int32_t* eb = hb + B.stride;
    int32_t* hc = eb + B.stride;
    const uint32_t p = wk.pair;
    const int q0 = P.coords[4 * p + 0], q1 = P.coords[4 * p + 1], t0 = P.coords[4 * p + 2], t1 = P.coords[4 * p + 3];
    const uint8_t* read = P.q_codes + P.q_off[P.pair_q[p]] + q0;
    const uint8_t* ref = P.db_codes + P.db_off[P.pair_t[p] - P.id_base] + t0;
    const int readLen = q1 - q0 + 1, refLen = t1 - t0 + 1;
    const int score = P.pair_score[p];
    const int w = wk.w, width = 2 * w + 3, width_d = 2 * w + 1;
    const int go = P.go, ge = P.ge;
    uint8_t* dir = B.dir + wk.dir_off;
    for (int j = 0; j < width + 1; ++j) { hb[j] = 0; eb[j] = 0; hc[j] = 0; }
    int best = 0;
    for (int i = 0; i < readLen; ++i) {
        const int beg = i - w > 0 ? i - w : 0;
        const int end = i + w < refLen - 1 ? i + w : refLen - 1;
        const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
        int f = 0, u = 0;
        hb[0] = 0; eb[0] = 0; hb[edge] = 0; eb[edge] = 0; hc[0] = 0;
        uint8_t* line = dir + (size_t)width_d * i;
        const int8_t* mrow = P.mat8 + read[i];          // mat8[target*32 + query]
        const int xi = beg, xp = (i - 1 - w > 0) ? i - 1 - w : 0;
        for (int j = beg; j <= end; ++j) {
            u = j - xi + 1;
            const int up = j - xp + 1, dg = up - 1;
            int open = i == 0 ? -go : hb[up] - go;
            int ext = i == 0 ? -ge : eb[up] - ge;
            const int e = open > ext ? open : ext;
            const unsigned de_open = open > ext ? 1u : 0u;
            eb[u] = e;
            open = hc[u - 1] - go;
            ext = f - ge;
            const unsigned df_open = open > ext ? 1u : 0u;
            f = open > ext ? open : ext;
            const int e1 = e > 0 ? e : 0, f1 = f > 0 ? f : 0;
            const int gap = e1 > f1 ? e1 : f1;
            const int dsc = hb[dg] + (int)__ldg(mrow + (int)ref[j] * 32);
            const int h = gap > dsc ? gap : dsc;
            hc[u] = h;
            if (h > best) best = h;
            const unsigned sel = gap <= dsc ? 0u : (e1 > f1 ? 1u : 2u);
            line[j - xi] = (uint8_t)(de_open | (df_open << 1) | (sel << 2));
        }
        for (int j = 1; j <= u; ++j) hb[j] = hc[j];
    }
    if (best < score) { B.status[tid] = 0; return; }
    // traceback from the bottom-right corner in state H until row 0 is reached (ssw.c:634-706)
    uint8_t* out = B.rev_paths + B.slot_off[p];
    const int cap = (int)(B.slot_off[p + 1] - B.slot_off[p]);
    int i = readLen - 1, j = refLen - 1, state = 2, n = 0, rc = 1;
    while (i > 0) {
        const int x = i - w > 0 ? i - w : 0;
        const int hi = i + w < refLen - 1 ? i + w : refLen - 1;
        if (j < x || j > hi || n >= cap - 1) { rc = -2; break; }
        const unsigned d = dir[(size_t)width_d * i + (j - x)];
        unsigned code;
        if (state == 0) code = (d & 1u) ? 3u : 2u;
        else if (state == 1) code = (d & 2u) ? 5u : 4u;
        else { const unsigned sel = d >> 2; code = sel == 0 ? 1u : sel == 1 ? ((d & 1u) ? 3u : 2u) : ((d & 2u) ? 5u : 4u); }
        switch (code) {
            case 1: --i; --j; state = 2; out[n++] = 1; break;
            case 2: --i; state = 0; out[n++] = 3; break;
            case 3: --i; state = 2; out[n++] = 3; break;
            case 4: --j; state = 1; out[n++] = 2; break;
            default: --j; state = 2; out[n++] = 2; break;
        }
    }
    if (rc == 1) { out[n++] = 1; B.path_len[p] = n; }
    B.status[tid] = rc;
}


// Warp-parallel banded_sw: lanes own consecutive band columns of one row, rows are sequential.
// Along a row the only serial dependency is F:  f[t] = max(hc[t-1] - go, f[t-1] - ge),  hc = max(g, f)  with
// g = max(e+, Hdiag + S) >= 0, which for ge <= go collapses to f[t] = max(g[t-1] - go, f[t-1] - ge): a max-plus
// prefix scan (5 shuffles).  Directions and values are the same integers banded_sw computes cell by cell.
constexpr int kBandWarps = 4;
constexpr int kNegBig = -(1 << 29);

// One banded_sw attempt of a warp with half band width w: fills the direction bytes, returns the banded maximum.
// bufA / bufB / E: three band rows of `2 * w + 4` ints (shared memory).
__device__ int band_fill(const int8_t* smatT, const uint8_t* read, int readLen, const uint8_t* ref, int refLen, int w, int go, int ge,
                         int32_t* bufA, int32_t* bufB, int32_t* E, uint8_t* dir, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int width = 2 * w + 3, width_d = 2 * w + 1;
    int32_t* prevH = bufA;
    int32_t* curH = bufB;
    for (int j = lane; j < width + 1; j += 32) { bufA[j] = 0; bufB[j] = 0; E[j] = 0; }
    __syncwarp();
    int best = 0;
    const int f0 = max(-go, -ge);
    for (int i = 0; i < readLen; ++i) {
        const int beg = i - w > 0 ? i - w : 0;
        const int end = i + w < refLen - 1 ? i + w : refLen - 1;
        const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
        const int xp = (i - 1 - w > 0) ? i - 1 - w : 0;
        if (lane == 0) { prevH[0] = 0; E[0] = 0; prevH[edge] = 0; E[edge] = 0; curH[0] = 0; }
        __syncwarp();
        const int8_t* mrow = smatT + (int)read[i] * 32;
        uint8_t* line = dir + (size_t)width_d * i;
        int carry_pm = kNegBig, carry_hc = 0, carry_f = 0;
        for (int c0 = beg; c0 <= end; c0 += 32) {
            const int j = c0 + lane;
            const bool act = j <= end;
            const int t = j - beg, u = t + 1, up = j - xp + 1;
            int ph = 0, pe = 0, pd = 0, sc = 0;
            if (act) { ph = prevH[up]; pe = E[up]; pd = prevH[up - 1]; sc = mrow[ref[j]]; }
            const int open_e = i == 0 ? -go : ph - go;
            const int ext_e = i == 0 ? -ge : pe - ge;
            const int e = open_e > ext_e ? open_e : ext_e;
            const unsigned de = open_e > ext_e ? 1u : 0u;
            const int e1 = e > 0 ? e : 0;
            const int dsc = pd + sc;
            const int g = e1 > dsc ? e1 : dsc;
            int a = act ? g - go + t * ge : kNegBig;
            int incl = a;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl = max(incl, y); }
            int excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = kNegBig;
            excl = max(excl, carry_pm);
            const int f = t == 0 ? f0 : max(f0 - t * ge, excl - (t - 1) * ge);
            const int hc = g > f ? g : f;
            int hc_left = __shfl_up_sync(FULL, hc, 1), f_left = __shfl_up_sync(FULL, f, 1);
            if (lane == 0) { hc_left = carry_hc; f_left = carry_f; }
            const unsigned df = (hc_left - go) > (f_left - ge) ? 1u : 0u;
            const int f1 = f > 0 ? f : 0;
            const int gap = e1 > f1 ? e1 : f1;
            const unsigned sel = gap <= dsc ? 0u : (e1 > f1 ? 1u : 2u);
            carry_pm = max(carry_pm, __shfl_sync(FULL, incl, 31));
            carry_hc = __shfl_sync(FULL, hc, 31);
            carry_f = __shfl_sync(FULL, f, 31);
            __syncwarp();
            if (act) {
                E[u] = e;
                curH[u] = hc;
                line[t] = (uint8_t)(de | (df << 1) | (sel << 2));
                best = max(best, hc);
            }
        }
        __syncwarp();
        int32_t* tmp = prevH; prevH = curH; curH = tmp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
    return best;
}

// Traceback over the direction bytes of a successful attempt (ssw.c:634-706): from the bottom-right corner in state H
// until row 0; diagonal runs are taken 32 cells at a time (every lane probes one cell of the diagonal).
// Writes the reversed path; returns its length or -2 when the walk leaves the band / the slot.
__device__ int band_trace(const uint8_t* dir, int readLen, int refLen, int w, uint8_t* out, int cap, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int width_d = 2 * w + 1;
    int i = readLen - 1, j = refLen - 1, state = 2, n = 0;
    while (i > 0) {
        const int ii = i - lane, jj = j - lane;
        unsigned d = 0xffu;
        if (ii > 0 && jj >= 0) {
            const int x = ii - w > 0 ? ii - w : 0;
            const int hi = ii + w < refLen - 1 ? ii + w : refLen - 1;
            if (jj >= x && jj <= hi) d = __ldcg(dir + (size_t)width_d * ii + (jj - x));
        }
        const unsigned d0 = __shfl_sync(FULL, d, 0);
        if (d0 == 0xffu) return -2;
        if (state == 2 && (d0 >> 2) == 0u) {
            const unsigned nd = __ballot_sync(FULL, !(d != 0xffu && (d >> 2) == 0u));
            int run = nd ? __ffs(nd) - 1 : 32;
            if (run > i) run = i;
            if (n + run >= cap) return -2;
            if (lane < run) out[n + lane] = 1;
            n += run; i -= run; j -= run;
            continue;
        }
        unsigned code;
        if (state == 0) code = (d0 & 1u) ? 3u : 2u;
        else if (state == 1) code = (d0 & 2u) ? 5u : 4u;
        else { const unsigned sel = d0 >> 2; code = sel == 1 ? ((d0 & 1u) ? 3u : 2u) : ((d0 & 2u) ? 5u : 4u); }
        if (n + 1 >= cap) return -2;
        uint8_t op;
        switch (code) {
            case 2: --i; state = 0; op = 3; break;
            case 3: --i; state = 2; op = 3; break;
            case 4: --j; state = 1; op = 2; break;
            default: --j; state = 2; op = 2; break;
        }
        if (lane == 0) out[n] = op;
        ++n;
    }
    if (lane == 0) out[n] = 1;            // the remaining cell is closed as one more match
    return n + 1;
}

// Warp-parallel banded_sw: lanes own consecutive band columns of one row, rows are sequential.
// Along a row the only serial dependency is F:  f[t] = max(hc[t-1] - go, f[t-1] - ge),  hc = max(g, f)  with
// g = max(e+, Hdiag + S) >= 0, which for ge <= go collapses to f[t] = max(g[t-1] - go, f[t-1] - ge): a max-plus
// prefix scan (5 shuffles).  Directions and values are the same integers banded_sw computes cell by cell.
// Host-sequenced variant: one attempt per work item (used for the bands the persistent kernel below hands back).
__global__ void __launch_bounds__(kBandWarps * 32) al_band_warp_kernel(AlParams P, BandParams B, unsigned long long* cursor) {
    extern __shared__ int32_t rows[];
    __shared__ int8_t smatT[32 * 32];            // [query letter][target letter]
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
        const int ql = i >> 5, tl = i & 31;
        smatT[i] = tl <= S4G_PAD_CODE ? P.mat8[tl * 32 + ql] : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    int32_t* bufA = rows + (size_t)warp * 3 * B.stride;
    int32_t* bufB = bufA + B.stride;
    int32_t* E = bufB + B.stride;
    while (true) {
        unsigned long long wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1ull);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= (unsigned long long)B.n_work) break;
        const BandWork wk = B.work[wi];
        const uint32_t p = wk.pair;
        const int q0 = P.coords[4 * p + 0], q1 = P.coords[4 * p + 1], t0 = P.coords[4 * p + 2], t1 = P.coords[4 * p + 3];
        const uint8_t* read = P.q_codes + P.q_off[P.pair_q[p]] + q0;
        const uint8_t* ref = P.db_codes + P.db_off[P.pair_t[p] - P.id_base] + t0;
        const int readLen = q1 - q0 + 1, refLen = t1 - t0 + 1;
        uint8_t* dir = B.dir + wk.dir_off;
        const int best = band_fill(smatT, read, readLen, ref, refLen, wk.w, P.go, P.ge, bufA, bufB, E, dir, lane);
        if (best < P.pair_score[p]) { if (lane == 0) B.status[wi] = 0; continue; }
        __threadfence_block();
        __syncwarp();
        const int n = band_trace(dir, readLen, refLen, wk.w, B.rev_paths + B.slot_off[p], (int)(B.slot_off[p + 1] - B.slot_off[p]), lane);
        if (lane == 0) {
            if (n > 0) B.path_len[p] = n;
            B.status[wi] = n > 0 ? 1 : -2;
        }
        __syncwarp();
    }
}

// Persistent variant: one warp takes a hit through ALL its band-doubling attempts (ssw.c:570-631) and the traceback,
// with its direction bytes in a per-warp scratch region that is reused from hit to hit.  Hits whose band outgrows the
// shared-memory rows (w > w_max) or the region are handed back in `overflow` (pair index, band width reached) for the
// host-sequenced kernel above.
struct PersistParams {
    int32_t w_max;               // largest half band width the shared-memory rows hold
    int64_t dir_per_warp;        // bytes of direction scratch per warp
    uint8_t* dir;
    uint8_t* rev_paths;
    const int64_t* slot_off;
    int32_t* path_len;
    unsigned long long* cursor;  // [0] work cursor, [1] overflow count, [2] error flags
    uint2* overflow;             // (pair, w)
    const uint32_t* order;       // processing order (largest hits first), or nullptr
    const uint2* in_list;        // (pair, half band width to start with) handed over by the group kernels, or nullptr = every hit from its first band
    const unsigned long long* in_count;
};

__global__ void __launch_bounds__(kBandWarps * 32) al_band_persistent_kernel(AlParams P, PersistParams B) {
    extern __shared__ int32_t rows[];
    __shared__ int8_t smatT[32 * 32];
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
        const int ql = i >> 5, tl = i & 31;
        smatT[i] = tl <= S4G_PAD_CODE ? P.mat8[tl * 32 + ql] : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const int stride = 2 * B.w_max + 5;
    int32_t* bufA = rows + (size_t)warp * 3 * stride;
    int32_t* bufB = bufA + stride;
    int32_t* E = bufB + stride;
    uint8_t* dir = B.dir + (size_t)(blockIdx.x * kBandWarps + warp) * B.dir_per_warp;
    while (true) {
        unsigned long long wi = 0;
        if (lane == 0) wi = atomicAdd(B.cursor, 1ull);
        wi = __shfl_sync(FULL, wi, 0);
        if ((long long)wi >= (B.in_list ? (long long)*B.in_count : P.n_pairs)) break;
        const uint32_t p = B.in_list ? B.in_list[wi].x : (B.order ? B.order[wi] : (uint32_t)wi);
        const int q0 = P.coords[4 * p + 0], q1 = P.coords[4 * p + 1], t0 = P.coords[4 * p + 2], t1 = P.coords[4 * p + 3];
        if (q0 < 0 || t0 < 0) continue;                      // reported by the sweeps
        const uint8_t* read = P.q_codes + P.q_off[P.pair_q[p]] + q0;
        const uint8_t* ref = P.db_codes + P.db_off[P.pair_t[p] - P.id_base] + t0;
        const int readLen = q1 - q0 + 1, refLen = t1 - t0 + 1;
        const int score = P.pair_score[p];
        int w = B.in_list ? (int)B.in_list[wi].y : abs(refLen - readLen) + 1;
        for (int round = 0; ; ++round) {
            if (round > 40) { if (lane == 0) atomicOr(B.cursor + 2, 2ull); break; }
            if (w > B.w_max || (int64_t)(2 * w + 1) * readLen > B.dir_per_warp) {
                if (lane == 0) B.overflow[atomicAdd(B.cursor + 1, 1ull)] = make_uint2(p, (uint32_t)w);
                break;
            }
            const int best = band_fill(smatT, read, readLen, ref, refLen, w, P.go, P.ge, bufA, bufB, E, dir, lane);
            if (best >= score) {
                __threadfence_block();
                __syncwarp();
                const int n = band_trace(dir, readLen, refLen, w, B.rev_paths + B.slot_off[p], (int)(B.slot_off[p + 1] - B.slot_off[p]), lane);
                if (lane == 0) { if (n > 0) B.path_len[p] = n; else atomicOr(B.cursor + 2, 4ull); }
                break;
            }
            w *= 2;
            __syncwarp();
        }
        __syncwarp();
    }
}

// ---- narrow bands: several hits per warp -------------------------------------------------------------------------
// The bands of near-diagonal alignments are a few cells wide (configs[1]: 2 w + 1 ~ 10 on average), and a band row costs the
// warp kernel above ~90 instructions however narrow it is -- it is bound by issue slots (ncu: 89 % issue active), with most
// lanes idle.  Here G = 8 or 16 lanes own a hit (half band width w <= (G - 1) / 2, so a row is at most one cell per lane) and
// 32 / G hits share the warp's instruction stream: same integers, same direction bytes, same traceback rules as band_fill /
// band_trace (ssw.c:549-727), shuffles confined to the group.  A hit whose band doubles beyond the group's width is handed
// on as (pair, w) -- to the next wider group kernel, finally to al_band_persistent_kernel, which continues at that width.
struct GroupParams {
    const uint2* in_list;              // (pair, w) to start from, or nullptr = every hit from its first band
    const unsigned long long* in_count;
    unsigned long long* cursor;
    uint2* out_list;
    unsigned long long* out_count;
    unsigned long long* err;           // bit 2: the traceback left the band or its slot
    uint8_t* dir;
    int64_t dir_per_group;
    uint8_t* rev_paths;
    const int64_t* slot_off;
    int32_t* path_len;
};

template <int G>
__global__ void __launch_bounds__(kBandWarps * 32) al_band_group_kernel(AlParams P, GroupParams B) {
    constexpr int kGroups = 32 / G, kWMax = (G - 1) / 2, kStride = 2 * kWMax + 5;
    __shared__ int8_t smatT[32 * 32];
    __shared__ int32_t rows[kBandWarps * kGroups * 3 * kStride];
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
        const int ql = i >> 5, tl = i & 31;
        smatT[i] = tl <= S4G_PAD_CODE ? P.mat8[tl * 32 + ql] : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1), grp = lane / G, gbase = lane & ~(G - 1);
    const unsigned FULL = 0xffffffffu;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
    int32_t* bufA = rows + (size_t)(warp * kGroups + grp) * 3 * kStride;
    int32_t* bufB = bufA + kStride;
    int32_t* E = bufB + kStride;
    uint8_t* dir = B.dir + (size_t)((blockIdx.x * kBandWarps + warp) * kGroups + grp) * B.dir_per_group;
    const long long n_work = B.in_list ? (long long)*B.in_count : P.n_pairs;
    const int go = P.go, ge = P.ge;
    const int f0 = max(-go, -ge);
    // the group's hit (the same values in all its lanes)
    bool has = false, exhausted = false;
    uint32_t p = 0;
    int w = 0, readLen = 0, refLen = 0, score = 0;
    const uint8_t *read = nullptr, *ref = nullptr;
    while (true) {
        // ---- (1) idle groups take the next hit they can hold
        while (!has && !exhausted) {
            unsigned long long wi = 0;
            if (gl == 0) wi = atomicAdd(B.cursor, 1ull);
            wi = __shfl_sync(gmask, wi, gbase);
            if ((long long)wi >= n_work) { exhausted = true; break; }
            p = B.in_list ? B.in_list[wi].x : (uint32_t)wi;
            const int q0 = P.coords[4 * p + 0], q1 = P.coords[4 * p + 1], t0 = P.coords[4 * p + 2], t1 = P.coords[4 * p + 3];
            if (q0 < 0 || t0 < 0) continue;                      // swAlign-rule hits, hits reported by the sweeps
            readLen = q1 - q0 + 1; refLen = t1 - t0 + 1;
            w = B.in_list ? (int)B.in_list[wi].y : abs(refLen - readLen) + 1;
            if (w > kWMax || (int64_t)(2 * w + 1) * readLen > B.dir_per_group) {
                if (gl == 0) B.out_list[atomicAdd(B.out_count, 1ull)] = make_uint2(p, (uint32_t)w);
                continue;
            }
            read = P.q_codes + P.q_off[P.pair_q[p]] + q0;
            ref = P.db_codes + P.db_off[P.pair_t[p] - P.id_base] + t0;
            score = P.pair_score[p];
            has = true;
        }
        __syncwarp();
        if (!__any_sync(FULL, has)) break;
        // ---- (2) one banded_sw attempt of every group that holds a hit, rows in lockstep
        const int rows_max = __reduce_max_sync(FULL, has ? readLen : 0);
        const int width = 2 * w + 3, width_d = 2 * w + 1;
        int32_t* prevH = bufA;
        int32_t* curH = bufB;
        for (int j = gl; j < kStride; j += G) { bufA[j] = 0; bufB[j] = 0; E[j] = 0; }
        __syncwarp();
        int best = 0;
        for (int i = 0; i < rows_max; ++i) {
            const bool rowact = has && i < readLen;
            const int beg = i - w > 0 ? i - w : 0;
            const int end = i + w < refLen - 1 ? i + w : refLen - 1;
            const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            const int xp = (i - 1 - w > 0) ? i - 1 - w : 0;
            if (gl == 0 && rowact) { prevH[0] = 0; E[0] = 0; prevH[edge] = 0; E[edge] = 0; curH[0] = 0; }
            __syncwarp();
            const int j = beg + gl;
            const bool act = rowact && j <= end;
            const int t = j - beg, u = t + 1, up = j - xp + 1;
            int ph = 0, pe = 0, pd = 0, sc = 0;
            if (act) { ph = prevH[up]; pe = E[up]; pd = prevH[up - 1]; sc = smatT[(int)read[i] * 32 + ref[j]]; }
            const int open_e = i == 0 ? -go : ph - go;
            const int ext_e = i == 0 ? -ge : pe - ge;
            const int e = open_e > ext_e ? open_e : ext_e;
            const unsigned de = open_e > ext_e ? 1u : 0u;
            const int e1 = e > 0 ? e : 0;
            const int dsc = pd + sc;
            const int g = e1 > dsc ? e1 : dsc;
            const int a = act ? g - go + t * ge : kNegBig;
            int incl = a;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o, G); if (gl >= o) incl = max(incl, y); }
            int excl = __shfl_up_sync(FULL, incl, 1, G);
            if (gl == 0) excl = kNegBig;
            const int f = t == 0 ? f0 : max(f0 - t * ge, excl - (t - 1) * ge);
            const int hc = g > f ? g : f;
            int hc_left = __shfl_up_sync(FULL, hc, 1, G), f_left = __shfl_up_sync(FULL, f, 1, G);
            if (gl == 0) { hc_left = 0; f_left = 0; }
            const unsigned df = (hc_left - go) > (f_left - ge) ? 1u : 0u;
            const int f1 = f > 0 ? f : 0;
            const int gap = e1 > f1 ? e1 : f1;
            const unsigned sel = gap <= dsc ? 0u : (e1 > f1 ? 1u : 2u);
            __syncwarp();
            if (act) {
                E[u] = e;
                curH[u] = hc;
                dir[(size_t)width_d * i + t] = (uint8_t)(de | (df << 1) | (sel << 2));
                best = max(best, hc);
            }
            __syncwarp();
            int32_t* tmp = prevH; prevH = curH; curH = tmp;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
        // ---- (3) groups whose band reached the score trace back; the others double their band (or hand the hit on)
        bool tracing = has && best >= score;
        if (has && !tracing) {
            w *= 2;
            if (w > kWMax || (int64_t)(2 * w + 1) * readLen > B.dir_per_group) {
                if (gl == 0) B.out_list[atomicAdd(B.out_count, 1ull)] = make_uint2(p, (uint32_t)w);
                has = false;
            }
        }
        __threadfence_block();
        __syncwarp();
        if (__any_sync(FULL, tracing)) {
            uint8_t* out = B.rev_paths + (tracing ? B.slot_off[p] : 0);
            const int cap = tracing ? (int)(B.slot_off[p + 1] - B.slot_off[p]) : 0;
            int i = readLen - 1, j = refLen - 1, state = 2, n = 0;
            bool fail = false;
            while (true) {
                const bool tr = tracing && !fail && i > 0;
                if (!__any_sync(FULL, tr)) break;
                const int ii = i - gl, jj = j - gl;
                unsigned d = 0xffu;
                if (tr && ii > 0 && jj >= 0) {
                    const int x = ii - w > 0 ? ii - w : 0;
                    const int hi = ii + w < refLen - 1 ? ii + w : refLen - 1;
                    if (jj >= x && jj <= hi) d = __ldcg(dir + (size_t)width_d * ii + (jj - x));
                }
                const unsigned d0 = __shfl_sync(FULL, d, gbase);
                const unsigned ndw = __ballot_sync(FULL, !(d != 0xffu && (d >> 2) == 0u));
                if (!tr) continue;
                if (d0 == 0xffu) { fail = true; continue; }
                if (state == 2 && (d0 >> 2) == 0u) {
                    const unsigned nd = (ndw & gmask) >> gbase;
                    int run = nd ? __ffs(nd) - 1 : G;
                    if (run > i) run = i;
                    if (n + run >= cap) { fail = true; continue; }
                    if (gl < run) out[n + gl] = 1;
                    n += run; i -= run; j -= run;
                    continue;
                }
                unsigned code;
                if (state == 0) code = (d0 & 1u) ? 3u : 2u;
                else if (state == 1) code = (d0 & 2u) ? 5u : 4u;
                else { const unsigned sl = d0 >> 2; code = sl == 1 ? ((d0 & 1u) ? 3u : 2u) : ((d0 & 2u) ? 5u : 4u); }
                if (n + 1 >= cap) { fail = true; continue; }
                uint8_t op;
                switch (code) {
                    case 2: --i; state = 0; op = 3; break;
                    case 3: --i; state = 2; op = 3; break;
                    case 4: --j; state = 1; op = 2; break;
                    default: --j; state = 2; op = 2; break;
                }
                if (gl == 0) out[n] = op;
                ++n;
            }
            if (tracing) {
                if (gl == 0) {
                    if (fail) atomicOr(B.err, 4ull);
                    else { out[n] = 1; B.path_len[p] = n + 1; }          // the remaining cell is closed as one more match
                }
                has = false;
            }
        }
        __syncwarp();
    }
}

// ---- swAlign-rule paths (score > 32767, or gap penalties SSW cannot take) ----------------------------------------
// Restates swAlign (vendor/swsharp/swsharp/src/cpu_module.c:1185-1413) on the rectangle [0..er] x [0..ec] that ends in
// its end cell, in linear working memory plus 4 bits per cell: move (STOP/DIAG/LEFT/UP with the reference's priority
// UP > LEFT > DIAG > STOP) and two flags saying whether the horizontal / vertical gap arriving in the cell was an
// extension (open/extend ties count as extension, :1291-1305) -- the gap lengths the reference stores per cell are the
// run lengths of these flags.  One warp per hit, same systolic sweep as above (lanes own 8 rows, 256 rows per pass).
// Direction words: [pass][lane][column], nibble r of a word = row 8 * lane + r of the pass.
constexpr int kSMIN = -1000000000;

struct SwaWork { uint32_t pair; uint32_t pad; int64_t dir_off; };   // dir_off in 32-bit words

__global__ void __launch_bounds__(kSwWarps * 32) al_swalign_kernel(AlParams P, const SwaWork* work, int n_work, uint32_t* dir_pool,
                                                                    uint8_t* rev_paths, const int64_t* slot_off, int32_t* path_len,
                                                                    unsigned long long* cursor) {
    extern __shared__ __align__(16) unsigned char ssm[];
    int8_t* smat = reinterpret_cast<int8_t*>(ssm);
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wb = ssm + kSwSmatBytes + warp * kSwWarpBytes;
    unsigned* prof = reinterpret_cast<unsigned*>(wb);
    unsigned short* ring = reinterpret_cast<unsigned short*>(prof + kSwProfWords);
    int* inH = reinterpret_cast<int*>(ring + 2 * kSwRing);
    int* inF = inH + kSwRing;
    int* outH = inF + kSwRing;
    int* outF = outH + 2 * kSwRing;
    const int gwarp = blockIdx.x * kSwWarps + warp;
    int32_t* bH = P.bound + (int64_t)gwarp * P.bound_stride;
    int32_t* bF = bH + P.bound_stride / 2;
    const unsigned FULL = 0xffffffffu;
    const int Q = P.go, R = P.ge;
    const char* prof_b = reinterpret_cast<const char*>(prof) + lane * 4;
    while (true) {
        unsigned long long wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1ull);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= (unsigned long long)n_work) break;
        const uint32_t p = work[wi].pair;
        uint32_t* dir = dir_pool + work[wi].dir_off;
        const int er = P.coords[4 * p + 1], ec = P.coords[4 * p + 3];
        const uint8_t* q = P.q_codes + P.q_off[P.pair_q[p]];
        const uint8_t* t = P.db_codes + P.db_off[P.pair_t[p] - P.id_base];
        const int rows = er + 1, cols = ec + 1;
        const int npass = (rows + 32 * kSwRows - 1) / (32 * kSwRows);
        // ---- fill
        for (int pass = 0; pass < npass; ++pass) {
            const int row0 = pass * 32 * kSwRows + lane * kSwRows;
            const bool first = pass == 0, last = pass == npass - 1;
            {
                int ql[kSwRows];
#pragma unroll
                for (int r = 0; r < kSwRows; ++r) ql[r] = row0 + r < rows ? (int)q[row0 + r] : -1;
                for (int letter = 0; letter <= S4G_PAD_CODE; ++letter) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        unsigned word = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int v = ql[4 * m + b] >= 0 ? (int)smat[letter * 32 + ql[4 * m + b]] : -128;
                            word |= (unsigned)(v & 0xff) << (8 * b);
                        }
                        prof[(letter * 2 + m) * 32 + lane] = word;
                    }
                }
            }
            for (int c = lane; c < kSwRing; c += 32) ring[kSwRing + c] = (unsigned short)(S4G_PAD_CODE * 256);
            int Hl[kSwRows], A[kSwRows];                 // score and horizontal-gap value of the cell to the left
#pragma unroll
            for (int r = 0; r < kSwRows; ++r) { Hl[r] = 0; A[r] = kSMIN; }
            int h_last = 0, d_last = kSMIN, diag_in = 0;
            uint32_t* dline = dir + ((size_t)pass * 32 + lane) * cols;
            const int nsteps = cols + 31;
            int flushed = 0;
            for (int s0 = 0; s0 < nsteps; s0 += kSwRing) {
                if (!last) {
                    const int upto = min(cols, s0 - 31);
                    for (int c = flushed + lane; c < upto; c += 32) { bH[c] = outH[c & (2 * kSwRing - 1)]; bF[c] = outF[c & (2 * kSwRing - 1)]; }
                    if (upto > flushed) flushed = upto;
                }
                {
                    const int base = s0 & (2 * kSwRing - 1);
#pragma unroll
                    for (int c = 0; c < kSwRing; c += 32) {
                        const int j = s0 + c + lane;
                        ring[base + c + lane] = (unsigned short)((j < cols ? (unsigned)t[j] : (unsigned)S4G_PAD_CODE) * 256u);
                        if (!first) { inH[c + lane] = j < cols ? bH[j] : 0; inF[c + lane] = j < cols ? bF[j] : kSMIN; }
                    }
                }
                __syncwarp();
                const int send = (nsteps - s0) < kSwRing ? (nsteps - s0) : kSwRing;
#pragma unroll 1
                for (int ss = 0; ss < send; ++ss) {
                    const int j = s0 + ss - lane;
                    const unsigned o = ring[j & (2 * kSwRing - 1)];
                    const unsigned w0 = *reinterpret_cast<const unsigned*>(prof_b + o);
                    const unsigned w1 = *reinterpret_cast<const unsigned*>(prof_b + o + 128);
                    int hs = __shfl_up_sync(FULL, h_last, 1);        // score of the cell above the lane's first row
                    int ha = __shfl_up_sync(FULL, d_last, 1);        // its vertical-gap value
                    if (lane == 0) { hs = first ? 0 : inH[ss]; ha = first ? kSMIN : inF[ss]; }
                    const bool live = j >= 0 && j < cols;
                    if (!live) { hs = 0; ha = kSMIN; }
                    int diag = diag_in;
                    diag_in = hs;
                    unsigned word = 0;
#pragma unroll
                    for (int r = 0; r < kSwRows; ++r) {
                        const unsigned w = r < 4 ? w0 : w1;
                        const int mch = diag + (int)(int8_t)((w >> (8 * (r & 3))) & 0xffu);
                        const int ins_o = Hl[r] - Q, ins_e = A[r] - R;
                        const int ins = max(ins_o, ins_e);
                        const int del_o = hs - Q, del_e = ha - R;
                        const int del = max(del_o, del_e);
                        const int scr = max(max(0, mch), max(ins, del));
                        const unsigned mv = del == scr ? 3u : ins == scr ? 2u : mch == scr ? 1u : 0u;
                        word |= (mv | (ins == ins_e ? 4u : 0u) | (del == del_e ? 8u : 0u)) << (4 * r);
                        diag = Hl[r];
                        if (live) { Hl[r] = scr; A[r] = ins; }
                        hs = scr; ha = del;
                    }
                    h_last = hs; d_last = ha;
                    if (live) dline[j] = word;
                    if (lane == 31 && !last && live) { outH[j & (2 * kSwRing - 1)] = h_last; outF[j & (2 * kSwRing - 1)] = d_last; }
                }
                __syncwarp();
            }
            if (!last) for (int c = flushed + lane; c < cols; c += 32) { bH[c] = outH[c & (2 * kSwRing - 1)]; bF[c] = outF[c & (2 * kSwRing - 1)]; }
            __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
        // ---- traceback (cpu_module.c:1349-1400): lane 0 walks, reading one nibble per step
        if (lane == 0) {
            uint8_t* out = rev_paths + slot_off[p];
            const int cap = (int)(slot_off[p + 1] - slot_off[p]);
            int r = er, c = ec, n = 0;
            bool ok = true;
            auto nib = [&](int rr, int cc) -> unsigned {
                const int ps = rr / (32 * kSwRows), ln = (rr % (32 * kSwRows)) / kSwRows, k = rr % kSwRows;
                return (__ldcg(dir + ((size_t)ps * 32 + ln) * cols + cc) >> (4 * k)) & 15u;
            };
            while (r >= 0 && c >= 0) {
                const unsigned d = nib(r, c), mv = d & 3u;
                if (mv == 1u) { if (n >= cap) { ok = false; break; } out[n++] = 1; --r; --c; }
                else if (mv == 2u) {
                    // LEFT: the gap consumes hGaps + 1 columns; hGaps = run of extension flags ending here
                    unsigned f = d;
                    while (true) {
                        if (n >= cap) { ok = false; break; }
                        out[n++] = 2; --c;
                        if (!(f & 4u) || c < 0) break;
                        f = nib(r, c);
                    }
                    if (!ok) break;
                } else if (mv == 3u) {
                    unsigned f = d;
                    while (true) {
                        if (n >= cap) { ok = false; break; }
                        out[n++] = 3; --r;
                        if (!(f & 8u) || r < 0) break;
                        f = nib(r, c);
                    }
                    if (!ok) break;
                } else { ++r; ++c; break; }
            }
            if (r == -1 || c == -1) { ++r; ++c; }
            if (ok) { P.coords[4 * p + 0] = r; P.coords[4 * p + 2] = c; path_len[p] = n; }
            else atomicOr(P.counters + 1, 4ull);
        }
        __syncwarp();
    }
}

__global__ void al_slot_sizes_kernel(AlParams P, int64_t* sizes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P.n_pairs) return;
    int64_t s = 0;
    if (i < P.n_pairs) {
        if (takes_swalign(P, P.pair_score[i])) { if (P.coords[4 * i + 1] >= 0) s = (int64_t)P.coords[4 * i + 1] + P.coords[4 * i + 3] + 4; }
        else if (P.coords[4 * i] >= 0) s = (P.coords[4 * i + 1] - P.coords[4 * i] + 1) + (P.coords[4 * i + 3] - P.coords[4 * i + 2] + 1) + 2;
    }
    sizes[i] = s;
}

__global__ void al_len64_kernel(const int32_t* path_len, int64_t n, int64_t* len64) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    len64[i] = (i < n && path_len[i] > 0) ? path_len[i] : 0;
}

// DP cells each traceback phase has to cover (statistic for the stage's roofline; see s4g_last_align_profile)
__global__ void al_cells_kernel(AlParams P, unsigned long long* acc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long e = 0, b = 0, p = 0;
    if (i < P.n_pairs) {
        const int q0 = P.coords[4 * i], q1 = P.coords[4 * i + 1], t0 = P.coords[4 * i + 2], t1 = P.coords[4 * i + 3];
        const uint32_t qi = P.pair_q[i];
        const unsigned long long qlen = (unsigned long long)(P.q_off[qi + 1] - P.q_off[qi]);
        if (q1 >= 0 && t1 >= 0) e = qlen * (unsigned long long)(t1 + 1);
        if (q0 >= 0 && t0 >= 0 && q1 >= q0 && t1 >= t0) {
            b = (unsigned long long)(q1 + 1) * (unsigned long long)(t1 - t0 + 1);
            const int dq = q1 - q0 + 1, dt = t1 - t0 + 1;
            p = (unsigned long long)dq * (unsigned long long)(2 * (abs(dt - dq) + 1) + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e += __shfl_down_sync(0xffffffffu, e, o); b += __shfl_down_sync(0xffffffffu, b, o); p += __shfl_down_sync(0xffffffffu, p, o); }
    if ((threadIdx.x & 31) == 0) { if (e) atomicAdd(acc, e); if (b) atomicAdd(acc + 1, b); if (p) atomicAdd(acc + 2, p); }
}

// forward-order copy of each reversed path into the packed output
__global__ void al_pack_paths_kernel(const uint8_t* rev, const int64_t* slot_off, const int32_t* path_len, const int64_t* out_off,
                                     int64_t n, uint8_t* out, int64_t cap) {
    const int64_t p = blockIdx.x;
    if (p >= n) return;
    const int len = path_len[p] > 0 ? path_len[p] : 0;
    const uint8_t* src = rev + slot_off[p];
    uint8_t* dst = out + out_off[p];
    if (out_off[p] + len > cap) return;
    for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[len - 1 - i];
}

}  // namespace

static int sw_align_impl(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_pairs, const uint32_t* pair_q,
                         const uint32_t* pair_t, const int32_t* pair_score, const int32_t* matrix, int gap_open,
                         int gap_extend, int32_t* out_coords, uint8_t* out_paths, int64_t path_capacity,
                         int64_t* out_path_offsets, int where);

namespace {
// NVLink-striped database: the targets of the hits are copied once into a resident buffer (coalesced peer reads at link
// bandwidth), so that the three traceback phases -- ring refills, band rows, direction walks: latency-sensitive gathers that
// touch every target several times -- run on local memory.  Hit h's target becomes sequence h of a temporary database.
__global__ void al_gather_lens_kernel(const uint32_t* pair_t, int64_t n, const int64_t* db_off, uint32_t id_base, int64_t* lens) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t l = 0;
    if (i < n) { const uint32_t t = pair_t[i] - id_base; l = db_off[t + 1] - db_off[t]; }
    lens[i] = l;
}
__global__ void al_gather_codes_kernel(const uint32_t* pair_t, int64_t n, const uint8_t* db_codes, const int64_t* db_off, uint32_t id_base,
                                       const int64_t* out_off, uint8_t* out, uint32_t* iota) {
    const int64_t h = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= n) return;
    const int lane = threadIdx.x & 31;
    const uint32_t t = pair_t[h] - id_base;
    const uint8_t* src = db_codes + db_off[t];
    uint8_t* dst = out + out_off[h];
    const int64_t len = out_off[h + 1] - out_off[h];
    // 4-byte body on the source's alignment (the destination is written byte-wise: it is local)
    int64_t head = (int64_t)((4 - (reinterpret_cast<uintptr_t>(src) & 3u)) & 3u);
    if (head > len) head = len;
    if (lane < head) dst[lane] = src[lane];
    const int64_t nw = (len - head) / 4;
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + head);
    for (int64_t w = lane; w < nw; w += 32) {
        const uint32_t v = s4[w];
        uint8_t* d = dst + head + 4 * w;
        d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16); d[3] = (uint8_t)(v >> 24);
    }
    const int64_t done = head + 4 * nw;
    if (lane < len - done) dst[done + lane] = src[done + lane];
    if (lane == 0) iota[h] = (uint32_t)h;
}
}  // namespace

extern "C" int s4g_sw_align(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_pairs, const uint32_t* pair_q,
                            const uint32_t* pair_t, const int32_t* pair_score, const int32_t* matrix, int gap_open,
                            int gap_extend, int32_t* out_coords, uint8_t* out_paths, int64_t path_capacity,
                            int64_t* out_path_offsets, int where) {
    const char* e = getenv("S4G_ALIGN_GATHER");
    const bool gather = db && db->borrowed_codes && n_pairs > 0 && n_pairs < ((int64_t)1 << 31) && pair_t && !(e && e[0] == '0');
    if (!ctx || !gather) return sw_align_impl(ctx, db, q, n_pairs, pair_q, pair_t, pair_score, matrix, gap_open, gap_extend, out_coords, out_paths, path_capacity, out_path_offsets, where);
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t* d_pt = pair_t;
    char* g = (char*)s4g_scratch(ctx, SLOT_AL_GATHER_IDX, 2 * sizeof(int64_t) * (size_t)(n_pairs + 1) + 2 * sizeof(uint32_t) * (size_t)n_pairs + 64);
    if (!g) return S4G_ERR_NOMEM;
    int64_t* d_lens = (int64_t*)g;
    int64_t* d_goff = d_lens + (n_pairs + 1);
    uint32_t* d_iota = (uint32_t*)(d_goff + (n_pairs + 1));
    uint32_t* d_pt_up = d_iota + n_pairs;
    if (where == S4G_HOST) {
        for (int64_t i = 0; i < n_pairs; ++i)
            if (pair_t[i] < db->id_base || pair_t[i] - db->id_base >= (uint64_t)db->n) { s4g_set_error(ctx, "pair %lld references an unknown target", (long long)i); return S4G_ERR_ARG; }
        S4G_CUDA(ctx, cudaMemcpyAsync(d_pt_up, pair_t, 4 * n_pairs, cudaMemcpyHostToDevice, st));
        d_pt = d_pt_up;
    }
    al_gather_lens_kernel<<<(unsigned)((n_pairs + 1 + 255) / 256), 256, 0, st>>>(d_pt, n_pairs, db->d_off, db->id_base, d_lens);
    S4G_CHECK_LAUNCH(ctx);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_lens, d_goff, (int)(n_pairs + 1), st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_lens, d_goff, (int)(n_pairs + 1), st));
        ctx->launches += 1;
    }
    s4g_db tdb;
    tdb.ctx = ctx; tdb.n = n_pairs; tdb.id_base = 0; tdb.max_len = db->max_len; tdb.d_off = d_goff;
    tdb.h_off.resize(n_pairs + 1);
    S4G_CUDA(ctx, cudaMemcpyAsync(tdb.h_off.data(), d_goff, sizeof(int64_t) * (n_pairs + 1), cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    if (where == S4G_DEVICE && tdb.h_off[n_pairs] < 0) return S4G_ERR_ARG;
    tdb.residues = (uint64_t)tdb.h_off[n_pairs];
    uint8_t* d_gcodes = (uint8_t*)s4g_scratch(ctx, SLOT_AL_GATHER, tdb.residues + S4G_DB_TAIL_PAD);
    if (!d_gcodes) return S4G_ERR_NOMEM;
    tdb.d_codes = d_gcodes;
    al_gather_codes_kernel<<<(unsigned)((n_pairs + 7) / 8), 256, 0, st>>>(d_pt, n_pairs, db->d_codes, db->d_off, db->id_base, d_goff, d_gcodes, d_iota);
    S4G_CHECK_LAUNCH(ctx);
    S4G_CUDA(ctx, cudaMemsetAsync(d_gcodes + tdb.residues, S4G_PAD_CODE, S4G_DB_TAIL_PAD, st));
    std::vector<uint32_t> h_iota;
    const uint32_t* t_arg = d_iota;
    if (where == S4G_HOST) { h_iota.resize(n_pairs); for (int64_t i = 0; i < n_pairs; ++i) h_iota[i] = (uint32_t)i; t_arg = h_iota.data(); }
    const int rc = sw_align_impl(ctx, &tdb, q, n_pairs, pair_q, t_arg, pair_score, matrix, gap_open, gap_extend, out_coords, out_paths, path_capacity, out_path_offsets, where);
    tdb.d_codes = nullptr; tdb.d_off = nullptr;           // scratch of the context, not the temporary's
    return rc;
}

static int sw_align_impl(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_pairs, const uint32_t* pair_q,
                         const uint32_t* pair_t, const int32_t* pair_score, const int32_t* matrix, int gap_open,
                         int gap_extend, int32_t* out_coords, uint8_t* out_paths, int64_t path_capacity,
                         int64_t* out_path_offsets, int where) {
    if (!ctx || !db || !q || n_pairs < 0 || !matrix || !out_path_offsets) return S4G_ERR_ARG;
    if (n_pairs > 0 && (!pair_q || !pair_t || !pair_score || !out_coords || !out_paths)) return S4G_ERR_ARG;
    if (n_pairs >= ((int64_t)1 << 31)) { s4g_set_error(ctx, "s4g_sw_align: %lld hits in one call (limit 2^31 - 1); split the batch", (long long)n_pairs); return S4G_ERR_CAPACITY; }
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (n_pairs == 0) {
        int64_t zero = 0;
        if (where == S4G_HOST) out_path_offsets[0] = 0;
        else S4G_CUDA(ctx, cudaMemcpyAsync(out_path_offsets, &zero, 8, cudaMemcpyHostToDevice, st));
        return S4G_OK;
    }
    const bool swalign_all = abs(gap_open) > 127 || abs(gap_extend) > 127;     // sse_module.c:215-236: SSW takes 8-bit penalties only

    // host copies of the pair arrays (the band rounds are sequenced on the host)
    std::vector<uint32_t> h_q(n_pairs), h_t(n_pairs);
    std::vector<int32_t> h_s(n_pairs);
    const uint32_t *d_pq, *d_pt; const int32_t* d_ps;
    if (where == S4G_HOST) {
        memcpy(h_q.data(), pair_q, 4 * n_pairs); memcpy(h_t.data(), pair_t, 4 * n_pairs); memcpy(h_s.data(), pair_score, 4 * n_pairs);
        uint32_t* a = (uint32_t*)s4g_scratch(ctx, SLOT_IO_A, 4 * n_pairs);
        uint32_t* b = (uint32_t*)s4g_scratch(ctx, SLOT_IO_B, 4 * n_pairs);
        int32_t* c = (int32_t*)s4g_scratch(ctx, SLOT_IO_C, 4 * n_pairs);
        if (!a || !b || !c) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cudaMemcpyAsync(a, pair_q, 4 * n_pairs, cudaMemcpyHostToDevice, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(b, pair_t, 4 * n_pairs, cudaMemcpyHostToDevice, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(c, pair_score, 4 * n_pairs, cudaMemcpyHostToDevice, st));
        d_pq = a; d_pt = b; d_ps = c;
    } else {
        S4G_CUDA(ctx, cudaMemcpyAsync(h_q.data(), pair_q, 4 * n_pairs, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(h_t.data(), pair_t, 4 * n_pairs, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(h_s.data(), pair_score, 4 * n_pairs, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaStreamSynchronize(st));
        d_pq = pair_q; d_pt = pair_t; d_ps = pair_score;
    }
    for (int64_t i = 0; i < n_pairs; ++i) {
        if (h_q[i] >= (uint32_t)q->n || h_t[i] < db->id_base || h_t[i] - db->id_base >= (uint64_t)db->n) { s4g_set_error(ctx, "pair %lld references an unknown query or target", (long long)i); return S4G_ERR_ARG; }
    }

    int8_t h_mat8[(S4G_PAD_CODE + 1) * 32];
    memset(h_mat8, 0, sizeof(h_mat8));
    int min_s = 0;
    for (int a = 0; a < S4G_NLET; ++a)
        for (int b = 0; b < S4G_NLET; ++b) {
            int v = matrix[b * S4G_NLET + a];
            if (v > 127 || v < -127) { s4g_set_error(ctx, "matrix entry %d outside int8 range", v); return S4G_ERR_ARG; }
            h_mat8[a * 32 + b] = (int8_t)v;
            if (v < min_s) min_s = v;
        }
    for (int b = 0; b < 32; ++b) h_mat8[S4G_PAD_CODE * 32 + b] = (int8_t)(min_s < -1 ? min_s : -1);

    const size_t sw_smem = kSwSmatBytes + (size_t)kSwWarps * kSwWarpBytes;
    S4G_CUDA(ctx, cudaFuncSetAttribute(al_sweep32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem));
    int sw_per_sm = 0;
    S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sw_per_sm, al_sweep32_kernel, kSwWarps * 32, sw_smem));
    if (sw_per_sm < 1) sw_per_sm = 1;
    const int sw_blocks = ctx->sm_count * sw_per_sm;
    const int64_t bound_stride = 2 * ((int64_t)std::max(db->max_len, q->max_len) + 64);
    char* misc = (char*)s4g_scratch(ctx, SLOT_AL_MISC, 256 + sizeof(h_mat8) + sizeof(int64_t) * 4 * (n_pairs + 1) + sizeof(int32_t) * 6 * n_pairs);
    int32_t* d_bound = (int32_t*)s4g_scratch(ctx, SLOT_SW_BOUND, sizeof(int32_t) * bound_stride * sw_blocks * kSwWarps);
    if (!misc || !d_bound) return S4G_ERR_NOMEM;
    unsigned long long* d_counters = (unsigned long long*)misc;
    int8_t* d_mat8 = (int8_t*)(misc + 64);
    int64_t* d_sizes = (int64_t*)(misc + 64 + ((sizeof(h_mat8) + 63) / 64) * 64);
    int64_t* d_slot_off = d_sizes + (n_pairs + 1);
    int64_t* d_len64 = d_slot_off + (n_pairs + 1);
    int64_t* d_out_off = d_len64 + (n_pairs + 1);
    int32_t* d_coords_own = (int32_t*)(d_out_off + (n_pairs + 1));
    int32_t* d_path_len = d_coords_own + 4 * n_pairs;
    int32_t* d_coords = where == S4G_DEVICE ? out_coords : d_coords_own;

    S4G_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 64, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_mat8, h_mat8, sizeof(h_mat8), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemsetAsync(d_path_len, 0xff, sizeof(int32_t) * n_pairs, st));
    S4G_CUDA(ctx, cudaMemsetAsync(d_coords, 0xff, sizeof(int32_t) * 4 * n_pairs, st));

    AlParams P;
    P.db_codes = db->d_codes; P.db_off = db->d_off; P.id_base = db->id_base;
    P.q_codes = q->d_codes; P.q_off = q->d_off;
    P.pair_q = d_pq; P.pair_t = d_pt; P.pair_score = d_ps; P.n_pairs = n_pairs;
    P.mat8 = d_mat8; P.go = gap_open; P.ge = gap_extend; P.coords = d_coords; P.counters = d_counters;
    P.bound = d_bound; P.bound_stride = bound_stride; P.swalign_all = swalign_all ? 1 : 0;
    int64_t n_swa = 0;
    for (int64_t i = 0; i < n_pairs; ++i) n_swa += (swalign_all || h_s[i] > 32767) ? 1 : 0;

    s4g_trace_start(ctx);
    S4G_CUDA(ctx, cudaEventRecord(ctx->ev_al[0], st));
    // 1: end cells -- packed s16x2 sweep, two hits of a query per warp; 32-bit sweep for queries beyond its reach
    {
        if (!swalign_all) {
            int rc = s4g_sw_forward_ends_device(ctx, db, q, n_pairs, d_pq, d_pt, d_ps, d_mat8, gap_open, gap_extend, d_coords, d_counters + 1);
            if (rc != S4G_OK) return rc;
        }
        if (swalign_all || n_swa > 0 || q->max_len > s4g_sw_long_query_rows()) {
            al_sweep32_kernel<<<sw_blocks, kSwWarps * 32, sw_smem, st>>>(P, 1, swalign_all ? 0 : s4g_sw_long_query_rows(), d_counters + 3);
            S4G_CHECK_LAUNCH(ctx);
        }
    }
    S4G_CUDA(ctx, cudaEventRecord(ctx->ev_al[1], st));
    s4g_trace_mark(ctx, "ends");
    // 2: begin cells -- reverse sweep from the end cell, stopped at the first column that holds the score: packed, two hits per
    // warp (align_rev.cu); the 32-bit sweep takes what is left (end rows beyond 1024; S4G_BEGINS=sweep32 forces it for all)
    {
        const char* e = getenv("S4G_BEGINS");
        const bool packed = !(e && strcmp(e, "sweep32") == 0);
        if (packed) {
            int rc = s4g_sw_reverse_begins_device(ctx, db, q, n_pairs, d_pq, d_pt, d_ps, d_mat8, gap_open, gap_extend, swalign_all ? 1 : 0, d_coords, d_counters + 1);
            if (rc != S4G_OK) return rc;
        }
        if (!packed || q->max_len > 1024) {
            al_sweep32_kernel<<<sw_blocks, kSwWarps * 32, sw_smem, st>>>(P, 0, 0, d_counters + 0);
            S4G_CHECK_LAUNCH(ctx);
        }
    }
    S4G_CUDA(ctx, cudaEventRecord(ctx->ev_al[2], st));
    s4g_trace_mark(ctx, "begins");
    // slots for the reversed paths (device scan; the host only needs an upper bound to size the buffer)
    al_slot_sizes_kernel<<<(unsigned)((n_pairs + 1 + 255) / 256), 256, 0, st>>>(P, d_sizes);
    S4G_CHECK_LAUNCH(ctx);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_sizes, d_slot_off, (int)(n_pairs + 1), st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_sizes, d_slot_off, (int)(n_pairs + 1), st));
        ctx->launches += 1;
    }
    int64_t total_slots = 0;
    for (int64_t i = 0; i < n_pairs; ++i)
        total_slots += (q->h_off[h_q[i] + 1] - q->h_off[h_q[i]]) + (db->h_off[h_t[i] - db->id_base + 1] - db->h_off[h_t[i] - db->id_base]) + 2;
    uint8_t* d_rev = (uint8_t*)s4g_scratch(ctx, SLOT_AL_OUT, (size_t)total_slots + 64);
    if (!d_rev) return S4G_ERR_NOMEM;

    // 3: banded traceback.  Persistent kernel: every warp takes a hit through all its band-doubling attempts; what
    // outgrows its shared-memory rows comes back in an overflow list for the host-sequenced rounds below.
    struct Pending { uint32_t pair; int32_t w, readLen, refLen; };
    std::vector<Pending> pending;
    std::vector<int32_t> h_coords;
    bool warp_rule = gap_extend <= gap_open;           // the max-plus scan of the warp kernels needs ge <= go
    { const char* e = getenv("S4G_BAND"); if (e && strcmp(e, "thread") == 0) warp_rule = false; }     // experiment: one thread per hit for every band
    unsigned long long h_flags[4] = {0, 0, 0, 0};
    if (warp_rule) {
        const int w_max = 256;
        const size_t p_smem = (size_t)kBandWarps * 3 * (2 * w_max + 5) * sizeof(int32_t);
        int per_sm = 0;
        S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, al_band_persistent_kernel, kBandWarps * 32, p_smem));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        const int grid = (int)std::min<int64_t>((n_pairs + kBandWarps - 1) / kBandWarps, (int64_t)ctx->sm_count * per_sm);
        const int64_t dir_per_warp = 640 * 1024;
        uint8_t* d_dirp = (uint8_t*)s4g_scratch(ctx, SLOT_AL_DIR, (size_t)grid * kBandWarps * dir_per_warp + 64);
        char* d_pw = (char*)s4g_scratch(ctx, SLOT_AL_WORK, 64 + sizeof(uint2) * n_pairs);
        if (!d_dirp || !d_pw) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cudaMemsetAsync(d_pw, 0, 64, st));
        PersistParams B;
        B.w_max = w_max; B.dir_per_warp = dir_per_warp; B.dir = d_dirp; B.rev_paths = d_rev; B.slot_off = d_slot_off; B.path_len = d_path_len;
        B.cursor = (unsigned long long*)d_pw; B.overflow = (uint2*)(d_pw + 64); B.order = nullptr; B.in_list = nullptr; B.in_count = nullptr;
        // narrow bands first, several hits per warp: 8 lanes per hit (w <= 3), then 16 (w <= 7); what outgrows them is handed on as
        // (pair, w) and the warp-per-hit kernel continues at that width.  S4G_BAND_GROUPS=0 / 8 / 16 selects the stages (default both).
        {
            const char* e = getenv("S4G_BAND_GROUPS");
            const int stages = e ? atoi(e) : 24;                 // bit 3: the 8-lane stage, bit 4: the 16-lane stage
            char* d_gl = (char*)s4g_scratch(ctx, SLOT_AL_GROUPS, 128 + 2 * sizeof(uint2) * (size_t)n_pairs);
            if (!d_gl) return S4G_ERR_NOMEM;
            S4G_CUDA(ctx, cudaMemsetAsync(d_gl, 0, 128, st));
            unsigned long long* g_cnt = (unsigned long long*)d_gl;      // [0] cursor8 [1] count A [2] cursor16 [3] count B
            uint2* listA = (uint2*)(d_gl + 128);
            uint2* listB = listA + n_pairs;
            GroupParams G;
            G.err = (unsigned long long*)d_pw + 2; G.dir = d_dirp; G.rev_paths = d_rev; G.slot_off = d_slot_off; G.path_len = d_path_len;
            const uint2* cur_list = nullptr;
            const unsigned long long* cur_count = nullptr;
            if (stages & 8) {
                G.in_list = nullptr; G.in_count = nullptr; G.cursor = g_cnt + 0; G.out_list = listA; G.out_count = g_cnt + 1; G.dir_per_group = dir_per_warp / 4;
                al_band_group_kernel<8><<<grid, kBandWarps * 32, 0, st>>>(P, G);
                S4G_CHECK_LAUNCH(ctx);
                cur_list = listA; cur_count = g_cnt + 1;
            }
            if (stages & 16) {
                G.in_list = cur_list; G.in_count = cur_count; G.cursor = g_cnt + 2; G.out_list = listB; G.out_count = g_cnt + 3; G.dir_per_group = dir_per_warp / 2;
                al_band_group_kernel<16><<<grid, kBandWarps * 32, 0, st>>>(P, G);
                S4G_CHECK_LAUNCH(ctx);
                cur_list = listB; cur_count = g_cnt + 3;
            }
            B.in_list = cur_list; B.in_count = cur_count;
        }
        al_band_persistent_kernel<<<grid, kBandWarps * 32, p_smem, st>>>(P, B);
        S4G_CHECK_LAUNCH(ctx);
        S4G_CUDA(ctx, cudaMemcpyAsync(h_flags, d_pw, 32, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_counters[2];
    S4G_CUDA(ctx, cudaMemcpyAsync(h_counters, d_counters, 16, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_counters[1] & 1ull) { s4g_set_error(ctx, "s4g_sw_align: a pair's score is not attained by any cell (score does not belong to the pair)"); return S4G_ERR_ARG; }
    if (h_flags[2] & 2ull) { s4g_set_error(ctx, "s4g_sw_align: band doubling did not converge"); return S4G_ERR_INTERNAL; }
    if (h_flags[2] & 4ull) { s4g_set_error(ctx, "s4g_sw_align: traceback left the band (the reference's behaviour is undefined there)"); return S4G_ERR_INTERNAL; }
    s4g_trace_mark(ctx, "band_persistent");
    if (ctx->trace) {
        unsigned long long g[4] = {0, 0, 0, 0};
        if (ctx->slot_ptr[SLOT_AL_GROUPS] && warp_rule) cudaMemcpy(g, ctx->slot_ptr[SLOT_AL_GROUPS], 32, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[s4g trace] align: %lld hits, %llu handed from the 8-lane to the 16-lane band kernel, %llu on to the warp kernel, %llu handed back to the host-sequenced band rounds\n",
                (long long)n_pairs, g[1], g[3], h_flags[1]);
    }
    const bool need_host_rounds = !warp_rule || h_flags[1] > 0;
    if (need_host_rounds || where == S4G_HOST) {
        h_coords.resize(4 * n_pairs);
        S4G_CUDA(ctx, cudaMemcpyAsync(h_coords.data(), d_coords, sizeof(int32_t) * 4 * n_pairs, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (!warp_rule) {
        pending.clear();
        for (int64_t i = 0; i < n_pairs; ++i) {
            if (swalign_all || h_s[i] > 32767) continue;
            const int readLen = h_coords[4 * i + 1] - h_coords[4 * i] + 1, refLen = h_coords[4 * i + 3] - h_coords[4 * i + 2] + 1;
            pending.push_back({(uint32_t)i, abs(refLen - readLen) + 1, readLen, refLen});
        }
    } else if (h_flags[1] > 0) {
        std::vector<uint2> ovf(h_flags[1]);
        S4G_CUDA(ctx, cudaMemcpy(ovf.data(), (char*)ctx->slot_ptr[SLOT_AL_WORK] + 64, sizeof(uint2) * ovf.size(), cudaMemcpyDeviceToHost));
        pending.resize(ovf.size());
        for (size_t k = 0; k < ovf.size(); ++k) {
            const int64_t i = ovf[k].x;
            pending[k] = {(uint32_t)i, (int32_t)ovf[k].y, h_coords[4 * i + 1] - h_coords[4 * i] + 1, h_coords[4 * i + 3] - h_coords[4 * i + 2] + 1};
        }
    }
    const size_t dir_budget = (size_t)4 << 30;
    const int threads = 64;
    const size_t smem_limit = 200 * 1024;
    int round = 0;
    while (!pending.empty()) {
        if (++round > 40) { s4g_set_error(ctx, "s4g_sw_align: band doubling did not converge"); return S4G_ERR_INTERNAL; }
        std::sort(pending.begin(), pending.end(), [](const Pending& a, const Pending& b) { return a.w < b.w; });
        std::vector<Pending> next;
        size_t pos = 0;
        while (pos < pending.size()) {
            // batch: hits of similar band width whose direction bytes fit the budget
            size_t end = pos, dir_bytes = 0;
            const int w_lo = pending[pos].w;
            std::vector<BandWork> work;
            int w_max = w_lo;
            while (end < pending.size() && pending[end].w <= 2 * w_lo + 2) {
                const size_t need = (size_t)(2 * pending[end].w + 1) * pending[end].readLen;
                if (!work.empty() && dir_bytes + need > dir_budget) break;
                work.push_back({pending[end].pair, pending[end].w, (int64_t)dir_bytes});
                dir_bytes += need;
                w_max = std::max(w_max, pending[end].w);
                ++end;
            }
            int stride = (2 * w_max + 5) | 1;
            const bool warp_path = gap_extend <= gap_open && (size_t)kBandWarps * 3 * stride * 4 <= smem_limit;
            int tpb = threads;
            while (tpb > 1 && (size_t)tpb * 3 * stride * 4 > smem_limit) tpb >>= 1;
            if (!warp_path && (size_t)tpb * 3 * stride * 4 > smem_limit) { s4g_set_error(ctx, "s4g_sw_align: band %d too wide for this build", w_max); return S4G_ERR_CAPACITY; }
            const size_t smem = warp_path ? (size_t)kBandWarps * 3 * stride * 4 : (size_t)tpb * 3 * stride * 4;
            BandWork* d_work = (BandWork*)s4g_scratch(ctx, SLOT_AL_WORK, sizeof(BandWork) * work.size() + sizeof(int32_t) * work.size() + 64);
            uint8_t* d_dir = (uint8_t*)s4g_scratch(ctx, SLOT_AL_DIR, dir_bytes + 64);
            if (!d_work || !d_dir) return S4G_ERR_NOMEM;
            int32_t* d_status = (int32_t*)(d_work + work.size());
            S4G_CUDA(ctx, cudaMemcpyAsync(d_work, work.data(), sizeof(BandWork) * work.size(), cudaMemcpyHostToDevice, st));
            BandParams B;
            B.work = d_work; B.n_work = (int32_t)work.size(); B.stride = stride; B.dir = d_dir; B.rev_paths = d_rev;
            B.slot_off = d_slot_off; B.path_len = d_path_len; B.status = d_status;
            if (warp_path) {
                S4G_CUDA(ctx, cudaFuncSetAttribute(al_band_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
                S4G_CUDA(ctx, cudaMemsetAsync(d_counters + 2, 0, 8, st));
                int per_sm = 0;
                S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, al_band_warp_kernel, kBandWarps * 32, smem));
                if (per_sm < 1) per_sm = 1;
                unsigned grid = (unsigned)std::min<size_t>((work.size() + kBandWarps - 1) / kBandWarps, (size_t)ctx->sm_count * per_sm);
                al_band_warp_kernel<<<grid, kBandWarps * 32, smem, st>>>(P, B, d_counters + 2);
            } else {
                S4G_CUDA(ctx, cudaFuncSetAttribute(al_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
                al_band_kernel<<<(unsigned)((work.size() + tpb - 1) / tpb), tpb, smem, st>>>(P, B);
            }
            S4G_CHECK_LAUNCH(ctx);
            std::vector<int32_t> h_status(work.size());
            S4G_CUDA(ctx, cudaMemcpyAsync(h_status.data(), d_status, sizeof(int32_t) * work.size(), cudaMemcpyDeviceToHost, st));
            S4G_CUDA(ctx, cudaStreamSynchronize(st));
            for (size_t i = 0; i < work.size(); ++i) {
                if (h_status[i] == 0) { Pending p2 = pending[pos + i]; p2.w *= 2; next.push_back(p2); }
                else if (h_status[i] < 0) { s4g_set_error(ctx, "s4g_sw_align: traceback left the band for pair %u (the reference's behaviour is undefined there)", work[i].pair); return S4G_ERR_INTERNAL; }
            }
            pos = end;
        }
        pending.swap(next);
    }

    // 3b: swAlign-rule hits -- full direction rectangles, batched by memory
    if (n_swa > 0) {
        if (h_coords.empty()) {
            h_coords.resize(4 * n_pairs);
            S4G_CUDA(ctx, cudaMemcpyAsync(h_coords.data(), d_coords, sizeof(int32_t) * 4 * n_pairs, cudaMemcpyDeviceToHost, st));
            S4G_CUDA(ctx, cudaStreamSynchronize(st));
        }
        S4G_CUDA(ctx, cudaFuncSetAttribute(al_swalign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem));
        const int64_t budget_words = (int64_t)2 << 30;          // 8 GiB of direction words per batch
        std::vector<SwaWork> work;
        int64_t used = 0;
        auto flush = [&]() -> int {
            if (work.empty()) return S4G_OK;
            SwaWork* d_work = (SwaWork*)s4g_scratch(ctx, SLOT_AL_WORK, sizeof(SwaWork) * work.size() + 64);
            uint32_t* d_dirw = (uint32_t*)s4g_scratch(ctx, SLOT_AL_DIR, sizeof(uint32_t) * (size_t)used + 64);
            if (!d_work || !d_dirw) return S4G_ERR_NOMEM;
            S4G_CUDA(ctx, cudaMemcpyAsync(d_work, work.data(), sizeof(SwaWork) * work.size(), cudaMemcpyHostToDevice, st));
            S4G_CUDA(ctx, cudaMemsetAsync(d_counters + 4, 0, 8, st));
            const int grid = (int)std::min<int64_t>((int64_t)(work.size() + kSwWarps - 1) / kSwWarps, sw_blocks);
            al_swalign_kernel<<<grid, kSwWarps * 32, sw_smem, st>>>(P, d_work, (int)work.size(), d_dirw, d_rev, d_slot_off, d_path_len, d_counters + 4);
            S4G_CHECK_LAUNCH(ctx);
            S4G_CUDA(ctx, cudaStreamSynchronize(st));           // the host vector is reused by the next batch
            work.clear();
            used = 0;
            return S4G_OK;
        };
        for (int64_t i = 0; i < n_pairs; ++i) {
            if (!(swalign_all || h_s[i] > 32767)) continue;
            const int64_t rows = (int64_t)h_coords[4 * i + 1] + 1, cols = (int64_t)h_coords[4 * i + 3] + 1;
            if (rows <= 0 || cols <= 0) continue;
            const int64_t words = ((rows + 32 * kSwRows - 1) / (32 * kSwRows)) * 32 * cols;
            if (!work.empty() && used + words > budget_words) { int rc = flush(); if (rc != S4G_OK) return rc; }
            work.push_back({(uint32_t)i, 0u, used});
            used += words;
        }
        int rc = flush();
        if (rc != S4G_OK) return rc;
        unsigned long long h_err = 0;
        S4G_CUDA(ctx, cudaMemcpy(&h_err, d_counters + 1, 8, cudaMemcpyDeviceToHost));
        if (h_err & 4ull) { s4g_set_error(ctx, "s4g_sw_align: a swAlign-rule path does not fit its slot (internal)"); return S4G_ERR_INTERNAL; }
        // coords of these hits changed on the device
        if (where == S4G_HOST) {
            S4G_CUDA(ctx, cudaMemcpyAsync(h_coords.data(), d_coords, sizeof(int32_t) * 4 * n_pairs, cudaMemcpyDeviceToHost, st));
            S4G_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    S4G_CUDA(ctx, cudaEventRecord(ctx->ev_al[3], st));
    ctx->al_timed = true;
    al_cells_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(P, d_counters + 5);
    S4G_CHECK_LAUNCH(ctx);
    s4g_trace_mark(ctx, "band_rounds");
    // pack: path offsets = exclusive scan of the lengths, then forward-order copy
    al_len64_kernel<<<(unsigned)((n_pairs + 1 + 255) / 256), 256, 0, st>>>(d_path_len, n_pairs, d_len64);
    S4G_CHECK_LAUNCH(ctx);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_len64, d_out_off, (int)(n_pairs + 1), st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_len64, d_out_off, (int)(n_pairs + 1), st));
        ctx->launches += 1;
    }
    int64_t h_total = 0;
    S4G_CUDA(ctx, cudaMemcpyAsync(&h_total, d_out_off + n_pairs, 8, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(ctx->al_cells, d_counters + 5, 24, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_total > path_capacity) { s4g_set_error(ctx, "s4g_sw_align: paths need %lld bytes, capacity %lld", (long long)h_total, (long long)path_capacity); return S4G_ERR_CAPACITY; }
    uint8_t* d_paths = where == S4G_DEVICE ? out_paths : (uint8_t*)s4g_scratch(ctx, SLOT_IO_D, (size_t)h_total + 64);
    if (!d_paths) return S4G_ERR_NOMEM;
    al_pack_paths_kernel<<<(unsigned)n_pairs, 64, 0, st>>>(d_rev, d_slot_off, d_path_len, d_out_off, n_pairs, d_paths, h_total);
    S4G_CHECK_LAUNCH(ctx);
    if (where == S4G_HOST) {
        memcpy(out_coords, h_coords.data(), sizeof(int32_t) * 4 * n_pairs);
        S4G_CUDA(ctx, cudaMemcpyAsync(out_paths, d_paths, (size_t)h_total, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(out_path_offsets, d_out_off, sizeof(int64_t) * (n_pairs + 1), cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaStreamSynchronize(st));
    } else {
        S4G_CUDA(ctx, cudaMemcpyAsync(out_path_offsets, d_out_off, sizeof(int64_t) * (n_pairs + 1), cudaMemcpyDeviceToDevice, st));
    }
    s4g_trace_mark(ctx, "pack");
    s4g_trace_report(ctx, "align");
    return S4G_OK;
}
