// Stage 2: Smith-Waterman affine-gap SCORES of every query against its candidates (sm_100a).
//
// Replaces the swimd scorer behind scoreDatabaseCpu (vendor/swsharp/swsharp/src/cpu_module.c:179-192,
// swimd/Swimd.cpp:139-449) and the legacy CUDA scorers behind scoreDatabasesGpu
// (src/gpu_module.h:280-291).  Result = exact integer max over all cells of
//   E[i][j] = max(E[i][j-1] - R, H[i][j-1] - Q)
//   F[i][j] = max(F[i-1][j] - R, H[i-1][j] - Q)
//   H[i][j] = max(0, H[i-1][j-1] + S[q_i][t_j], E[i][j], F[i][j])        (Q = gap open, R = gap extend)
//
// Kernel design (inter-sequence, one warp per PAIR of candidates of the same query):
//   * the two targets of a pair ride in the two 16-bit halves of every register (s16x2), and all
//     recurrences are Blackwell DPX instructions: VIADDMNMX.S16x2(.RELU), VIMNMX.S16x2, VIMNMX3.S16x2,
//     plus VIADD.16x2 for H-Q (which issues on a different pipe than the DPX/ALU ops -- measured,
//     tools/dpx_microbench.cu);
//   * query rows are striped over the 32 lanes, K consecutive rows per lane held in registers
//     (H and E, 2*K registers); lane L works on target column (step - L), so the anti-diagonal
//     dependency is carried by two warp shuffles per step (H and F of the lane's last row);
//   * the query profile (int8, [27 letters][K/4 words][32 lanes]) sits in shared memory in a layout that
//     makes every lane's load bank-conflict free whatever letter each lane is looking at; one PRMT with
//     sign replication builds the packed s16x2 substitution word for both targets;
//   * target residues are staged per warp through a small shared-memory ring as pre-multiplied profile
//     row offsets (2 x LDS.U16 per step, zero ALU work);
//   * pairs whose 16-bit score could have wrapped (best > 32767 - max(S)) and queries longer than
//     32*32 rows are re-run by the 32-bit multi-pass kernel (exact, any length).
//
// Work decomposition: candidates of each query are sorted by length (device radix sort) and paired
// neighbour-wise; a persistent grid of CTAs pulls (query, block of pairs) tiles from an atomic counter,
// builds the query profile once per tile and lets its warps pull pairs from the tile.  The tile table runs
// over the queries in DESCENDING LENGTH (ScoreParams::q_order): the two CTAs of an SM then execute the same
// row-class instantiation (one 6-9 KB step loop in the 32 KB instruction cache instead of two) and the
// expensive tiles go first.
//
// The shipped score kernel (sw_score_packed2_kernel / stream_pairs_packed2) takes TWO stream columns per
// step: a lane works on columns 2(step - lane) and 2(step - lane) + 1; the first column's H is a temporary
// (the diagonal input of the next row's second cell), which removes the one register move per cell ptxas
// puts into the one-column loop, and halves shuffles, ring reads and loop control per column.  The
// one-column forms (stream_pairs_packed, score_pair_packed) stay for the end-cell mode and as A/B arms.
#include <cstring>
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;                 // warps per CTA
constexpr int kTilePairs = 64;            // pairs per (query, tile)
constexpr int kRing = 64;                 // ring refill granularity (columns), >= 32
constexpr int kMaxK = 32;                 // rows per lane in the packed kernel -> queries up to 1024
constexpr int kGenK = 8;                  // rows per lane in the 32-bit kernel (256 rows per pass)
constexpr int kStripColsMin = 4096;       // boundary rows of the striped kernel: at least this many columns per target, grown to the shard's longest sequence

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

struct ScoreParams {
    const uint8_t* db_codes;
    const int64_t* db_off;
    uint32_t id_base;
    const uint8_t* q_codes;
    const int64_t* q_off;
    int32_t nq;
    const uint32_t* cand_ids;       // original order
    const int64_t* cand_off;        // nq+1
    const uint32_t* sorted_idx;     // positions into cand_ids, grouped by query, longest target first
    const int64_t* tile_start;      // nq+1: exclusive scan of packed-kernel tiles per query, queries taken in q_order
    const int32_t* q_order;         // queries by descending length (nullptr: as given): co-resident CTAs run the same row class
                                    // (one step loop in the instruction cache) and the long tiles go first
    const int8_t* mat8;             // 27 x 32 int8 (row = target letter incl. pad, col = query letter)
    int32_t* out;                   // scores, original order
    unsigned long long* counters;   // [0] tile counter, [1] overflow count, [2] generic work counter
    uint32_t* overflow;             // list of cand positions to re-run in 32 bit
    int32_t* bound;                 // generic kernel: per-warp boundary rows
    int64_t bound_stride;           // ints per warp
    int32_t gap_open, gap_extend;
    int32_t ovf_limit;              // 32767 - max(matrix)
    // end-cell mode (stage 3, step 1): cand_ids = targets of the kept hits, out = coords (4 per hit)
    const int32_t* pair_score;
    // striped kernel (queries longer than 32 * kMaxK rows)
    const int64_t* long_tile_start;   // nq+1: exclusive scan of its tiles per query
    unsigned* strip_bound;            // per CTA: kTilePairs x 2 x strip_cols packed boundary rows (H, F)
    int32_t strip_cols;               // columns per boundary row (targets beyond it take the 32-bit kernel)
};

// ------------------------------------------------------------------------------------------------------
// packed s16x2 kernel body for one pair, K rows per lane

// TRACK = false: returns the packed maxima of the two targets.
// TRACK = true : the scores are known (score2, packed); finds for each target the first column, then the first row in
//                it, whose H equals the score -- SSW's end cell (ssw.c:283-308,491-512) -- as (col << 10 | row) in
//                found1 / found2, and stops as soon as both are settled.
constexpr unsigned kNotFound = 0xffffffffu;

template <int K, bool TRACK>
__device__ __forceinline__ unsigned score_pair_packed(const unsigned* __restrict__ prof_lane,   // smem, + lane
                                                      unsigned short* ring1, unsigned short* ring2,
                                                      const uint8_t* __restrict__ t1, int len1,
                                                      const uint8_t* __restrict__ t2, int len2,
                                                      unsigned negQ, unsigned negR, int lane,
                                                      unsigned score2 = 0, int qlen = 0, unsigned* found1 = nullptr,
                                                      unsigned* found2 = nullptr) {
    constexpr int KW = (K + 3) / 4;
    constexpr unsigned kRowBytes = KW * 128;             // bytes per profile letter row
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr int kRingMask = 2 * kRing - 1;
    const unsigned FULL = 0xffffffffu;

    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0;
    unsigned fnd1 = kNotFound, fnd2 = kNotFound;

    const int maxlen = len1 > len2 ? len1 : len2;
    const int nsteps = maxlen + 31;

    // columns -kRing..-1 read as padding
    for (int c = lane; c < kRing; c += 32) { ring1[kRing + c] = kPadOff; ring2[kRing + c] = kPadOff; }

    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);

    for (int s0 = 0; s0 < nsteps; s0 += kRing) {
        if (TRACK && s0 > 0) {
            // a target is settled once every lane has passed the column of its first hit (or its end)
            unsigned m1 = fnd1, m2 = fnd2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { m1 = min(m1, __shfl_xor_sync(FULL, m1, o)); m2 = min(m2, __shfl_xor_sync(FULL, m2, o)); }
            const bool d1 = (m1 != kNotFound && (int)(m1 >> 10) + 32 <= s0) || len1 + 31 <= s0;
            const bool d2 = (m2 != kNotFound && (int)(m2 >> 10) + 32 <= s0) || len2 + 31 <= s0;
            if (d1 && d2) break;
        }
        // refill: columns s0 .. s0+kRing-1 into ring slot (s0/kRing)&1
        {
            const int base = s0 & kRingMask;
#pragma unroll
            for (int c = 0; c < kRing; c += 32) {
                int j = s0 + c + lane;
                unsigned o1 = j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff;
                unsigned o2 = j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff;
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
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                w1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1 + m * 128);
                w2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2 + m * 128);
            }
            unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
            unsigned f = __shfl_up_sync(FULL, f_out, 1);
            if (lane == 0) { h_up = 0; f = 0; }
            // Row r needs t_r = Hdiag + S = H_old[r-1] + S[r]; it is formed one row ahead (VIADD.16x2, which issues on the
            // FMA pipe) so that H[r] can be overwritten in place without register moves, and the cell update itself
            // is three DPX instructions on the ALU pipe: VIMNMX3.RELU, VIADDMNMX, VIADDMNMX.
            unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;
            diag_in = h_up;
            if (TRACK) best = 0;                                            // per-column maximum in end-cell mode
#pragma unroll
            for (int r = 0; r < K; ++r) {
                unsigned t_next = 0;
                if (r + 1 < K) {
                    const unsigned sel = ((r + 1) & 3) == 0 ? 0xC480u : ((r + 1) & 3) == 1 ? 0xD591u : ((r + 1) & 3) == 2 ? 0xE6A2u : 0xF7B3u;
                    t_next = __vadd2(H[r], prmt(w1[(r + 1) >> 2], w2[(r + 1) >> 2], sel));
                }
                const unsigned h = __vimax3_s16x2_relu(t, E[r], f);          // H = max(Hdiag + S, E, F, 0)
                H[r] = h;
                const unsigned hq = __vadd2(h, negQ);                      // H - Q
                E[r] = __viaddmax_s16x2(E[r], negR, hq);                   // E of the next column
                f = __viaddmax_s16x2(f, negR, hq);                         // F of the next row
                // the matrix maximum is always reached by a diagonal step (gaps only lower a score), so tracking
                // t = Hdiag + S instead of H finds the same maximum -- and gives the add a second consumer, which keeps
                // ptxas from folding it back into a VIADDMNMX on the ALU pipe
                if (r & 1) best = __vimax3_s16x2(best, t_prev, t);
                else if (r == K - 1) best = __vmaxs2(best, t);
                t_prev = t;
                t = t_next;
            }
            h_last = H[K - 1];
            f_out = f;
            if (TRACK) {
                const unsigned eq = __vcmpeq2(best, score2);
                if (eq) {
                    const int col = s0 + ss - lane;
                    if ((eq & 0xffffu) && col < len1) {
                        int rr = -1;
#pragma unroll
                        for (int r = K - 1; r >= 0; --r) if ((H[r] & 0xffffu) == (score2 & 0xffffu) && lane * K + r < qlen) rr = r;
                        if (rr >= 0) fnd1 = min(fnd1, ((unsigned)col << 10) | (unsigned)(lane * K + rr));
                    }
                    if ((eq >> 16) && col < len2) {
                        int rr = -1;
#pragma unroll
                        for (int r = K - 1; r >= 0; --r) if ((H[r] >> 16) == (score2 >> 16) && lane * K + r < qlen) rr = r;
                        if (rr >= 0) fnd2 = min(fnd2, ((unsigned)col << 10) | (unsigned)(lane * K + rr));
                    }
                }
            }
        }
        __syncwarp();
    }
    if (TRACK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { fnd1 = min(fnd1, __shfl_xor_sync(FULL, fnd1, o)); fnd2 = min(fnd2, __shfl_xor_sync(FULL, fnd2, o)); }
        *found1 = fnd1; *found2 = fnd2;
        return 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = __vmaxs2(best, __shfl_xor_sync(FULL, best, o));
    return best;
}

// ------------------------------------------------------------------------------------------------------
// streaming form of the packed sweep (score mode): the pairs a warp takes from its tile are laid end to end as ONE
// stream of target columns, so the 31-step fill/drain of the systolic wavefront is paid once per warp and tile instead
// of once per pair (candidate lists of random data are dominated by ~100-residue targets, where the drain is 29 % of the
// steps).  The first column of every pair carries a flag (bit 15 of its ring entry); a lane that reaches it
//   * hands its running maximum of the finished pair down the lanes (a third SHFL.UP rides the wavefront: lane L gets
//     lane L-1's merged maximum exactly when it crosses the same boundary one step later; lane 31 writes the score),
//   * clears its H / E rows, diagonal input and maximum.
// The flag block is the only divergent code and runs once per lane and pair.  After the last pair a flagged sentinel
// column and 31 pad columns flush the wavefront.
constexpr int kDescRing = 16;             // output positions of the pairs in flight
constexpr int kMinCols = 8;               // pairs are padded to this many columns (bounds the pairs in flight)
#ifndef S4G_STREAM_TILE
#define S4G_STREAM_TILE 256
#endif
#ifndef S4G_TRACK_TILE
#define S4G_TRACK_TILE 64
#endif
constexpr int kStreamTilePairs = S4G_STREAM_TILE;     // pairs per (query, tile) of the streaming score kernel
constexpr int kTrackTilePairs = S4G_TRACK_TILE;       // pairs per (query, tile) of the end-cell kernel
constexpr unsigned kNoTarget = 0xffffffffu;

template <int K>
__device__ __forceinline__ void stream_pairs_packed(const ScoreParams& P, const unsigned* __restrict__ prof_lane, unsigned short* ring1,
                                                    unsigned short* ring2, uint2* desc, int* s_next, int64_t cbeg, int64_t cend,
                                                    int pair_end, unsigned negQ, unsigned negR, int lane) {
    constexpr int KW = (K + 3) / 4;
    constexpr unsigned kRowBytes = KW * 128;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr unsigned kFlag = 0x8000u;
    constexpr int kRingMask = 2 * kRing - 1;
    constexpr int kOpen = 0x7fffffff;
    const unsigned FULL = 0xffffffffu;
    static_assert(kPadOff < kFlag, "profile offsets must leave bit 15 free");

    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0, b_out = 0;
    int n31 = 0;                          // boundaries this lane has crossed (used by lane 31)

    // fill state (warp uniform): current pair and the next column of it to be staged
    const uint8_t *t1 = nullptr, *t2 = nullptr;
    int len1 = 0, len2 = 0, L = 0, pos = 0, n_pulled = 0, end_col = kOpen;

    for (int c = lane; c < kRing; c += 32) { ring1[kRing + c] = kPadOff; ring2[kRing + c] = kPadOff; }
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);

    for (int s0 = 0;; s0 += kRing) {
        // ---- stage stream columns s0 .. s0+kRing-1
        int col = s0;
        while (col < s0 + kRing) {
            if (pos >= L && end_col == kOpen) {
                int p = 0;
                if (lane == 0) p = atomicAdd(s_next, 1);
                p = __shfl_sync(FULL, p, 0);
                if (p < pair_end) {
                    const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
                    const uint32_t c1 = P.sorted_idx[i1];
                    const bool has2 = i2 < cend;
                    const uint32_t c2 = has2 ? P.sorted_idx[i2] : c1;
                    const uint32_t g1 = P.cand_ids[c1] - P.id_base, g2 = P.cand_ids[c2] - P.id_base;
                    const int64_t a1 = P.db_off[g1], b1 = P.db_off[g1 + 1];
                    const int64_t a2 = P.db_off[g2], b2 = P.db_off[g2 + 1];
                    t1 = P.db_codes + a1; len1 = (int)(b1 - a1);
                    t2 = P.db_codes + a2; len2 = has2 ? (int)(b2 - a2) : 0;
                    L = len1 > len2 ? len1 : len2;
                    if (L < kMinCols) L = kMinCols;
                    pos = 0;
                    if (lane == 0) desc[n_pulled & (kDescRing - 1)] = make_uint2(c1, has2 ? c2 : kNoTarget);
                    ++n_pulled;
                } else {
                    end_col = col;                                   // the sentinel boundary sits here
                }
            }
            if (end_col != kOpen) {
                for (int c = col + lane; c < s0 + kRing; c += 32) {
                    ring1[c & kRingMask] = (unsigned short)(c == end_col ? (kPadOff | kFlag) : kPadOff);
                    ring2[c & kRingMask] = (unsigned short)kPadOff;
                }
                break;
            }
            const int n = min(L - pos, s0 + kRing - col);
            for (int i = lane; i < n; i += 32) {
                const int j = pos + i;
                unsigned o1 = j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff;
                const unsigned o2 = j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff;
                if (j == 0) o1 |= kFlag;
                ring1[(col + i) & kRingMask] = (unsigned short)o1;
                ring2[(col + i) & kRingMask] = (unsigned short)o2;
            }
            col += n; pos += n;
        }
        if (n_pulled == 0) return;                                   // the tile was empty for this warp
        __syncwarp();
        // lane 31 crosses the sentinel at step end_col + 31
        const int send = end_col == kOpen ? kRing : min(kRing, end_col + 32 - s0);
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int j = (s0 + ss - lane) & kRingMask;
            int o1 = (short)ring1[j];
            const unsigned o2 = ring2[j];
            unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
            unsigned f = __shfl_up_sync(FULL, f_out, 1);
            unsigned b_in = __shfl_up_sync(FULL, b_out, 1);
            if (lane == 0) { h_up = 0; f = 0; b_in = 0; }
            if (o1 < 0) {                                            // first column of a pair (or the sentinel)
                o1 &= 0x7fff;
                b_out = __vmaxs2(best, b_in);
                if (lane == 31) {
                    if (n31 > 0) {
                        const uint2 d = desc[(n31 - 1) & (kDescRing - 1)];
                        const int s1 = (int)(b_out & 0xffffu), s2 = (int)(b_out >> 16);
                        if (s1 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = d.x; else P.out[d.x] = s1;
                        if (d.y != kNoTarget) { if (s2 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = d.y; else P.out[d.y] = s2; }
                    }
                    ++n31;
                }
                best = 0; diag_in = 0;
#pragma unroll
                for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
            }
            unsigned w1[KW], w2[KW];
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                w1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1 + m * 128);
                w2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2 + m * 128);
            }
            unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;
            diag_in = h_up;
#pragma unroll
            for (int r = 0; r < K; ++r) {                            // cell update: see score_pair_packed
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
        }
        __syncwarp();
        if (end_col != kOpen && s0 + kRing >= end_col + 32) break;
    }
}

// Two columns per step (round 2, the shipped score path): a lane takes the columns 2(step - lane) and 2(step - lane) + 1 of the
// stream through its K rows in one pass.  Per row the first column's H is a temporary (it is the diagonal input of the next row's
// second cell), only the second column's H and the E entering the next step are carried over the loop: ptxas no longer parks the
// new H in a temporary and moves it back (one IMAD.MOV per cell in the one-column loop: it forms Hdiag + S in place in H's register
// at the top of the step), and shuffles, ring reads and loop control are paid once per two columns.  SASS of the step loop:
// K = 17: 8.4 issue slots per cell against 10.3, K = 32: 7.7 against 9.1; the ALU pipe (4.5 ops per cell) is the only bound left.
// Pairs are padded to an even number of columns (boundaries sit on even stream columns: the flag rides the first column).
constexpr int kDescRing2 = 32;            // pairs in flight: 64 staged columns + 62 columns back to lane 31, 8 columns per pair at least

template <int K>
__device__ __forceinline__ void stream_pairs_packed2(const ScoreParams& P, const unsigned* __restrict__ prof_lane, unsigned short* ring1,
                                                     unsigned short* ring2, uint2* desc, int* s_next, int64_t cbeg, int64_t cend,
                                                     int pair_end, unsigned negQ, unsigned negR, int lane) {
    constexpr int KW = (K + 3) / 4;
    constexpr unsigned kRowBytes = KW * 128;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr unsigned kFlag = 0x8000u;
    constexpr int kRingMask = 2 * kRing - 1;
    constexpr int kSteps = kRing / 2;     // steps per refill
    constexpr int kOpen = 0x7fffffff;
    const unsigned FULL = 0xffffffffu;
    static_assert(kPadOff < kFlag, "profile offsets must leave bit 15 free");
    static_assert(kMinCols % 2 == 0 && kRing % 2 == 0, "even columns");

    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, hA_last = 0, hB_last = 0, fA_out = 0, fB_out = 0, diag_in = 0, b_out = 0;
    int n31 = 0;                          // boundaries this lane has crossed (used by lane 31)

    // fill state (warp uniform): current pair and the next column of it to be staged
    const uint8_t *t1 = nullptr, *t2 = nullptr;
    int len1 = 0, len2 = 0, L = 0, pos = 0, n_pulled = 0, end_col = kOpen;

    for (int c = lane; c < kRing; c += 32) { ring1[kRing + c] = kPadOff; ring2[kRing + c] = kPadOff; }
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
    const unsigned* ring1w = reinterpret_cast<const unsigned*>(ring1);
    const unsigned* ring2w = reinterpret_cast<const unsigned*>(ring2);

    for (int s0 = 0;; s0 += kRing) {
        // ---- stage stream columns s0 .. s0+kRing-1
        int col = s0;
        while (col < s0 + kRing) {
            if (pos >= L && end_col == kOpen) {
                int p = 0;
                if (lane == 0) p = atomicAdd(s_next, 1);
                p = __shfl_sync(FULL, p, 0);
                if (p < pair_end) {
                    const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
                    const uint32_t c1 = P.sorted_idx[i1];
                    const bool has2 = i2 < cend;
                    const uint32_t c2 = has2 ? P.sorted_idx[i2] : c1;
                    const uint32_t g1 = P.cand_ids[c1] - P.id_base, g2 = P.cand_ids[c2] - P.id_base;
                    const int64_t a1 = P.db_off[g1], b1 = P.db_off[g1 + 1];
                    const int64_t a2 = P.db_off[g2], b2 = P.db_off[g2 + 1];
                    t1 = P.db_codes + a1; len1 = (int)(b1 - a1);
                    t2 = P.db_codes + a2; len2 = has2 ? (int)(b2 - a2) : 0;
                    L = len1 > len2 ? len1 : len2;
                    L = (L + 1) & ~1;                                // pad columns change no maximum
                    if (L < kMinCols) L = kMinCols;
                    pos = 0;
                    if (lane == 0) desc[n_pulled & (kDescRing2 - 1)] = make_uint2(c1, has2 ? c2 : kNoTarget);
                    ++n_pulled;
                } else {
                    end_col = col;                                   // the sentinel boundary sits here (an even column)
                }
            }
            if (end_col != kOpen) {
                for (int c = col + lane; c < s0 + kRing; c += 32) {
                    ring1[c & kRingMask] = (unsigned short)(c == end_col ? (kPadOff | kFlag) : kPadOff);
                    ring2[c & kRingMask] = (unsigned short)kPadOff;
                }
                break;
            }
            const int n = min(L - pos, s0 + kRing - col);
            for (int i = lane; i < n; i += 32) {
                const int j = pos + i;
                unsigned o1 = j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff;
                const unsigned o2 = j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff;
                if (j == 0) o1 |= kFlag;
                ring1[(col + i) & kRingMask] = (unsigned short)o1;
                ring2[(col + i) & kRingMask] = (unsigned short)o2;
            }
            col += n; pos += n;
        }
        if (n_pulled == 0) return;                                   // the tile was empty for this warp
        __syncwarp();
        // lane 31 crosses the sentinel at step end_col / 2 + 31
        const int S0 = s0 >> 1;
        const int send = end_col == kOpen ? kSteps : min(kSteps, (end_col >> 1) + 32 - S0);
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int j = (S0 + ss - lane) & (kRing - 1);            // column pair of this lane in the ring of kRing pairs
            const unsigned p1 = ring1w[j], p2 = ring2w[j];
            unsigned o1a = p1 & 0xffffu;
            const unsigned o1b = p1 >> 16, o2a = p2 & 0xffffu, o2b = p2 >> 16;
            unsigned hA_up = __shfl_up_sync(FULL, hA_last, 1);
            unsigned hB_up = __shfl_up_sync(FULL, hB_last, 1);
            unsigned fA = __shfl_up_sync(FULL, fA_out, 1);
            unsigned fB = __shfl_up_sync(FULL, fB_out, 1);
            unsigned b_in = __shfl_up_sync(FULL, b_out, 1);
            if (lane == 0) { hA_up = 0; hB_up = 0; fA = 0; fB = 0; b_in = 0; }
            if (o1a & kFlag) {                                       // first column of a pair (or the sentinel)
                o1a &= 0x7fffu;
                b_out = __vmaxs2(best, b_in);
                if (lane == 31) {
                    if (n31 > 0) {
                        const uint2 d = desc[(n31 - 1) & (kDescRing2 - 1)];
                        const int s1 = (int)(b_out & 0xffffu), s2 = (int)(b_out >> 16);
                        if (s1 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = d.x; else P.out[d.x] = s1;
                        if (d.y != kNoTarget) { if (s2 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = d.y; else P.out[d.y] = s2; }
                    }
                    ++n31;
                }
                best = 0; diag_in = 0;
#pragma unroll
                for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
            }
            unsigned wa1[KW], wa2[KW], wb1[KW], wb2[KW];
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                wa1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1a + m * 128);
                wa2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2a + m * 128);
                wb1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1b + m * 128);
                wb2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2b + m * 128);
            }
            // cell A = (row, first column), cell B = (row, second column); recurrences as in score_pair_packed
            unsigned tA = __vadd2(diag_in, prmt(wa1[0], wa2[0], 0xC480u));       // H(row above, column before A) + S
            unsigned tB = __vadd2(hA_up, prmt(wb1[0], wb2[0], 0xC480u));         // H(row above, A) + S
            diag_in = hB_up;
#pragma unroll
            for (int r = 0; r < K; ++r) {
                unsigned tA_next = 0, tB_next = 0;
                const unsigned sel = ((r + 1) & 3) == 0 ? 0xC480u : ((r + 1) & 3) == 1 ? 0xD591u : ((r + 1) & 3) == 2 ? 0xE6A2u : 0xF7B3u;
                if (r + 1 < K) tA_next = __vadd2(H[r], prmt(wa1[(r + 1) >> 2], wa2[(r + 1) >> 2], sel));
                const unsigned hA = __vimax3_s16x2_relu(tA, E[r], fA);
                const unsigned hqA = __vadd2(hA, negQ);
                const unsigned eB = __viaddmax_s16x2(E[r], negR, hqA);           // E of column B
                fA = __viaddmax_s16x2(fA, negR, hqA);
                if (r + 1 < K) tB_next = __vadd2(hA, prmt(wb1[(r + 1) >> 2], wb2[(r + 1) >> 2], sel));
                const unsigned hB = __vimax3_s16x2_relu(tB, eB, fB);
                H[r] = hB;
                const unsigned hqB = __vadd2(hB, negQ);
                E[r] = __viaddmax_s16x2(eB, negR, hqB);                          // E of the next step's column A
                fB = __viaddmax_s16x2(fB, negR, hqB);
                best = __vimax3_s16x2(best, tA, tB);
                if (r == K - 1) hA_last = hA;
                tA = tA_next; tB = tB_next;
            }
            hB_last = H[K - 1];
            fA_out = fA; fB_out = fB;
        }
        __syncwarp();
        if (end_col != kOpen && S0 + kSteps >= (end_col >> 1) + 32) break;
    }
}

// Build the int8 profile of one query into shared memory: prof[letter][m][lane] words, byte b of word
// (m, lane) = S[q[lane*K + 4m + b]][letter]; rows beyond the query read 0; pad letter row reads mat8 row 26.
template <int K>
__device__ void build_profile(unsigned* prof, const int8_t* smat, const uint8_t* q, int qlen) {
    constexpr int KW = (K + 3) / 4;
    for (int w = threadIdx.x; w < (S4G_PAD_CODE + 1) * KW * 32; w += blockDim.x) {
        const int lane = w & 31, m = (w >> 5) % KW, letter = w / (KW * 32);
        unsigned word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int rr = 4 * m + b;
            const int row = lane * K + rr;
            int v = 0;
            if (rr < K && row < qlen) v = smat[letter * 32 + q[row]];
            word |= (unsigned)(v & 0xff) << (8 * b);
        }
        prof[w] = word;
    }
}

template <int K, int MODE>
__device__ void run_tile(const ScoreParams& P, unsigned* prof, const int8_t* smat, unsigned short* rings,
                         int* s_next, int q, int pair_begin, int pair_end) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t qo = P.q_off[q];
    const int qlen = (int)(P.q_off[q + 1] - qo);
    build_profile<K>(prof, smat, P.q_codes + qo, qlen);
    if (threadIdx.x == 0) *s_next = pair_begin;
    __syncthreads();
    const int64_t cbeg = P.cand_off[q], cend = P.cand_off[q + 1];
    unsigned short* ring1 = rings + warp * (4 * kRing);
    unsigned short* ring2 = ring1 + 2 * kRing;
    const unsigned negQ = ((unsigned)(-P.gap_open) & 0xffffu) * 0x10001u;
    const unsigned negR = ((unsigned)(-P.gap_extend) & 0xffffu) * 0x10001u;
    constexpr bool TRACK = MODE == 1;
    if (MODE == 2) {
        uint2* desc = reinterpret_cast<uint2*>(rings + kWarps * (4 * kRing)) + warp * kDescRing2;
        stream_pairs_packed2<K>(P, prof + lane, ring1, ring2, desc, s_next, cbeg, cend, pair_end, negQ, negR, lane);
        __syncthreads();
        return;
    }
    if (MODE == 0) {
        uint2* desc = reinterpret_cast<uint2*>(rings + kWarps * (4 * kRing)) + warp * kDescRing;
        stream_pairs_packed<K>(P, prof + lane, ring1, ring2, desc, s_next, cbeg, cend, pair_end, negQ, negR, lane);
        __syncthreads();
        return;
    }
    while (true) {
        int p = 0;
        if (lane == 0) p = atomicAdd(s_next, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= pair_end) break;
        const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
        const uint32_t c1 = P.sorted_idx[i1];
        const bool has2 = i2 < cend;
        const uint32_t c2 = has2 ? P.sorted_idx[i2] : c1;
        const int64_t a1 = P.db_off[P.cand_ids[c1] - P.id_base], b1 = P.db_off[P.cand_ids[c1] - P.id_base + 1];
        const int64_t a2 = P.db_off[P.cand_ids[c2] - P.id_base], b2 = P.db_off[P.cand_ids[c2] - P.id_base + 1];
        if (TRACK) {
            // missing second target: a score no cell can reach
            const unsigned sc2 = ((unsigned)P.pair_score[c1] & 0xffffu) | (has2 ? (unsigned)P.pair_score[c2] << 16 : 0x7fff0000u);
            unsigned f1, f2;
            score_pair_packed<K, true>(prof + lane, ring1, ring2, P.db_codes + a1, (int)(b1 - a1), P.db_codes + a2,
                                       has2 ? (int)(b2 - a2) : 0, negQ, negR, lane, sc2, qlen, &f1, &f2);
            if (lane == 0) {
                if (f1 == kNotFound) atomicOr(&P.counters[3], 1ull);
                P.out[4 * (int64_t)c1 + 1] = f1 == kNotFound ? -1 : (int)(f1 & 1023u);
                P.out[4 * (int64_t)c1 + 3] = f1 == kNotFound ? -1 : (int)(f1 >> 10);
                if (has2) {
                    if (f2 == kNotFound) atomicOr(&P.counters[3], 1ull);
                    P.out[4 * (int64_t)c2 + 1] = f2 == kNotFound ? -1 : (int)(f2 & 1023u);
                    P.out[4 * (int64_t)c2 + 3] = f2 == kNotFound ? -1 : (int)(f2 >> 10);
                }
            }
            continue;
        }
        const unsigned best = score_pair_packed<K, false>(prof + lane, ring1, ring2, P.db_codes + a1, (int)(b1 - a1),
                                                          P.db_codes + a2, has2 ? (int)(b2 - a2) : 0, negQ, negR, lane);
        if (lane == 0) {
            const int s1 = (int)(best & 0xffffu), s2 = (int)(best >> 16);
            if (s1 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = c1; else P.out[c1] = s1;
            if (has2) { if (s2 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = c2; else P.out[c2] = s2; }
        }
    }
    __syncthreads();
}

template <int MODE>
__device__ __forceinline__ void packed_kernel_body(const ScoreParams& P) {
    constexpr bool TRACK = MODE == 1;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* prof = reinterpret_cast<unsigned*>(smem);                                   // 27*8*32 words max
    int8_t* smat = reinterpret_cast<int8_t*>(smem + (S4G_PAD_CODE + 1) * 8 * 32 * 4);      // 27*32
    unsigned short* rings = reinterpret_cast<unsigned short*>(smat + (S4G_PAD_CODE + 1) * 32);
    __shared__ int s_next;
    __shared__ long long s_tile;
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    const long long total = P.tile_start[P.nq];
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(&P.counters[0], 1ull);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= total) break;
        // query of this tile: last q with tile_start[q] <= tile
        int lo = 0, hi = P.nq;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P.tile_start[mid] <= tile) lo = mid; else hi = mid; }
        const int q = P.q_order ? P.q_order[lo] : lo;
        const int qlen = (int)(P.q_off[q + 1] - P.q_off[q]);
        const long long n_c = P.cand_off[q + 1] - P.cand_off[q];
        const int n_pairs = (int)((n_c + 1) >> 1);
        constexpr int TP = TRACK ? kTrackTilePairs : kStreamTilePairs;
        const int pb = (int)(tile - P.tile_start[lo]) * TP;
        const int pe = pb + TP < n_pairs ? pb + TP : n_pairs;
        // rows per lane: exactly ceil(qlen / 32) (every class from 2 to 32 has its own instantiation: an even-only dispatch pads
        // the average configs[1] query by 32 rows, 5.8 percent of the cells)
        const int K = (qlen + 31) >> 5;
        switch (K) {
            case 0: case 1: case 2: run_tile<2, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 3: run_tile<3, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 4: run_tile<4, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 5: run_tile<5, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 6: run_tile<6, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 7: run_tile<7, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 8: run_tile<8, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 9: run_tile<9, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 10: run_tile<10, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 11: run_tile<11, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 12: run_tile<12, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 13: run_tile<13, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 14: run_tile<14, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 15: run_tile<15, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 16: run_tile<16, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 17: run_tile<17, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 18: run_tile<18, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 19: run_tile<19, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 20: run_tile<20, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 21: run_tile<21, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 22: run_tile<22, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 23: run_tile<23, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 24: run_tile<24, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 25: run_tile<25, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 26: run_tile<26, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 27: run_tile<27, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 28: run_tile<28, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 29: run_tile<29, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 30: run_tile<30, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            case 31: run_tile<31, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
            default: run_tile<32, MODE>(P, prof, smat, rings, &s_next, q, pb, pe); break;
        }
    }
}

__global__ void __launch_bounds__(kWarps * 32, 2) sw_score_packed_kernel(ScoreParams P) { packed_kernel_body<0>(P); }
// the same with two stream columns per step (stream_pairs_packed2): the shipped form
__global__ void __launch_bounds__(kWarps * 32, 2) sw_score_packed2_kernel(ScoreParams P) { packed_kernel_body<2>(P); }

// End cells of the kept hits with the same systolic sweep (stage 3, step 1).
__global__ void __launch_bounds__(kWarps * 32, 2) al_forward_packed_kernel(ScoreParams P) { packed_kernel_body<1>(P); }

// ------------------------------------------------------------------------------------------------------
// striped kernel for long queries (intra-sequence): the query is cut into stripes of 32 * kMaxK = 1024 rows; a stripe
// is swept exactly like a short query (same packed cell update, two targets per warp, profile of the stripe shared by
// the CTA), and the last row of a stripe (H and F, packed for both targets) is handed to the next stripe through a
// per-pair boundary buffer in global memory, staged through shared memory 64 columns at a time.

struct StripSmem {
    unsigned short *ring1, *ring2;    // [128] profile row offsets of the two targets
    unsigned *inH, *inF;              // [64]  boundary of the previous stripe for the current block of columns
    unsigned *outH, *outF;            // [128] boundary produced by lane 31
};
constexpr int kStripWarpBytes = 4 * kRing * 2 + 2 * kRing * 4 + 4 * kRing * 4;
// streaming form (stream_stripe): rings of 2 * kRing stream columns for the target offsets (2 x u16), the incoming boundary (H, F),
// the outgoing boundary (H, F) and its destination (u32), plus the pairs in flight
constexpr int kStreamStripWarpBytes = 2 * (2 * kRing) * 2 + 5 * (2 * kRing) * 4 + 32 * 4;      // 32 = kDescRing2 descriptors (the one-column form uses 16)

__device__ __forceinline__ unsigned sweep_stripe(const unsigned* __restrict__ prof_lane, const StripSmem& S, const uint8_t* __restrict__ t1,
                                                 int len1, const uint8_t* __restrict__ t2, int len2, unsigned negQ, unsigned negR,
                                                 unsigned* bH, unsigned* bF, bool first, bool last, int lane) {
    constexpr int K = kMaxK, KW = K / 4;
    constexpr unsigned kRowBytes = KW * 128;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr int kRingMask = 2 * kRing - 1;
    const unsigned FULL = 0xffffffffu;
    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0;
    const int maxlen = len1 > len2 ? len1 : len2;
    const int nsteps = maxlen + 31;
    for (int c = lane; c < kRing; c += 32) { S.ring1[kRing + c] = kPadOff; S.ring2[kRing + c] = kPadOff; }
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
    int flushed = 0;
    for (int s0 = 0; s0 < nsteps; s0 += kRing) {
        if (!last) {
            const int upto = min(maxlen, s0 - 31);
            for (int c = flushed + lane; c < upto; c += 32) { bH[c] = S.outH[c & kRingMask]; bF[c] = S.outF[c & kRingMask]; }
            if (upto > flushed) flushed = upto;
        }
        {
            const int base = s0 & kRingMask;
#pragma unroll
            for (int c = 0; c < kRing; c += 32) {
                const int j = s0 + c + lane;
                S.ring1[base + c + lane] = (unsigned short)(j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff);
                S.ring2[base + c + lane] = (unsigned short)(j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff);
                if (!first) { S.inH[c + lane] = j < maxlen ? bH[j] : 0u; S.inF[c + lane] = j < maxlen ? bF[j] : 0u; }
            }
        }
        __syncwarp();
        const int send = (nsteps - s0) < kRing ? (nsteps - s0) : kRing;
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int j = s0 + ss - lane;
            const unsigned o1 = S.ring1[j & kRingMask];
            const unsigned o2 = S.ring2[j & kRingMask];
            unsigned w1[KW], w2[KW];
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                w1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1 + m * 128);
                w2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2 + m * 128);
            }
            unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
            unsigned f = __shfl_up_sync(FULL, f_out, 1);
            if (lane == 0) { h_up = first ? 0u : S.inH[ss]; f = first ? 0u : S.inF[ss]; }
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
                t_prev = t;
                t = t_next;
            }
            h_last = H[K - 1];
            f_out = f;
            if (lane == 31 && !last && j >= 0 && j < maxlen) { S.outH[j & kRingMask] = h_last; S.outF[j & kRingMask] = f_out; }
        }
        __syncwarp();
    }
    if (!last) for (int c = flushed + lane; c < maxlen; c += 32) { bH[c] = S.outH[c & kRingMask]; bF[c] = S.outF[c & kRingMask]; }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = __vmaxs2(best, __shfl_xor_sync(FULL, best, o));
    return best;
}

// Streaming form of a stripe sweep: the pairs a warp takes from the tile are laid end to end as one stream of target columns
// (see stream_pairs_packed), so the 31-step fill/drain of the wavefront is paid once per warp, tile and stripe instead of once
// per pair and stripe -- the candidates of a titin-like query are ~100-residue targets, where the drain is a quarter of all
// steps.  Per stream column the rings carry the two profile offsets, the boundary values coming in from the previous stripe
// (lane 0) and, for what lane 31 produces, the place in the CTA's boundary buffer they go to.
struct StreamStripSmem {
    unsigned short *ring1, *ring2;    // [2 kRing]
    unsigned *inH, *inF;              // [2 kRing]
    unsigned *outH, *outF, *oaddr;    // [2 kRing]
    unsigned* desc;                   // [kDescRing] tile-relative pair index of the pairs in flight
};

__device__ __forceinline__ void stream_stripe(const ScoreParams& P, const unsigned* __restrict__ prof_lane, const StreamStripSmem& S, int* s_next,
                                              unsigned* s_best, unsigned* cta_bound, int64_t cbeg, int64_t cend, int pb, int pe, unsigned negQ,
                                              unsigned negR, bool first, bool last, int lane) {
    constexpr int K = kMaxK, KW = K / 4;
    constexpr unsigned kRowBytes = KW * 128;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr unsigned kFlag = 0x8000u;
    constexpr int kRingMask = 2 * kRing - 1;
    constexpr int kOpen = 0x7fffffff;
    constexpr unsigned kNoAddr = 0xffffffffu;
    const unsigned FULL = 0xffffffffu;
    const unsigned strip_cols = (unsigned)P.strip_cols;

    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0, b_out = 0;
    int n31 = 0;
    const uint8_t *t1 = nullptr, *t2 = nullptr;
    int len1 = 0, len2 = 0, L = 0, pos = 0, n_pulled = 0, end_col = kOpen, cur_pair = 0;
    for (int c = lane; c < kRing; c += 32) { S.ring1[kRing + c] = kPadOff; S.ring2[kRing + c] = kPadOff; S.inH[kRing + c] = 0u; S.inF[kRing + c] = 0u; S.oaddr[kRing + c] = kNoAddr; }
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
    int flushed = 0;
    auto flush_out = [&](int upto) {          // boundary values of the stream columns lane 31 has finished -> the pairs' boundary rows
        if (last) return;
        for (int c = flushed + lane; c < upto; c += 32) {
            const unsigned a = S.oaddr[c & kRingMask];
            if (a != kNoAddr) { cta_bound[a] = S.outH[c & kRingMask]; cta_bound[a + strip_cols] = S.outF[c & kRingMask]; }
        }
        if (upto > flushed) flushed = upto;
    };

    for (int s0 = 0;; s0 += kRing) {
        flush_out(s0 - 31);
        __syncwarp();
        // ---- stage stream columns s0 .. s0+kRing-1
        int col = s0;
        while (col < s0 + kRing) {
            if (pos >= L && end_col == kOpen) {
                int p = 0;
                if (lane == 0) p = atomicAdd(s_next, 1);
                p = __shfl_sync(FULL, p, 0);
                if (p < pe) {
                    const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
                    const uint32_t c1 = P.sorted_idx[i1];
                    const bool has2 = i2 < cend;
                    const uint32_t c2 = has2 ? P.sorted_idx[i2] : c1;
                    const uint32_t g1 = P.cand_ids[c1] - P.id_base, g2 = P.cand_ids[c2] - P.id_base;
                    const int64_t a1 = P.db_off[g1], b1 = P.db_off[g1 + 1];
                    const int64_t a2 = P.db_off[g2], b2 = P.db_off[g2 + 1];
                    len1 = (int)(b1 - a1); len2 = has2 ? (int)(b2 - a2) : 0;
                    if (len1 > P.strip_cols || len2 > P.strip_cols) {       // too long for the boundary buffer: 32-bit kernel
                        if (lane == 0) s_best[p - pb] = 0x7fff7fffu;
                        pos = 0; L = 0;
                        continue;
                    }
                    t1 = P.db_codes + a1; t2 = P.db_codes + a2;
                    L = len1 > len2 ? len1 : len2;
                    if (L < kMinCols) L = kMinCols;
                    pos = 0;
                    cur_pair = p - pb;
                    if (lane == 0) S.desc[n_pulled & (kDescRing - 1)] = (unsigned)cur_pair;
                    ++n_pulled;
                } else {
                    end_col = col;
                }
            }
            if (end_col != kOpen) {
                for (int c = col + lane; c < s0 + kRing; c += 32) {
                    S.ring1[c & kRingMask] = (unsigned short)(c == end_col ? (kPadOff | kFlag) : kPadOff);
                    S.ring2[c & kRingMask] = (unsigned short)kPadOff;
                    S.inH[c & kRingMask] = 0u; S.inF[c & kRingMask] = 0u; S.oaddr[c & kRingMask] = kNoAddr;
                }
                break;
            }
            const int n = min(L - pos, s0 + kRing - col);
            const unsigned row = (unsigned)cur_pair * 2u * strip_cols;
            const int maxlen = len1 > len2 ? len1 : len2;
            for (int i = lane; i < n; i += 32) {
                const int j = pos + i;
                unsigned o1 = j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff;
                const unsigned o2 = j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff;
                if (j == 0) o1 |= kFlag;
                const int x = (col + i) & kRingMask;
                S.ring1[x] = (unsigned short)o1;
                S.ring2[x] = (unsigned short)o2;
                const bool real = j < maxlen;
                S.inH[x] = (!first && real) ? cta_bound[row + j] : 0u;
                S.inF[x] = (!first && real) ? cta_bound[row + strip_cols + j] : 0u;
                S.oaddr[x] = real ? row + (unsigned)j : kNoAddr;
            }
            col += n; pos += n;
        }
        if (n_pulled == 0) return;
        __syncwarp();
        const int send = end_col == kOpen ? kRing : min(kRing, end_col + 32 - s0);
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int j = (s0 + ss - lane) & kRingMask;
            int o1 = (short)S.ring1[j];
            const unsigned o2 = S.ring2[j];
            unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
            unsigned f = __shfl_up_sync(FULL, f_out, 1);
            unsigned b_in = __shfl_up_sync(FULL, b_out, 1);
            if (lane == 0) { h_up = S.inH[j]; f = S.inF[j]; b_in = 0; }
            if (o1 < 0) {                                            // first column of a pair (or the sentinel)
                o1 &= 0x7fff;
                b_out = __vmaxs2(best, b_in);
                if (lane == 31) {
                    if (n31 > 0) { const unsigned d = S.desc[(n31 - 1) & (kDescRing - 1)]; s_best[d] = __vmaxs2(s_best[d], b_out); }
                    ++n31;
                }
                best = 0; diag_in = 0;
#pragma unroll
                for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
            }
            unsigned w1[KW], w2[KW];
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                w1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1 + m * 128);
                w2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2 + m * 128);
            }
            unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;
            diag_in = h_up;
#pragma unroll
            for (int r = 0; r < K; ++r) {                            // cell update: see score_pair_packed
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
                t_prev = t;
                t = t_next;
            }
            h_last = H[K - 1];
            f_out = f;
            if (lane == 31 && !last) { S.outH[j] = h_last; S.outF[j] = f_out; }
        }
        __syncwarp();
        if (end_col != kOpen && s0 + kRing >= end_col + 32) { flush_out(end_col); break; }
    }
    __syncwarp();
}

// The same with two stream columns per step (see stream_pairs_packed2: no register moves, shuffles and ring reads once per two
// columns): lane 0 takes the boundary values of both columns from the in-rings (one 8-byte read each), lane 31 leaves both in the
// out-rings; lane 31 runs 62 columns behind lane 0, so finished columns are flushed up to s0 - 62 and 32 pairs can be in flight.
__device__ __forceinline__ void stream_stripe2(const ScoreParams& P, const unsigned* __restrict__ prof_lane, const StreamStripSmem& S, int* s_next,
                                              unsigned* s_best, unsigned* cta_bound, int64_t cbeg, int64_t cend, int pb, int pe, unsigned negQ,
                                              unsigned negR, bool first, bool last, int lane) {
    constexpr int K = kMaxK, KW = K / 4;
    constexpr unsigned kRowBytes = KW * 128;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr unsigned kFlag = 0x8000u;
    constexpr int kRingMask = 2 * kRing - 1;
    constexpr int kSteps = kRing / 2;     // steps per refill (two stream columns per step)
    constexpr int kOpen = 0x7fffffff;
    constexpr unsigned kNoAddr = 0xffffffffu;
    const unsigned FULL = 0xffffffffu;
    const unsigned strip_cols = (unsigned)P.strip_cols;

    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, hA_last = 0, hB_last = 0, fA_out = 0, fB_out = 0, diag_in = 0, b_out = 0;
    int n31 = 0;
    const uint8_t *t1 = nullptr, *t2 = nullptr;
    int len1 = 0, len2 = 0, L = 0, pos = 0, n_pulled = 0, end_col = kOpen, cur_pair = 0;
    for (int c = lane; c < kRing; c += 32) { S.ring1[kRing + c] = kPadOff; S.ring2[kRing + c] = kPadOff; S.inH[kRing + c] = 0u; S.inF[kRing + c] = 0u; S.oaddr[kRing + c] = kNoAddr; }
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
    int flushed = 0;
    auto flush_out = [&](int upto) {          // boundary values of the stream columns lane 31 has finished -> the pairs' boundary rows
        if (last) return;
        for (int c = flushed + lane; c < upto; c += 32) {
            const unsigned a = S.oaddr[c & kRingMask];
            if (a != kNoAddr) { cta_bound[a] = S.outH[c & kRingMask]; cta_bound[a + strip_cols] = S.outF[c & kRingMask]; }
        }
        if (upto > flushed) flushed = upto;
    };

    for (int s0 = 0;; s0 += kRing) {
        flush_out(s0 - 62);
        __syncwarp();
        // ---- stage stream columns s0 .. s0+kRing-1
        int col = s0;
        while (col < s0 + kRing) {
            if (pos >= L && end_col == kOpen) {
                int p = 0;
                if (lane == 0) p = atomicAdd(s_next, 1);
                p = __shfl_sync(FULL, p, 0);
                if (p < pe) {
                    const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
                    const uint32_t c1 = P.sorted_idx[i1];
                    const bool has2 = i2 < cend;
                    const uint32_t c2 = has2 ? P.sorted_idx[i2] : c1;
                    const uint32_t g1 = P.cand_ids[c1] - P.id_base, g2 = P.cand_ids[c2] - P.id_base;
                    const int64_t a1 = P.db_off[g1], b1 = P.db_off[g1 + 1];
                    const int64_t a2 = P.db_off[g2], b2 = P.db_off[g2 + 1];
                    len1 = (int)(b1 - a1); len2 = has2 ? (int)(b2 - a2) : 0;
                    if (len1 > P.strip_cols || len2 > P.strip_cols) {       // too long for the boundary buffer: 32-bit kernel
                        if (lane == 0) s_best[p - pb] = 0x7fff7fffu;
                        pos = 0; L = 0;
                        continue;
                    }
                    t1 = P.db_codes + a1; t2 = P.db_codes + a2;
                    L = len1 > len2 ? len1 : len2;
                    L = (L + 1) & ~1;                                // boundaries on even stream columns; a pad column changes no maximum
                    if (L < kMinCols) L = kMinCols;
                    pos = 0;
                    cur_pair = p - pb;
                    if (lane == 0) S.desc[n_pulled & (kDescRing2 - 1)] = (unsigned)cur_pair;
                    ++n_pulled;
                } else {
                    end_col = col;
                }
            }
            if (end_col != kOpen) {
                for (int c = col + lane; c < s0 + kRing; c += 32) {
                    S.ring1[c & kRingMask] = (unsigned short)(c == end_col ? (kPadOff | kFlag) : kPadOff);
                    S.ring2[c & kRingMask] = (unsigned short)kPadOff;
                    S.inH[c & kRingMask] = 0u; S.inF[c & kRingMask] = 0u; S.oaddr[c & kRingMask] = kNoAddr;
                }
                break;
            }
            const int n = min(L - pos, s0 + kRing - col);
            const unsigned row = (unsigned)cur_pair * 2u * strip_cols;
            const int maxlen = len1 > len2 ? len1 : len2;
            for (int i = lane; i < n; i += 32) {
                const int j = pos + i;
                unsigned o1 = j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff;
                const unsigned o2 = j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff;
                if (j == 0) o1 |= kFlag;
                const int x = (col + i) & kRingMask;
                S.ring1[x] = (unsigned short)o1;
                S.ring2[x] = (unsigned short)o2;
                const bool real = j < maxlen;
                S.inH[x] = (!first && real) ? cta_bound[row + j] : 0u;
                S.inF[x] = (!first && real) ? cta_bound[row + strip_cols + j] : 0u;
                S.oaddr[x] = real ? row + (unsigned)j : kNoAddr;
            }
            col += n; pos += n;
        }
        if (n_pulled == 0) return;
        __syncwarp();
        const int S0 = s0 >> 1;
        const int send = end_col == kOpen ? kSteps : min(kSteps, (end_col >> 1) + 32 - S0);
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int jp = (S0 + ss - lane) & (kRing - 1);           // column pair of this lane; its columns sit at ring slots 2 jp, 2 jp + 1
            const unsigned p1 = reinterpret_cast<const unsigned*>(S.ring1)[jp], p2 = reinterpret_cast<const unsigned*>(S.ring2)[jp];
            unsigned o1a = p1 & 0xffffu;
            const unsigned o1b = p1 >> 16, o2a = p2 & 0xffffu, o2b = p2 >> 16;
            unsigned hA_up = __shfl_up_sync(FULL, hA_last, 1);
            unsigned hB_up = __shfl_up_sync(FULL, hB_last, 1);
            unsigned fA = __shfl_up_sync(FULL, fA_out, 1);
            unsigned fB = __shfl_up_sync(FULL, fB_out, 1);
            unsigned b_in = __shfl_up_sync(FULL, b_out, 1);
            if (lane == 0) {
                const uint2 ih = reinterpret_cast<const uint2*>(S.inH)[jp], jf = reinterpret_cast<const uint2*>(S.inF)[jp];
                hA_up = ih.x; hB_up = ih.y; fA = jf.x; fB = jf.y; b_in = 0;
            }
            if (o1a & kFlag) {                                       // first column of a pair (or the sentinel)
                o1a &= 0x7fffu;
                b_out = __vmaxs2(best, b_in);
                if (lane == 31) {
                    if (n31 > 0) { const unsigned d = S.desc[(n31 - 1) & (kDescRing2 - 1)]; s_best[d] = __vmaxs2(s_best[d], b_out); }
                    ++n31;
                }
                best = 0; diag_in = 0;
#pragma unroll
                for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
            }
            unsigned wa1[KW], wa2[KW], wb1[KW], wb2[KW];
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                wa1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1a + m * 128);
                wa2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2a + m * 128);
                wb1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1b + m * 128);
                wb2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2b + m * 128);
            }
            unsigned tA = __vadd2(diag_in, prmt(wa1[0], wa2[0], 0xC480u));       // cells A / B of a row: see stream_pairs_packed2
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
            if (lane == 31 && !last) {
                reinterpret_cast<uint2*>(S.outH)[jp] = make_uint2(hA_last, hB_last);
                reinterpret_cast<uint2*>(S.outF)[jp] = make_uint2(fA_out, fB_out);
            }
        }
        __syncwarp();
        if (end_col != kOpen && S0 + kSteps >= (end_col >> 1) + 32) { flush_out(end_col); break; }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kWarps * 32, 2) sw_score_striped_kernel(ScoreParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* prof = reinterpret_cast<unsigned*>(smem);
    int8_t* smat = reinterpret_cast<int8_t*>(smem + (S4G_PAD_CODE + 1) * 8 * 32 * 4);
    unsigned char* wbase = reinterpret_cast<unsigned char*>(smat + (S4G_PAD_CODE + 1) * 32);
    __shared__ long long s_tile;
    __shared__ unsigned s_best[kTilePairs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StripSmem S;
    {
        unsigned char* wb = wbase + warp * kStripWarpBytes;
        S.ring1 = reinterpret_cast<unsigned short*>(wb);
        S.ring2 = S.ring1 + 2 * kRing;
        S.inH = reinterpret_cast<unsigned*>(S.ring2 + 2 * kRing);
        S.inF = S.inH + kRing;
        S.outH = S.inF + kRing;
        S.outF = S.outH + 2 * kRing;
    }
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    const long long total = P.long_tile_start[P.nq];
    const unsigned negQ = ((unsigned)(-P.gap_open) & 0xffffu) * 0x10001u;
    const unsigned negR = ((unsigned)(-P.gap_extend) & 0xffffu) * 0x10001u;
    unsigned* cta_bound = P.strip_bound + (size_t)blockIdx.x * kTilePairs * 2 * P.strip_cols;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(&P.counters[4], 1ull);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= total) break;
        int lo = 0, hi = P.nq;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P.long_tile_start[mid] <= tile) lo = mid; else hi = mid; }
        const int q = lo;
        const int64_t qo = P.q_off[q];
        const int qlen = (int)(P.q_off[q + 1] - qo);
        const int64_t cbeg = P.cand_off[q], cend = P.cand_off[q + 1];
        const int n_pairs = (int)((cend - cbeg + 1) >> 1);
        const int pb = (int)(tile - P.long_tile_start[q]) * kTilePairs;
        const int pe = pb + kTilePairs < n_pairs ? pb + kTilePairs : n_pairs;
        const int npass = (qlen + 32 * kMaxK - 1) / (32 * kMaxK);
        for (int i = threadIdx.x; i < kTilePairs; i += blockDim.x) s_best[i] = 0;
        for (int pass = 0; pass < npass; ++pass) {
            __syncthreads();
            build_profile<kMaxK>(prof, smat, P.q_codes + qo + (int64_t)pass * 32 * kMaxK, qlen - pass * 32 * kMaxK);
            __syncthreads();
            for (int p = pb + warp; p < pe; p += kWarps) {
                const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
                const uint32_t c1 = P.sorted_idx[i1];
                const bool has2 = i2 < cend;
                const uint32_t c2 = has2 ? P.sorted_idx[i2] : c1;
                const int64_t a1 = P.db_off[P.cand_ids[c1] - P.id_base], b1 = P.db_off[P.cand_ids[c1] - P.id_base + 1];
                const int64_t a2 = P.db_off[P.cand_ids[c2] - P.id_base], b2 = P.db_off[P.cand_ids[c2] - P.id_base + 1];
                const int len1 = (int)(b1 - a1), len2 = has2 ? (int)(b2 - a2) : 0;
                if (len1 > P.strip_cols || len2 > P.strip_cols) {       // too long for the boundary buffer: 32-bit kernel
                    if (lane == 0) s_best[p - pb] = 0x7fff7fffu;
                    continue;
                }
                unsigned* bH = cta_bound + (size_t)(p - pb) * 2 * P.strip_cols;
                const unsigned best = sweep_stripe(prof + lane, S, P.db_codes + a1, len1, P.db_codes + a2, len2, negQ, negR, bH, bH + P.strip_cols,
                                                   pass == 0, pass == npass - 1, lane);
                if (lane == 0) s_best[p - pb] = __vmaxs2(s_best[p - pb], best);
            }
        }
        __syncthreads();
        for (int p = pb + threadIdx.x; p < pe; p += blockDim.x) {
            const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
            const uint32_t c1 = P.sorted_idx[i1];
            const unsigned best = s_best[p - pb];
            const int s1 = (int)(best & 0xffffu), s2 = (int)(best >> 16);
            if (s1 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = c1; else P.out[c1] = s1;
            if (i2 < cend) {
                const uint32_t c2 = P.sorted_idx[i2];
                if (s2 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = c2; else P.out[c2] = s2;
            }
        }
    }
}

// striped kernel, streaming form (TWO: two stream columns per step, the default; S4G_STRIPED=stream1 selects one column per step,
// S4G_STRIPED=pairs the pair-by-pair kernel above)
template <bool TWO>
__device__ __forceinline__ void striped_stream_body(const ScoreParams& P) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* prof = reinterpret_cast<unsigned*>(smem);
    int8_t* smat = reinterpret_cast<int8_t*>(smem + (S4G_PAD_CODE + 1) * 8 * 32 * 4);
    unsigned char* wbase = reinterpret_cast<unsigned char*>(smat + (S4G_PAD_CODE + 1) * 32);
    __shared__ long long s_tile;
    __shared__ int s_next;
    __shared__ unsigned s_best[kTilePairs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StreamStripSmem S;
    {
        unsigned char* wb = wbase + warp * kStreamStripWarpBytes;
        S.ring1 = reinterpret_cast<unsigned short*>(wb);
        S.ring2 = S.ring1 + 2 * kRing;
        S.inH = reinterpret_cast<unsigned*>(S.ring2 + 2 * kRing);
        S.inF = S.inH + 2 * kRing;
        S.outH = S.inF + 2 * kRing;
        S.outF = S.outH + 2 * kRing;
        S.oaddr = S.outF + 2 * kRing;
        S.desc = S.oaddr + 2 * kRing;
    }
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    const long long total = P.long_tile_start[P.nq];
    const unsigned negQ = ((unsigned)(-P.gap_open) & 0xffffu) * 0x10001u;
    const unsigned negR = ((unsigned)(-P.gap_extend) & 0xffffu) * 0x10001u;
    unsigned* cta_bound = P.strip_bound + (size_t)blockIdx.x * kTilePairs * 2 * P.strip_cols;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(&P.counters[4], 1ull);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= total) break;
        int lo = 0, hi = P.nq;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P.long_tile_start[mid] <= tile) lo = mid; else hi = mid; }
        const int q = lo;
        const int64_t qo = P.q_off[q];
        const int qlen = (int)(P.q_off[q + 1] - qo);
        const int64_t cbeg = P.cand_off[q], cend = P.cand_off[q + 1];
        const int n_pairs = (int)((cend - cbeg + 1) >> 1);
        const int pb = (int)(tile - P.long_tile_start[q]) * kTilePairs;
        const int pe = pb + kTilePairs < n_pairs ? pb + kTilePairs : n_pairs;
        const int npass = (qlen + 32 * kMaxK - 1) / (32 * kMaxK);
        for (int i = threadIdx.x; i < kTilePairs; i += blockDim.x) s_best[i] = 0;
        for (int pass = 0; pass < npass; ++pass) {
            __syncthreads();
            build_profile<kMaxK>(prof, smat, P.q_codes + qo + (int64_t)pass * 32 * kMaxK, qlen - pass * 32 * kMaxK);
            if (threadIdx.x == 0) s_next = pb;
            __syncthreads();
            if (TWO) stream_stripe2(P, prof + lane, S, &s_next, s_best, cta_bound, cbeg, cend, pb, pe, negQ, negR, pass == 0, pass == npass - 1, lane);
            else stream_stripe(P, prof + lane, S, &s_next, s_best, cta_bound, cbeg, cend, pb, pe, negQ, negR, pass == 0, pass == npass - 1, lane);
        }
        __syncthreads();
        for (int p = pb + threadIdx.x; p < pe; p += blockDim.x) {
            const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
            const uint32_t c1 = P.sorted_idx[i1];
            const unsigned best = s_best[p - pb];
            const int s1 = (int)(best & 0xffffu), s2 = (int)(best >> 16);
            if (s1 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = c1; else P.out[c1] = s1;
            if (i2 < cend) {
                const uint32_t c2 = P.sorted_idx[i2];
                if (s2 > P.ovf_limit) P.overflow[atomicAdd(&P.counters[1], 1ull)] = c2; else P.out[c2] = s2;
            }
        }
    }
}

__global__ void __launch_bounds__(kWarps * 32, 2) sw_score_striped_stream_kernel(ScoreParams P) { striped_stream_body<false>(P); }
__global__ void __launch_bounds__(kWarps * 32, 2) sw_score_striped_stream2_kernel(ScoreParams P) { striped_stream_body<true>(P); }

// ------------------------------------------------------------------------------------------------------
// End cells of the kept hits of LONG queries (stage 3, step 1): the striped sweep with the end-cell rule of
// score_pair_packed<K, true>.  SSW's end cell is the first column, then the first row in it, whose H equals the score
// (ssw.c:283-308,491-512).  Stripes are swept top down; a stripe that finds the score in column c leaves only the columns
// before c to the stripes below it (a cell further down in the same column has a larger row), so the column range -- and the
// boundary rows handed on -- shrink from stripe to stripe.  Two hits of a query per warp, 16-bit; hits under the swAlign rules
// (score > 32767) are left to al_sweep32_kernel.
__device__ __forceinline__ void track_stripe(const unsigned* __restrict__ prof_lane, const StripSmem& S, const uint8_t* __restrict__ t1, int len1,
                                             const uint8_t* __restrict__ t2, int len2, unsigned score2, int rows_valid, unsigned negQ, unsigned negR,
                                             unsigned* bH, unsigned* bF, bool first, bool last, int lane, unsigned* found1, unsigned* found2) {
    constexpr int K = kMaxK, KW = K / 4;
    constexpr unsigned kRowBytes = KW * 128;
    constexpr unsigned kPadOff = S4G_PAD_CODE * kRowBytes;
    constexpr int kRingMask = 2 * kRing - 1;
    const unsigned FULL = 0xffffffffu;
    unsigned H[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { H[r] = 0; E[r] = 0; }
    unsigned best = 0, h_last = 0, f_out = 0, diag_in = 0;
    unsigned fnd1 = kNotFound, fnd2 = kNotFound;
    const int maxlen = len1 > len2 ? len1 : len2;
    const int nsteps = maxlen + 31;
    for (int c = lane; c < kRing; c += 32) { S.ring1[kRing + c] = kPadOff; S.ring2[kRing + c] = kPadOff; }
    const char* prof_bytes = reinterpret_cast<const char*>(prof_lane);
    int flushed = 0, done_cols = 0;
    for (int s0 = 0; s0 < nsteps; s0 += kRing) {
        if (!last) {
            const int upto = min(maxlen, s0 - 31);
            for (int c = flushed + lane; c < upto; c += 32) { bH[c] = S.outH[c & kRingMask]; bF[c] = S.outF[c & kRingMask]; }
            if (upto > flushed) flushed = upto;
        }
        if (s0 > 0) {
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
                const int j = s0 + c + lane;
                S.ring1[base + c + lane] = (unsigned short)(j < len1 ? (unsigned)t1[j] * kRowBytes : kPadOff);
                S.ring2[base + c + lane] = (unsigned short)(j < len2 ? (unsigned)t2[j] * kRowBytes : kPadOff);
                if (!first) { S.inH[c + lane] = j < maxlen ? bH[j] : 0u; S.inF[c + lane] = j < maxlen ? bF[j] : 0u; }
            }
        }
        __syncwarp();
        const int send = (nsteps - s0) < kRing ? (nsteps - s0) : kRing;
#pragma unroll 1
        for (int ss = 0; ss < send; ++ss) {
            const int j = s0 + ss - lane;
            const unsigned o1 = S.ring1[j & kRingMask];
            const unsigned o2 = S.ring2[j & kRingMask];
            unsigned w1[KW], w2[KW];
#pragma unroll
            for (int m = 0; m < KW; ++m) {
                w1[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o1 + m * 128);
                w2[m] = *reinterpret_cast<const unsigned*>(prof_bytes + o2 + m * 128);
            }
            unsigned h_up = __shfl_up_sync(FULL, h_last, 1);
            unsigned f = __shfl_up_sync(FULL, f_out, 1);
            if (lane == 0) { h_up = first ? 0u : S.inH[ss]; f = first ? 0u : S.inF[ss]; }
            unsigned t = __vadd2(diag_in, prmt(w1[0], w2[0], 0xC480u)), t_prev = 0;
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
                t_prev = t;
                t = t_next;
            }
            h_last = H[K - 1];
            f_out = f;
            if (lane == 31 && !last && j >= 0 && j < maxlen) { S.outH[j & kRingMask] = h_last; S.outF[j & kRingMask] = f_out; }
            // the score is reached by a diagonal step (also across a stripe boundary: row 0 forms Hdiag + S from the boundary row), so
            // the column maximum of Hdiag + S gates the search for the row
            const unsigned eq = __vcmpeq2(best, score2);
            if (eq && j >= 0) {
                if ((eq & 0xffffu) && j < len1) {
                    int rr = -1;
#pragma unroll
                    for (int r = K - 1; r >= 0; --r) if ((H[r] & 0xffffu) == (score2 & 0xffffu) && lane * K + r < rows_valid) rr = r;
                    if (rr >= 0) fnd1 = min(fnd1, ((unsigned)j << 10) | (unsigned)(lane * K + rr));
                }
                if ((eq >> 16) && j < len2) {
                    int rr = -1;
#pragma unroll
                    for (int r = K - 1; r >= 0; --r) if ((H[r] >> 16) == (score2 >> 16) && lane * K + r < rows_valid) rr = r;
                    if (rr >= 0) fnd2 = min(fnd2, ((unsigned)j << 10) | (unsigned)(lane * K + rr));
                }
            }
        }
        done_cols = s0 + send - 31;
        __syncwarp();
    }
    if (!last) { const int upto = min(maxlen, done_cols); for (int c = flushed + lane; c < upto; c += 32) { bH[c] = S.outH[c & kRingMask]; bF[c] = S.outF[c & kRingMask]; } }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { fnd1 = min(fnd1, __shfl_xor_sync(FULL, fnd1, o)); fnd2 = min(fnd2, __shfl_xor_sync(FULL, fnd2, o)); }
    *found1 = fnd1; *found2 = fnd2;
}

__global__ void __launch_bounds__(kWarps * 32, 2) al_forward_striped_kernel(ScoreParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* prof = reinterpret_cast<unsigned*>(smem);
    int8_t* smat = reinterpret_cast<int8_t*>(smem + (S4G_PAD_CODE + 1) * 8 * 32 * 4);
    unsigned char* wbase = reinterpret_cast<unsigned char*>(smat + (S4G_PAD_CODE + 1) * 32);
    __shared__ long long s_tile;
    __shared__ int s_lim[kTilePairs][2];            // columns of each hit that can still hold the end cell
    __shared__ int s_col[kTilePairs][2], s_row[kTilePairs][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StripSmem S;
    {
        unsigned char* wb = wbase + warp * kStripWarpBytes;
        S.ring1 = reinterpret_cast<unsigned short*>(wb);
        S.ring2 = S.ring1 + 2 * kRing;
        S.inH = reinterpret_cast<unsigned*>(S.ring2 + 2 * kRing);
        S.inF = S.inH + kRing;
        S.outH = S.inF + kRing;
        S.outF = S.outH + 2 * kRing;
    }
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    const long long total = P.long_tile_start[P.nq];
    const unsigned negQ = ((unsigned)(-P.gap_open) & 0xffffu) * 0x10001u;
    const unsigned negR = ((unsigned)(-P.gap_extend) & 0xffffu) * 0x10001u;
    unsigned* cta_bound = P.strip_bound + (size_t)blockIdx.x * kTilePairs * 2 * P.strip_cols;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(&P.counters[4], 1ull);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= total) break;
        int lo = 0, hi = P.nq;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P.long_tile_start[mid] <= tile) lo = mid; else hi = mid; }
        const int q = lo;
        const int64_t qo = P.q_off[q];
        const int qlen = (int)(P.q_off[q + 1] - qo);
        const int64_t cbeg = P.cand_off[q], cend = P.cand_off[q + 1];
        const int n_pairs = (int)((cend - cbeg + 1) >> 1);
        const int pb = (int)(tile - P.long_tile_start[q]) * kTilePairs;
        const int pe = pb + kTilePairs < n_pairs ? pb + kTilePairs : n_pairs;
        const int npass = (qlen + 32 * kMaxK - 1) / (32 * kMaxK);
        for (int x = threadIdx.x; x < 2 * (pe - pb); x += blockDim.x) {
            const int p = pb + (x >> 1), h = x & 1;
            const int64_t i = cbeg + 2 * (int64_t)p + h;
            int lim = 0;
            if (i < cend) {
                const uint32_t c = P.sorted_idx[i];
                const uint32_t g = P.cand_ids[c] - P.id_base;
                const int len = (int)(P.db_off[g + 1] - P.db_off[g]);
                const int sc = P.pair_score[c];
                if (sc > 0 && sc <= 32767 && len <= P.strip_cols) lim = len;          // others: al_sweep32_kernel
            }
            s_lim[x >> 1][h] = lim; s_col[x >> 1][h] = -1; s_row[x >> 1][h] = -1;
        }
        for (int pass = 0; pass < npass; ++pass) {
            __syncthreads();
            build_profile<kMaxK>(prof, smat, P.q_codes + qo + (int64_t)pass * 32 * kMaxK, qlen - pass * 32 * kMaxK);
            __syncthreads();
            for (int p = pb + warp; p < pe; p += kWarps) {
                const int l1 = s_lim[p - pb][0], l2 = s_lim[p - pb][1];
                if (l1 == 0 && l2 == 0) continue;
                const int64_t i1 = cbeg + 2 * (int64_t)p, i2 = i1 + 1;
                const uint32_t c1 = P.sorted_idx[i1];
                const uint32_t c2 = i2 < cend ? P.sorted_idx[i2] : c1;
                const int64_t a1 = P.db_off[P.cand_ids[c1] - P.id_base], a2 = P.db_off[P.cand_ids[c2] - P.id_base];
                const unsigned sc2 = (l1 > 0 ? ((unsigned)P.pair_score[c1] & 0xffffu) : 0x7fffu) | ((l2 > 0 ? (unsigned)P.pair_score[c2] : 0x7fffu) << 16);
                unsigned* bH = cta_bound + (size_t)(p - pb) * 2 * P.strip_cols;
                unsigned f1, f2;
                track_stripe(prof + lane, S, P.db_codes + a1, l1, P.db_codes + a2, l2, sc2, min(32 * kMaxK, qlen - pass * 32 * kMaxK), negQ, negR, bH,
                             bH + P.strip_cols, pass == 0, pass == npass - 1, lane, &f1, &f2);
                if (lane == 0) {
                    if (f1 != kNotFound) { s_col[p - pb][0] = (int)(f1 >> 10); s_row[p - pb][0] = pass * 32 * kMaxK + (int)(f1 & 1023u); s_lim[p - pb][0] = (int)(f1 >> 10); }
                    if (f2 != kNotFound) { s_col[p - pb][1] = (int)(f2 >> 10); s_row[p - pb][1] = pass * 32 * kMaxK + (int)(f2 & 1023u); s_lim[p - pb][1] = (int)(f2 >> 10); }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int x = threadIdx.x; x < 2 * (pe - pb); x += blockDim.x) {
            const int p = pb + (x >> 1), h = x & 1;
            const int64_t i = cbeg + 2 * (int64_t)p + h;
            if (i >= cend) continue;
            const uint32_t c = P.sorted_idx[i];
            const uint32_t g = P.cand_ids[c] - P.id_base;
            const int len = (int)(P.db_off[g + 1] - P.db_off[g]);
            const int sc = P.pair_score[c];
            if (!(sc > 0 && sc <= 32767 && len <= P.strip_cols)) continue;
            if (s_col[x >> 1][h] < 0) { atomicOr(&P.counters[3], 1ull); continue; }
            P.out[4 * (int64_t)c + 1] = s_row[x >> 1][h];
            P.out[4 * (int64_t)c + 3] = s_col[x >> 1][h];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// exact 32-bit kernel: one warp per (query, target), any query length (256-row passes, the boundary row
// H/F travels through a per-warp global scratch that stays in L2).  Used for 16-bit overflow re-runs and
// for queries longer than 32*kMaxK rows.

struct GenericWork {
    const uint32_t* list;         // cand positions (overflow list) or nullptr = "all candidates of long queries"
    const unsigned long long* n_list;   // device count for `list`
    const uint32_t* long_idx;     // for the long-query path: sorted positions
    long long n_long;
};

__device__ __forceinline__ int find_query(const int64_t* cand_off, int nq, int64_t pos) {
    int lo = 0, hi = nq;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cand_off[mid] <= pos) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(kWarps * 32) sw_score_generic_kernel(ScoreParams P, GenericWork W) {
    __shared__ int8_t smat[(S4G_PAD_CODE + 1) * 32];
    for (int i = threadIdx.x; i < (S4G_PAD_CODE + 1) * 32; i += blockDim.x) smat[i] = P.mat8[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kWarps + (threadIdx.x >> 5);
    int32_t* bH = P.bound + (int64_t)gwarp * P.bound_stride;
    int32_t* bF = bH + P.bound_stride / 2;
    const unsigned FULL = 0xffffffffu;
    const long long n_work = W.list ? (long long)*W.n_list : W.n_long;
    const int Q = P.gap_open, R = P.gap_extend;
    while (true) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(&P.counters[2], 1ull);
        w = __shfl_sync(FULL, w, 0);
        if ((long long)w >= n_work) break;
        const uint32_t c = W.list ? W.list[w] : W.long_idx[w];
        const int q = find_query(P.cand_off, P.nq, (int64_t)c);
        const uint8_t* qs = P.q_codes + P.q_off[q];
        const int qlen = (int)(P.q_off[q + 1] - P.q_off[q]);
        const uint32_t tid = P.cand_ids[c] - P.id_base;
        const uint8_t* ts = P.db_codes + P.db_off[tid];
        const int tlen = (int)(P.db_off[tid + 1] - P.db_off[tid]);
        const int npass = (qlen + 32 * kGenK - 1) / (32 * kGenK);
        int best = 0;
        for (int pass = 0; pass < npass; ++pass) {
            const int row0 = pass * 32 * kGenK + lane * kGenK;
            int ql[kGenK];
#pragma unroll
            for (int r = 0; r < kGenK; ++r) ql[r] = row0 + r < qlen ? qs[row0 + r] : -1;
            int H[kGenK], E[kGenK];
#pragma unroll
            for (int r = 0; r < kGenK; ++r) { H[r] = 0; E[r] = 0; }
            int h_last = 0, f_out = 0, diag_in = 0;
            const bool first = pass == 0, last = pass == npass - 1;
            for (int s = 0; s < tlen + 31; ++s) {
                const int j = s - lane;
                int h_up = __shfl_up_sync(FULL, h_last, 1);
                int f = __shfl_up_sync(FULL, f_out, 1);
                if (lane == 0) {
                    if (first || j >= tlen) { h_up = 0; f = 0; }
                    else { h_up = __ldcg(bH + j); f = __ldcg(bF + j); }
                }
                int hd = diag_in;
                diag_in = h_up;
                const bool live = j >= 0 && j < tlen;
                const int8_t* srow = smat + (live ? ts[j] : S4G_PAD_CODE) * 32;
#pragma unroll
                for (int r = 0; r < kGenK; ++r) {
                    const int sc = ql[r] >= 0 ? (int)srow[ql[r]] : 0;
                    int h = __vimax3_s32_relu(hd + sc, E[r], f);
                    if (!live) h = 0;
                    hd = H[r];
                    H[r] = h;
                    const int hq = h - Q;
                    E[r] = __viaddmax_s32(E[r], -R, hq);
                    f = __viaddmax_s32(f, -R, hq);
                    if (!live) { E[r] = 0; f = 0; }
                    best = max(best, h);
                }
                h_last = H[kGenK - 1];
                f_out = f;
                if (lane == 31 && !last && live) { __stcg(bH + j, h_last); __stcg(bF + j, f_out); }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
        if (lane == 0) P.out[c] = best;
    }
}

// ------------------------------------------------------------------------------------------------------
// preparation kernels

__global__ void make_keys_kernel(ScoreParams P, int64_t n_pairs, unsigned long long* keys, uint32_t* vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int q = find_query(P.cand_off, P.nq, i);
    const uint32_t t = P.cand_ids[i] - P.id_base;
    const unsigned len = (unsigned)(P.db_off[t + 1] - P.db_off[t]);
    keys[i] = ((unsigned long long)q << 32) | (unsigned long long)(0xffffffffu - len);
    vals[i] = (uint32_t)i;
}

// tiles per query for the packed kernel (0 for long queries) and for the striped kernel (0 for short ones)
__global__ void count_tiles_kernel(ScoreParams P, int64_t* tiles, int64_t* long_cands, int packed_tile_pairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P.nq) return;
    int64_t t = 0, l = 0;
    if (i < P.nq) {
        const int q = P.q_order ? P.q_order[i] : i;                      // packed tiles: position i of the tile table is query q
        const int64_t n_c = P.cand_off[q + 1] - P.cand_off[q];
        const int qlen = (int)(P.q_off[q + 1] - P.q_off[q]);
        if (qlen <= 32 * kMaxK) t = ((n_c + 1) / 2 + packed_tile_pairs - 1) / packed_tile_pairs;
        const int64_t n_l = P.cand_off[i + 1] - P.cand_off[i];           // striped tiles stay in query order
        if ((int)(P.q_off[i + 1] - P.q_off[i]) > 32 * kMaxK) l = ((n_l + 1) / 2 + kTilePairs - 1) / kTilePairs;
    }
    tiles[i] = t;
    long_cands[i] = l;
}


// ---- end-cell mode: the kept hits as (query, target) work, grouped by query and sorted by target length ----

__global__ void hit_keys_kernel(ScoreParams P, const uint32_t* pair_q, int64_t n, unsigned long long* keys, uint32_t* vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = P.cand_ids[i] - P.id_base;
    const unsigned len = (unsigned)(P.db_off[t + 1] - P.db_off[t]);
    keys[i] = ((unsigned long long)pair_q[i] << 32) | (unsigned long long)(0xffffffffu - len);
    vals[i] = (uint32_t)i;
}

// first position of every query in the sorted key list (nq + 1 entries)
__global__ void hit_qstart_kernel(const unsigned long long* sorted_keys, int64_t n, int nq, int64_t* qstart) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > nq) return;
    const unsigned long long want = (unsigned long long)q << 32;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted_keys[mid] < want) lo = mid + 1; else hi = mid; }
    qstart[q] = lo;
}

}  // namespace

int s4g_sw_score_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* d_cand_ids,
                        const int64_t* d_cand_off, int64_t n_pairs, const int32_t* h_matrix, int gap_open,
                        int gap_extend, int32_t* d_out) {
    cudaStream_t st = ctx->stream;
    const int nq = q->n;
    // int8 matrix, rows = target letter (27, row 26 = pad), cols = query letter (32)
    int8_t h_mat8[(S4G_PAD_CODE + 1) * 32];
    int max_s = 0, min_s = 0;
    memset(h_mat8, 0, sizeof(h_mat8));
    for (int a = 0; a < S4G_NLET; ++a)
        for (int b = 0; b < S4G_NLET; ++b) {
            int v = h_matrix[b * S4G_NLET + a];   // S[query=b][target=a]
            if (v > 127 || v < -127) { s4g_set_error(ctx, "matrix entry %d outside int8 range", v); return S4G_ERR_ARG; }
            h_mat8[a * 32 + b] = (int8_t)v;
            if (v > max_s) max_s = v;
            if (v < min_s) min_s = v;
        }
    for (int b = 0; b < 32; ++b) h_mat8[S4G_PAD_CODE * 32 + b] = (int8_t)(min_s < -1 ? min_s : -1);
    if (gap_open + max_s >= 16000 || gap_extend >= 16000) { s4g_set_error(ctx, "gap penalties too large for the 16-bit kernel"); return S4G_ERR_ARG; }

    int8_t* d_mat8 = (int8_t*)s4g_scratch(ctx, SLOT_SW_MAT, sizeof(h_mat8));
    unsigned long long* d_keys = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_KEYS, sizeof(unsigned long long) * n_pairs);
    unsigned long long* d_keys2 = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_KEYS2, sizeof(unsigned long long) * n_pairs);
    uint32_t* d_vals = (uint32_t*)s4g_scratch(ctx, SLOT_SW_VALS, sizeof(uint32_t) * n_pairs);
    uint32_t* d_vals2 = (uint32_t*)s4g_scratch(ctx, SLOT_SW_VALS2, sizeof(uint32_t) * n_pairs);
    int64_t* d_tiles = (int64_t*)s4g_scratch(ctx, SLOT_SW_TILES, sizeof(int64_t) * 5 * (nq + 1));
    unsigned long long* d_counters = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_MISC, 64);
    uint32_t* d_ovf = (uint32_t*)s4g_scratch(ctx, SLOT_SW_OVF, sizeof(uint32_t) * 2 * n_pairs);
    if (!d_mat8 || !d_keys || !d_keys2 || !d_vals || !d_vals2 || !d_tiles || !d_counters || !d_ovf) return S4G_ERR_NOMEM;
    int64_t* d_tile_cnt = d_tiles, *d_tile_start = d_tiles + (nq + 1), *d_long_cnt = d_tiles + 2 * (nq + 1), *d_long_start = d_tiles + 3 * (nq + 1);

    const int gen_blocks = ctx->sm_count * 2;
    const int64_t bound_stride = 2 * ((int64_t)db->max_len + 64);
    int32_t* d_bound = (int32_t*)s4g_scratch(ctx, SLOT_SW_BOUND, sizeof(int32_t) * bound_stride * gen_blocks * kWarps);
    if (!d_bound) return S4G_ERR_NOMEM;

    S4G_CUDA(ctx, cudaMemcpyAsync(d_mat8, h_mat8, sizeof(h_mat8), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 64, st));

    ScoreParams P;
    P.db_codes = db->d_codes; P.db_off = db->d_off; P.id_base = db->id_base;
    P.q_codes = q->d_codes; P.q_off = q->d_off; P.nq = nq;
    P.cand_ids = d_cand_ids; P.cand_off = d_cand_off;
    P.sorted_idx = d_vals2; P.tile_start = d_tile_start; P.mat8 = d_mat8; P.out = d_out;
    { const char* e = getenv("S4G_TILE_ORDER"); P.q_order = (e && strcmp(e, "query") == 0) ? nullptr : q->d_len_order; }
    P.counters = d_counters; P.overflow = d_ovf; P.bound = d_bound; P.bound_stride = bound_stride;
    P.long_tile_start = d_long_start; P.strip_bound = nullptr; P.strip_cols = 0; P.pair_score = nullptr;
    P.gap_open = gap_open; P.gap_extend = gap_extend; P.ovf_limit = 32767 - max_s;

    // 1. sort candidates of each query by target length (longest first)
    {
        const int threads = 256;
        const int blocks = (int)((n_pairs + threads - 1) / threads);
        make_keys_kernel<<<blocks, threads, 0, st>>>(P, n_pairs, d_keys, d_vals);
        S4G_CHECK_LAUNCH(ctx);
        int qbits = 1;
        while ((1ll << qbits) < nq) ++qbits;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n_pairs, 0, 32 + qbits, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp_bytes);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n_pairs, 0, 32 + qbits, st));
        ctx->launches += 4;
    }
    // 2. tile table
    {
        count_tiles_kernel<<<(nq + 1 + 255) / 256, 256, 0, st>>>(P, d_tile_cnt, d_long_cnt, kStreamTilePairs);
        S4G_CHECK_LAUNCH(ctx);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_tile_cnt, d_tile_start, nq + 1, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp_bytes);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_tile_cnt, d_tile_start, nq + 1, st));
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_long_cnt, d_long_start, nq + 1, st));
        ctx->launches += 2;
    }
    // 3. packed kernel (persistent grid)
    {
        // two stream columns per step (S4G_SCORE=1col: the one-column loop, kept for comparison)
        const char* sc = getenv("S4G_SCORE");
        const bool two = !(sc && strcmp(sc, "1col") == 0);
        void (*score_kernel)(ScoreParams) = two ? sw_score_packed2_kernel : sw_score_packed_kernel;
        const size_t smem = (S4G_PAD_CODE + 1) * 8 * 32 * 4 + (S4G_PAD_CODE + 1) * 32 + kWarps * 4 * kRing * sizeof(unsigned short) +
                            kWarps * (two ? kDescRing2 : kDescRing) * sizeof(uint2);
        S4G_CUDA(ctx, cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_kernel, kWarps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        if (const char* e = getenv("S4G_SW_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // tools/corun_experiment.py
        S4G_CUDA(ctx, cudaEventRecord(ctx->ev_sw0, st));
        score_kernel<<<ctx->sm_count * per_sm, kWarps * 32, smem, st>>>(P);
        S4G_CHECK_LAUNCH(ctx);
    }
    // 4. long queries: striped kernel (stripes of 1024 rows, boundary rows through a per-CTA buffer)
    if (q->max_len > 32 * kMaxK) {
        const char* sv = getenv("S4G_STRIPED");
        const bool stream = !(sv && strcmp(sv, "pairs") == 0);
        const bool stream1 = sv && strcmp(sv, "stream1") == 0;
        void (*strip_kernel)(ScoreParams) = stream ? (stream1 ? sw_score_striped_stream_kernel : sw_score_striped_stream2_kernel) : sw_score_striped_kernel;
        const size_t smem = (S4G_PAD_CODE + 1) * 8 * 32 * 4 + (S4G_PAD_CODE + 1) * 32 + kWarps * (stream ? kStreamStripWarpBytes : kStripWarpBytes);
        S4G_CUDA(ctx, cudaFuncSetAttribute(strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, strip_kernel, kWarps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        const int grid = ctx->sm_count * per_sm;
        // boundary rows sized for the longest sequence of the shard (a titin-like query meets titin-like targets), within 16 GiB
        const size_t per_col = sizeof(unsigned) * (size_t)grid * kTilePairs * 2;
        int64_t strip_cols = std::max<int64_t>(kStripColsMin, ((int64_t)db->max_len + 63) / 64 * 64);
        strip_cols = std::min<int64_t>(strip_cols, (int64_t)(((size_t)16 << 30) / per_col) / 64 * 64);
        unsigned* d_strip = (unsigned*)s4g_scratch(ctx, SLOT_SW_STRIP, per_col * (size_t)strip_cols);
        P.strip_cols = (int32_t)strip_cols;
        if (!d_strip) return S4G_ERR_NOMEM;
        P.strip_bound = d_strip;
        strip_kernel<<<grid, kWarps * 32, smem, st>>>(P);
        S4G_CHECK_LAUNCH(ctx);
    }
    S4G_CUDA(ctx, cudaEventRecord(ctx->ev_sw1, st));      // s4g_last_sw_kernel_ms: packed + striped kernels
    ctx->sw_timed = true;
    // 5. exact 32-bit kernel for what the 16-bit kernels could not settle (scores near 32767, targets longer than the
    //    striped kernel's boundary buffer)
    {
        GenericWork W; W.list = d_ovf; W.n_list = d_counters + 1; W.long_idx = nullptr; W.n_long = 0;
        sw_score_generic_kernel<<<gen_blocks, kWarps * 32, 0, st>>>(P, W);
        S4G_CHECK_LAUNCH(ctx);
    }
    return S4G_OK;
}

// Stage 3, step 1 for queries of at most 32 * kMaxK residues: end cell (query row -> coords[4i+1], target column ->
// coords[4i+3]) of every kept hit, two hits of the same query per warp.  Hits of longer queries are left untouched
// (align.cu sweeps them with its 32-bit kernel).  d_flags[0] bit 0 is set when a score is not attained.
int s4g_sw_forward_ends_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n, const uint32_t* d_pair_q, const uint32_t* d_pair_t,
                               const int32_t* d_pair_score, const int8_t* d_mat8, int gap_open, int gap_extend, int32_t* d_coords,
                               unsigned long long* d_flags) {
    cudaStream_t st = ctx->stream;
    const int nq = q->n;
    unsigned long long* d_keys = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_KEYS, sizeof(unsigned long long) * n);
    unsigned long long* d_keys2 = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_KEYS2, sizeof(unsigned long long) * n);
    uint32_t* d_vals = (uint32_t*)s4g_scratch(ctx, SLOT_SW_VALS, sizeof(uint32_t) * n);
    uint32_t* d_vals2 = (uint32_t*)s4g_scratch(ctx, SLOT_SW_VALS2, sizeof(uint32_t) * n);
    int64_t* d_tiles = (int64_t*)s4g_scratch(ctx, SLOT_SW_TILES, sizeof(int64_t) * 5 * (nq + 1));
    unsigned long long* d_counters = (unsigned long long*)s4g_scratch(ctx, SLOT_SW_MISC, 64);
    if (!d_keys || !d_keys2 || !d_vals || !d_vals2 || !d_tiles || !d_counters) return S4G_ERR_NOMEM;
    int64_t* d_tile_cnt = d_tiles, *d_tile_start = d_tiles + (nq + 1), *d_long_cnt = d_tiles + 2 * (nq + 1), *d_qstart = d_tiles + 4 * (nq + 1);
    S4G_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 64, st));

    ScoreParams P;
    memset(&P, 0, sizeof(P));
    P.db_codes = db->d_codes; P.db_off = db->d_off; P.id_base = db->id_base;
    P.q_codes = q->d_codes; P.q_off = q->d_off; P.nq = nq;
    P.cand_ids = d_pair_t; P.cand_off = d_qstart;
    P.sorted_idx = d_vals2; P.tile_start = d_tile_start; P.mat8 = d_mat8; P.out = d_coords;
    { const char* e = getenv("S4G_TILE_ORDER"); P.q_order = (e && strcmp(e, "query") == 0) ? nullptr : q->d_len_order; }
    P.counters = d_counters; P.gap_open = gap_open; P.gap_extend = gap_extend; P.pair_score = d_pair_score;

    hit_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, d_pair_q, n, d_keys, d_vals);
    S4G_CHECK_LAUNCH(ctx);
    {
        int qbits = 1;
        while ((1ll << qbits) < nq) ++qbits;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 32 + qbits, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp_bytes);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 32 + qbits, st));
        ctx->launches += 4;
    }
    hit_qstart_kernel<<<(nq + 1 + 255) / 256, 256, 0, st>>>(d_keys2, n, nq, d_qstart);
    S4G_CHECK_LAUNCH(ctx);
    count_tiles_kernel<<<(nq + 1 + 255) / 256, 256, 0, st>>>(P, d_tile_cnt, d_long_cnt, kTrackTilePairs);
    S4G_CHECK_LAUNCH(ctx);
    {
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_tile_cnt, d_tile_start, nq + 1, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp_bytes);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_tile_cnt, d_tile_start, nq + 1, st));
        ctx->launches += 1;
    }
    const size_t smem = (S4G_PAD_CODE + 1) * 8 * 32 * 4 + (S4G_PAD_CODE + 1) * 32 + kWarps * 4 * kRing * sizeof(unsigned short);
    S4G_CUDA(ctx, cudaFuncSetAttribute(al_forward_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, al_forward_packed_kernel, kWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    al_forward_packed_kernel<<<ctx->sm_count * per_sm, kWarps * 32, smem, st>>>(P);
    S4G_CHECK_LAUNCH(ctx);
    // long queries: striped end-cell sweep (S4G_ENDS=sweep32 leaves them to align.cu's 32-bit kernel)
    const char* ev = getenv("S4G_ENDS");
    if (q->max_len > 32 * kMaxK && !(ev && strcmp(ev, "sweep32") == 0)) {
        int64_t* d_long_start = d_tiles + 3 * (nq + 1);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_long_cnt, d_long_start, nq + 1, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp_bytes);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_long_cnt, d_long_start, nq + 1, st));
        ctx->launches += 1;
        const size_t ssmem = (S4G_PAD_CODE + 1) * 8 * 32 * 4 + (S4G_PAD_CODE + 1) * 32 + kWarps * kStripWarpBytes;
        S4G_CUDA(ctx, cudaFuncSetAttribute(al_forward_striped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
        int sper = 0;
        S4G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sper, al_forward_striped_kernel, kWarps * 32, ssmem));
        if (sper < 1) sper = 1;
        const int grid = ctx->sm_count * sper;
        const size_t per_col = sizeof(unsigned) * (size_t)grid * kTilePairs * 2;
        int64_t strip_cols = std::max<int64_t>(kStripColsMin, ((int64_t)db->max_len + 63) / 64 * 64);
        strip_cols = std::min<int64_t>(strip_cols, (int64_t)(((size_t)16 << 30) / per_col) / 64 * 64);
        unsigned* d_strip = (unsigned*)s4g_scratch(ctx, SLOT_SW_STRIP, per_col * (size_t)strip_cols);
        if (!d_strip) return S4G_ERR_NOMEM;
        P.long_tile_start = d_long_start; P.strip_bound = d_strip; P.strip_cols = (int32_t)strip_cols;
        al_forward_striped_kernel<<<grid, kWarps * 32, ssmem, st>>>(P);
        S4G_CHECK_LAUNCH(ctx);
    }
    // fold the error flag into the caller's flag word
    S4G_CUDA(ctx, cudaMemcpyAsync(d_flags, d_counters + 3, 8, cudaMemcpyDeviceToDevice, st));
    return S4G_OK;
}

int s4g_sw_long_query_rows() { return 32 * kMaxK; }
