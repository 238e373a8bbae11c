// Stage 1: k-mer candidate prefilter (sm_100a).
//
// Replaces searchDatabase / threadSearchDatabase (sift4g/src/database_search.cpp:66-253), the query k-mer
// index (sift4g/src/hash.cpp:21-90) and longestIncreasingSubsequence (database_search.cpp:255-280):
//   for every database sequence d and every query q sharing at least one k-mer with it,
//       score(q,d) = LIS_strict(query positions of the shared k-mers, in database order) / (float) len(d)
//   and every query keeps its max_candidates best sequences.  Ties at the cut-off are broken by ascending
//   sequence id (a deterministic member of the reference's thread-count dependent tie family).
//
// Device data structures (all L2 resident for the configurations of BASELINE.json):
//   bitrank[2^(5k)/32] = {32 presence bits, number of present k-mers before this word}   (one 8 B load)
//   bucket_start[D+1]  = first entry of each present k-mer in `hits`
//   hits[]             = (query << 32 | position) sorted by (k-mer, query, position)        (hash.cpp:80-84)
// The database streams from HBM exactly once per call: one warp per sequence, 32 k-mer positions per step.
// Hits of a sequence are gathered in the reference's emission order into a per-warp shared-memory buffer,
// bitonic-sorted by (query, emission order), and each query's run is reduced with an in-place patience LIS.
// Sequences with more hits than the buffer holds are deferred to a global-memory path (radix sort + one
// thread per run).  Scored pairs that beat the query's current cut-off are appended to that query's
// candidate buffer; buffers are compacted (segmented sort, keep the best N, raise the cut-off) whenever
// they could overflow during the next chunk of sequences.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <type_traits>
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kSeqBatch = 4;           // sequences claimed per atomic
constexpr int kGCap = 1024;             // surviving hits per sequence handled in the per-warp global scratch
constexpr unsigned long long kNoThr = ~0ull;

struct PfParams {
    const uint8_t* db_codes;
    int64_t local_lo, local_hi;       // bytes of db_codes in this GPU's HBM (an NVLink-striped view: the rest is read from peers)
    const int64_t* db_off;
    const uint32_t* order;            // scan order: local sequence indices by ascending length
    uint32_t id_base;
    int64_t seq_begin, seq_end;
    int k;
    uint32_t mask;
    const uint2* bitrank;
    const uint32_t* bucket_start;
    const unsigned long long* hits;
    const unsigned long long* entry;  // per index k-mer: the queries of its <= 3 hits (bits 63..62 = count), else (bucket size << 32 | first entry in `hits`)
    const int64_t* q_hit_start;       // first index entry of every query (nq + 1): equal neighbours = a query without k-mers
    int nq;
    unsigned long long* thr;
    uint32_t* count;
    unsigned long long* cand;
    uint32_t cap;
    unsigned long long* counters;     // [0] sequence cursor [1] deferred sequences [2] pool cursor [3] error flags
    uint32_t* def_seq;                // deferred: local sequence index (in the shard, not in the scan order)
    unsigned long long* def_off;      // deferred: offset into the pool
    uint32_t max_deferred;
    unsigned long long* pool_keys;    // (slot << 48 | q << 28 | order)
    uint32_t* pool_vals;              // query position
    unsigned long long pool_cap;
    unsigned long long* gbuf;         // per-warp global scratch (kGCap entries) for sequences with many surviving hits
    int trace;                        // S4G_TRACE: counters[5] / [6] count the flagged sequences re-walked / replayed from the rank queue
    int no_replay;                    // S4G_PF_REPLAY=0: every flagged sequence is re-walked (the form before r04, for comparison)
};

__device__ __forceinline__ unsigned long long cand_key(float score, uint32_t id) {
    return ((unsigned long long)(~__float_as_uint(score)) << 32) | id;   // ascending = (score desc, id asc)
}

__device__ __forceinline__ void emit(const PfParams& P, uint32_t q, int lis, int len, uint32_t id) {
    const float score = __fdiv_rn((float)lis, (float)len);               // database_search.cpp:228-229
    const unsigned long long key = cand_key(score, id);
    if (key < __ldcg(P.thr + q)) {
        const uint32_t slot = atomicAdd(P.count + q, 1u);
        if (slot < P.cap) P.cand[(size_t)q * P.cap + slot] = key;
        else atomicOr(P.counters + 3, 1ull);
    }
}

// k-mer at position j (needs j + k <= len)
__device__ __forceinline__ uint32_t kmer_at(const uint8_t* s, int j, int k) {
    uint32_t v = 0;
    for (int i = 0; i < k; ++i) v = (v << 5) | s[j + i];
    return v;
}

__device__ __forceinline__ void lookup(const PfParams& P, uint32_t kmer, uint32_t& b, uint32_t& c) {
    const uint2 br = __ldg(P.bitrank + (kmer >> 5));
    const uint32_t bit = kmer & 31u;
    c = 0; b = 0;
    if ((br.x >> bit) & 1u) {
        const uint32_t r = br.y + __popc(br.x & ((1u << bit) - 1u));
        b = __ldg(P.bucket_start + r);
        c = __ldg(P.bucket_start + r + 1) - b;
    }
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// in-place patience LIS over the positions of entries [a, a+n): strictly increasing
__device__ int lis_inplace(unsigned long long* buf, int a, int n) {
    uint32_t* tails = reinterpret_cast<uint32_t*>(buf + a);
    int len = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t x = (uint32_t)(buf[a + i] & 0x3fffffu);
        if (len == 0 || tails[len - 1] < x) { tails[len++] = x; continue; }     // collinear hits: O(1)
        int lo = 0, hi = len - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (tails[mid] < x) lo = mid + 1; else hi = mid; }
        tails[lo] = x;
    }
    return len;
}

// k-mers of one scan step of a warp: 128 positions starting at `base`, 4 consecutive positions per lane.  probe[i] says
// whether position j0 + i takes part at all: inside the sequence and not a consecutive duplicate of the k-mer in front of it
// (database_search.cpp:212-214).
template <bool kStaged = false>
__device__ __forceinline__ void step_kmers(const PfParams& P, const uint8_t* seq, int npos, int base, int lane, uint32_t& carry,
                                           uint32_t (&km)[4], bool (&probe)[4]) {
    const unsigned FULL = 0xffffffffu;
    const int k = P.k;
    const int j0 = base + 4 * lane;
    // 8 residue bytes of this lane through 3 aligned 32-bit loads (bytes beyond the sequence only reach k-mers at
    // positions >= npos, which are discarded; the database buffer has S4G_DB_TAIL_PAD readable bytes after its end)
    const uintptr_t addr = reinterpret_cast<uintptr_t>(seq + j0);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(addr & 3u) * 8u;
    uint32_t lo = 0, hi = 0;
    if (j0 < npos) {
        uint32_t w0, w1, w2;
        if (kStaged) {                                                     // `seq` points into the warp's staging slot: LDS, not generic loads
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(w);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(sa));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(sa));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(sa));
        }
        else { w0 = __ldg(w); w1 = __ldg(w + 1); w2 = __ldg(w + 2); }
        lo = __funnelshift_r(w0, w1, sh);
        hi = __funnelshift_r(w1, w2, sh);
    }
    uint32_t by[8];
#pragma unroll
    for (int x = 0; x < 4; ++x) { by[x] = (lo >> (8 * x)) & 0x1fu; by[x + 4] = (hi >> (8 * x)) & 0x1fu; }
    {
        uint32_t v = (((by[0] << 5) | by[1]) << 5) | by[2];
        if (k >= 4) v = (v << 5) | by[3];
        if (k >= 5) v = (v << 5) | by[4];
        km[0] = v;
#pragma unroll
        for (int i = 1; i < 4; ++i) {
            const uint32_t nb = k == 5 ? by[i + 4] : (k == 4 ? by[i + 3] : by[i + 2]);
            km[i] = ((km[i - 1] << 5) | nb) & P.mask;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) if (j0 + i >= npos) km[i] = 0xfffffffeu;
    uint32_t prev0 = __shfl_up_sync(FULL, km[3], 1);
    if (lane == 0) prev0 = carry;
    carry = __shfl_sync(FULL, km[3], 31);
    probe[0] = j0 < npos && !(j0 > 0 && km[0] == prev0);
#pragma unroll
    for (int i = 1; i < 4; ++i) probe[i] = (j0 + i < npos) && km[i] != km[i - 1];
}

// One scan step of a warp (the re-walk and the deferred path): the lane's buckets -- first entry hb[i] and size hc[i] for
// each of its positions (0 when the k-mer is absent from the index, is a consecutive duplicate or lies beyond the sequence).
__device__ __forceinline__ void scan_step(const PfParams& P, const uint8_t* seq, int npos, int base, int lane, uint32_t& carry,
                                          uint32_t (&hb)[4], uint32_t (&hc)[4]) {
    uint32_t km[4];
    bool probe[4];
    step_kmers(P, seq, npos, base, lane, carry, km, probe);
    uint2 br[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) br[i] = probe[i] ? __ldg(P.bitrank + (km[i] >> 5)) : make_uint2(0u, 0u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t bit = km[i] & 31u;
        hb[i] = 0; hc[i] = 0;
        if (probe[i] && ((br[i].x >> bit) & 1u)) {
            const uint32_t r = br[i].y + __popc(br[i].x & ((1u << bit) - 1u));
            hb[i] = __ldg(P.bucket_start + r);
            hc[i] = __ldg(P.bucket_start + r + 1) - hb[i];
        }
    }
}

// Publish the lane's buckets of one scan step in the warp's shared-memory tables (start offset of every one of the
// 128 positions inside the step's hit list + first index entry) so that the step's hits can be enumerated
// load-balanced: hit x of the step belongs to the last position p with off_s[p] <= x.
__device__ __forceinline__ void publish_step(uint32_t* off_s, uint32_t* hb_s, int lane, uint32_t excl, const uint32_t (&hb)[4],
                                             const uint32_t (&hc)[4]) {
    uint32_t o = excl;
#pragma unroll
    for (int i = 0; i < 4; ++i) { off_s[4 * lane + i] = o; hb_s[4 * lane + i] = hb[i]; o += hc[i]; }
    __syncwarp();
}

__device__ __forceinline__ uint32_t step_hit_index(const uint32_t* off_s, const uint32_t* hb_s, uint32_t x) {
    int p = 0;
#pragma unroll
    for (int b = 64; b > 0; b >>= 1) if (off_s[p + b] <= x) p += b;
    return hb_s[p] + (x - off_s[p]);
}

// Lane-owned enumeration: hit j of the lane's own (up to four) buckets, in emission order.  No table, no search: used
// whenever the buckets of a step are spread evenly enough over the lanes (see lane_owned()).
__device__ __forceinline__ uint32_t lane_hit_index(const uint32_t (&hb)[4], const uint32_t (&hc)[4], uint32_t j) {
    const uint32_t p1 = hc[0], p2 = p1 + hc[1], p3 = p2 + hc[2];
    uint32_t base = hb[0], pre = 0;
    if (j >= p1) { base = hb[1]; pre = p1; }
    if (j >= p2) { base = hb[2]; pre = p2; }
    if (j >= p3) { base = hb[3]; pre = p3; }
    return base + (j - pre);
}

// The warp walks max-over-lanes(c) rounds when every lane enumerates its own hits; the load-balanced enumeration walks
// total/32 rounds but pays a 7-level table search per hit.  Lane-owned only for sparse steps (small query batches: a
// step holds a few dozen hits in buckets of one or two): measured 1 % faster there, but 30 % slower on dense steps
// (20 000 queries, ~435 hits per step), where every lane reading its own bucket turns one coalesced request into 32.
__device__ __forceinline__ bool lane_owned(uint32_t cmax, uint32_t total) { return total <= 64u && cmax <= 6u; }

// Exact per-sequence hit counters: one 16-bit counter per query of the batch (two per 32-bit word), private to the warp.
__device__ __forceinline__ uint32_t count_of(const uint32_t* cnt, uint32_t q) { return reinterpret_cast<const unsigned short*>(cnt)[q]; }

// Can a (query, sequence) pair still beat the query's cut-off?  LIS <= number of hits, so a pair whose hit COUNT cannot reach
// cut-off x length is no candidate.  The cut-off comes from the CTA's shared table: upper 16 bits of the float score, i.e.
// rounded down -- conservative; 0 while the query has no cut-off yet.
__device__ __forceinline__ bool may_pass(const unsigned short* qthr, uint32_t count, uint32_t q, float flen) {
    const float th = __uint_as_float((uint32_t)qthr[q] << 16);
    return !((float)count < th * flen);
}

constexpr uint32_t kRankRing = 256, kBucketRing = 64;      // entries of the two per-warp queues of pass A (1 KB + 512 B)
constexpr uint32_t kLaneBucket = 24;      // index buckets up to this size are walked by one lane, larger ones by the warp

// The scan kernel.  Per sequence (one warp):
//   pass A  walks the k-mer positions (4 per lane and step) and COUNTS the hits of every query in the warp's exact counters.
//           Nothing is stored: a shared-memory atomic returns the previous count, and the one hit that takes a query's count
//           across  cut-off x length  raises a flag.  The presence/rank probe is a random 32-byte sector access, of which an
//           SM completes one per clock (tools/gather_microbench.cu: 290 G/s whole GPU); k-mers of the index with up to three
//           hits carry the hits' queries in their bucket entry (one access instead of two).  A Bloom filter in shared memory in front of the
//           probe was built and measured: it halves the probes and makes the scan slower (the kernel is bound by issued
//           instructions, not by the L1 wavefront rate -- profiles/r02_prefilter.md).
//   no flag -> the sequence holds no candidate for any query: clear the counters, next sequence (practically every
//           sequence once the cut-offs have settled).
//   pass B  re-walks the sequence and keeps the hits of the queries whose exact count reaches their cut-off; a query with
//           exactly one hit has LIS = 1 and is emitted straight away, the other survivors are bitonic-sorted by
//           (query, emission order) and each query's run is reduced by the in-place LIS.
// Shared memory: per warp  scap x 8 B (sort buffer; the step tables of pass B alias it) + cnt_words x 4 B counters;
// per CTA  the 2-byte cut-off table.
// Residue staging (kStage): every warp keeps a ring of kStageSlots segments of up to kStageSeg k-mer positions in shared memory,
// filled by TMA bulk copies (cp.async.bulk, one mbarrier per slot) that run up to kStageSlots - 1 segments ahead of the walk --
// across the sequences of a batch and into the next batch.  The walk then reads residues with LDS instead of three LDG per lane
// and step, and the copies hide the latency of the residue stream: ~0.35 us from the local HBM, ~1.7 us from a peer's HBM over
// NVLink (tools/peer_read_microbench.cu), where prefetch.global.L2 is no option.  Segment boundaries are multiples of the
// 128-position step, so the walk is step for step the unstaged one.
#ifndef S4G_STAGE_SEG
#define S4G_STAGE_SEG 256
#endif
#ifndef S4G_STAGE_SLOTS
#define S4G_STAGE_SLOTS 2
#endif
constexpr int kStageSeg = S4G_STAGE_SEG;             // k-mer positions per segment (a multiple of the 128-position step)
constexpr int kStageSlot = 16 + kStageSeg + 16;      // bytes: alignment head + residues + k-1 tail and the lanes' 8-byte windows
constexpr int kStageSlots = S4G_STAGE_SLOTS;         // every KB of shared memory is a KB less L1 for the index probes: keep it small
constexpr int kStageWarpBytes = kStageSlots * (kStageSlot + 16 + 8);     // + descriptor + mbarrier per slot

struct StageDesc { long long a; uint32_t s; int32_t len; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool kStage>
__device__ __forceinline__ void pf_scan_body(const PfParams& P, int scap, int cnt_words) {
    extern __shared__ unsigned long long sbuf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned long long* buf = sbuf + (size_t)warp * scap;
    uint32_t* cnt_all = reinterpret_cast<uint32_t*>(sbuf + (size_t)nwarps * scap);
    uint32_t* cnt = cnt_all + (size_t)warp * cnt_words;
    unsigned short* qthr = reinterpret_cast<unsigned short*>(cnt_all + (size_t)nwarps * cnt_words);
    // staging area behind the cut-off table (16-byte aligned): per warp kStageSlots x (slot | descriptor) + mbarriers
    unsigned char* stage_base = reinterpret_cast<unsigned char*>(sbuf) + (((size_t)nwarps * scap * 8 + (size_t)nwarps * cnt_words * 4 + (((size_t)P.nq + 7) & ~(size_t)7) * 2 + 15) & ~(size_t)15);
    unsigned char* st_slots = stage_base + (size_t)warp * kStageWarpBytes;
    StageDesc* st_desc = reinterpret_cast<StageDesc*>(st_slots + kStageSlots * kStageSlot);
    unsigned long long* st_bar = reinterpret_cast<unsigned long long*>(st_desc + kStageSlots);
    if (kStage && lane == 0) {
        for (int i = 0; i < kStageSlots; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(st_bar + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t* off_s = reinterpret_cast<uint32_t*>(buf);      // pass B only (step tables); its survivors go to the global scratch
    uint32_t* hb_s = off_s + 128;
    uint32_t* ring_r = reinterpret_cast<uint32_t*>(buf);     // pass A only: queue of index ranks (31 left over + 128 of a step)
    unsigned long long* ring_m = buf + kRankRing / 2;        // pass A only: queue of buckets (31 left over + 32 of a take)
    unsigned short* ring_p = reinterpret_cast<unsigned short*>(ring_m + kBucketRing);   // sequence position of every queued rank (the replay of pass B)
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr int kHitUnroll = 2;
    const unsigned FULL = 0xffffffffu;
    const int k = P.k;
    for (int i = threadIdx.x; i < nwarps * cnt_words; i += blockDim.x) cnt_all[i] = 0;
    for (int q = threadIdx.x; q < P.nq; q += blockDim.x) qthr[q] = (unsigned short)((~(uint32_t)(__ldcg(P.thr + q) >> 32)) >> 16);   // no cut-off yet -> 0
    __syncthreads();
    // smallest cut-off score of the batch (0 while some query has none)
    float thr_min;
    {
        uint32_t m = 0xffffu;
        for (int q = lane; q < P.nq; q += 32) if (P.q_hit_start[q + 1] > P.q_hit_start[q]) m = min(m, (uint32_t)qthr[q]);   // queries shorter than k never get one
        thr_min = __uint_as_float(__reduce_min_sync(FULL, m) << 16);
    }
    // Sequences are claimed in batches of kSeqBatch positions of the scan order (lane bi holds sequence bi of the batch: index,
    // begin, end) with an L2 prefetch of their residues -- resident ones only: prefetch.global.L2 on a peer's page costs ~60x
    // the load it is meant to hide (tools/peer_read_microbench.cu: 121 ms vs 1.9 ms for the same walk).
    long long my_a = 0, my_e = 0;
    uint32_t my_idx = 0;
    int b_i = 0, b_n = 0;                                   // next sequence of the batch, sequences in the batch
    auto next_sequence = [&](long long& s, int64_t& a, int& len) -> bool {
        while (true) {
            if (b_i >= b_n) {
                long long s0 = 0;
                if (lane == 0) s0 = (long long)atomicAdd(P.counters + 0, (unsigned long long)kSeqBatch);
                s0 = __shfl_sync(FULL, s0, 0) + P.seq_begin;
                if (s0 >= P.seq_end) return false;
                my_a = 0; my_e = 0; my_idx = 0;
                if (lane < kSeqBatch && s0 + lane < P.seq_end) {
                    my_idx = __ldg(P.order + s0 + lane);
                    my_a = P.db_off[my_idx];
                    my_e = P.db_off[my_idx + 1];
                }
#pragma unroll
                for (int bi = 0; bi < kSeqBatch; ++bi) {
                    const long long b0 = __shfl_sync(FULL, my_a, bi), b1 = __shfl_sync(FULL, my_e, bi);
                    if (b0 >= P.local_lo && b1 <= P.local_hi)
                        for (long long x = b0 + 128ll * lane; x < b1; x += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.db_codes + x));
                }
                b_i = 0;
                b_n = (int)min((long long)kSeqBatch, P.seq_end - s0);
            }
            s = (long long)__shfl_sync(FULL, my_idx, b_i);
            a = __shfl_sync(FULL, my_a, b_i);
            len = (int)(__shfl_sync(FULL, my_e, b_i) - a);
            ++b_i;
            if (len >= k) return true;
        }
    };
    // staging: the producer walks the same sequence stream ahead of the consumer, one segment per slot
    uint32_t st_prod = 0, st_cons = 0;                      // segments issued / consumed
    long long p_s = 0; int64_t p_a = 0; int p_len = 0, p_pos = -1;   // sequence being cut into segments (p_pos < 0: fetch the next one)
    bool p_end = false;
    auto produce_ahead = [&]() {
        while (st_prod - st_cons < (uint32_t)kStageSlots && !p_end) {
            if (p_pos < 0) {
                if (!next_sequence(p_s, p_a, p_len)) { p_end = true; break; }
                p_pos = 0;
            }
            const int p_npos = p_len - k + 1;
            const int seg_n = min(kStageSeg, p_npos - p_pos);
            const uint32_t slot = st_prod % kStageSlots;
            if (lane == 0) {
                const uint8_t* src = P.db_codes + p_a + p_pos;
                const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
                const uint32_t bytes = (head + (uint32_t)seg_n + 12u + 15u) & ~15u;          // <= kStageSlot; the buffer's tail pad keeps it readable
                st_desc[slot].a = p_a; st_desc[slot].s = (uint32_t)p_s; st_desc[slot].len = p_len;
                const uint32_t bar = smem_u32(st_bar + slot);
                // (no proxy fence: the slot was only READ through the generic proxy, and those loads have delivered their data --
                // the k-mers of the segment are computed -- before the warp gets here)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(st_slots + slot * kStageSlot)), "l"(src - head), "r"(bytes), "r"(bar) : "memory");
            }
            ++st_prod;
            p_pos += kStageSeg;
            if (p_pos >= p_npos) p_pos = -1;
        }
        __syncwarp();
    };
    auto stage_wait = [&](uint32_t seg) {
        const uint32_t bar = smem_u32(st_bar + seg % kStageSlots), parity = (seg / kStageSlots) & 1u;
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    };
    __syncthreads();                                        // mbarriers initialised
    {
        while (true) {
            long long s; int64_t a; int len;
            if (kStage) {
                produce_ahead();
                if (st_cons == st_prod) break;
                const StageDesc d = st_desc[st_cons % kStageSlots];
                s = d.s; a = d.a; len = d.len;
            } else {
                if (!next_sequence(s, a, len)) break;
            }
            const uint8_t* seq = P.db_codes + a;
            const uint8_t* seg_seq = seq;                    // staged: the current segment's bytes, addressed by sequence position
            const int npos = len - k + 1;
            const float flen = (float)len * 0.99999f;      // margin >> float rounding: dropping stays exact
            // ---- pass A: count; `crossed` = one of this lane's hits took a query across its cut-off.
            // need_min: no query can reach its cut-off in this sequence with fewer hits (from the smallest cut-off of the batch;
            // 1 while some query has none), so the exact per-query test only runs for the rare counts that get that far.
            const uint32_t need_min = max(1u, (uint32_t)ceilf(thr_min * flen));
            uint32_t nh = 0, carry = 0xffffffffu;            // nh: hits counted by this lane
            bool crossed = false;
            bool wide_bucket = false;                        // a bucket beyond 63 hits: the replay's order field cannot number its hits
            auto count_hit = [&](uint32_t q) {
                const uint32_t old = atomicAdd(cnt + (q >> 1), (q & 1u) ? 0x10000u : 1u);
                const uint32_t oc = (q & 1u) ? (old >> 16) : (old & 0xffffu);
                if (oc + 1u >= need_min) {
                    // count oc -> oc + 1 crosses the query's own threshold (a query without cut-off crosses with its first hit)
                    const float need = fmaxf(__uint_as_float((uint32_t)qthr[q] << 16) * flen, 0.5f);
                    crossed |= ((float)(oc + 1u) >= need) && ((float)oc < need);
                }
            };
            // Only ~1 position in 6 of a 1 000-query batch meets a k-mer of the index, so the lanes do not chase their own
            // positions any further than the presence bit: the index ranks of the k-mers found are queued in the warp's ring and
            // taken out 32 at a time -- entry load and counting run with full warps.  An entry either holds the queries of the
            // k-mer's (up to three) hits, counted at once, or names a bucket; buckets are queued once more and walked 32 at a
            // time, one lane each (the buckets of a batch of a few thousand queries are small), the rare large ones by the warp.
            uint32_t r_head = 0, r_n = 0, m_head = 0, m_n = 0;        // both rings: first item, items queued (warp-uniform)
            auto flush_buckets = [&](uint32_t n) {                  // n <= 32
                unsigned long long e = 0ull;
                if ((uint32_t)lane < n) e = ring_m[(m_head + lane) & (kBucketRing - 1u)];
                m_head += n; m_n -= n;
                const uint32_t st = (uint32_t)e, c = (uint32_t)(e >> 32);
                nh += c;
                const bool big = c > kLaneBucket;
                if (!big) {
                    for (uint32_t j = 0; j < c; j += 2) {
                        const unsigned long long h0 = __ldg(P.hits + st + j);
                        const unsigned long long h1 = j + 1 < c ? __ldg(P.hits + st + j + 1) : 0ull;
                        count_hit((uint32_t)(h0 >> 32));
                        if (j + 1 < c) count_hit((uint32_t)(h1 >> 32));
                    }
                }
                unsigned m = __ballot_sync(FULL, big);
                while (m) {                                          // a frequent k-mer: the whole warp walks its bucket
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t bs = __shfl_sync(FULL, st, src), bc = __shfl_sync(FULL, c, src);
                    for (uint32_t x = lane; x < bc; x += 32) count_hit((uint32_t)(__ldg(P.hits + bs + x) >> 32));
                }
                __syncwarp();
            };
            auto take = [&](uint32_t n) {                           // n <= 32
                unsigned long long e = 0ull;
                if ((uint32_t)lane < n) e = __ldg(P.entry + ring_r[(r_head + lane) & (kRankRing - 1u)]);
                r_head += n; r_n -= n;
                const uint32_t inl = (uint32_t)(e >> 62);            // hits whose queries ride in the entry
                if (inl >= 1u) count_hit((uint32_t)e & 0xfffffu);
                if (inl >= 2u) count_hit((uint32_t)(e >> 20) & 0xfffffu);
                if (inl == 3u) count_hit((uint32_t)(e >> 40) & 0xfffffu);
                nh += inl;
                const bool bucket = e != 0ull && inl == 0u;
                wide_bucket |= bucket && (uint32_t)(e >> 32) > 63u;
                const unsigned m = __ballot_sync(FULL, bucket);
                if (m) {
                    if (bucket) ring_m[(m_head + m_n + __popc(m & lt_mask)) & (kBucketRing - 1u)] = e;
                    m_n += __popc(m);
                    __syncwarp();
                    if (m_n >= 32u) flush_buckets(32u);
                }
            };
            for (int base = 0;; base += 128) {
                const bool drain = base >= npos;
                if (!drain) {
                    uint32_t km[4];
                    bool probe[4];
                    if (kStage) {
                        if ((base % kStageSeg) == 0) {               // next segment of this sequence: release the old slot, refill ahead
                            if (base > 0) { ++st_cons; produce_ahead(); }
                            stage_wait(st_cons);
                            seg_seq = st_slots + (st_cons % kStageSlots) * kStageSlot + (reinterpret_cast<uintptr_t>(seq + base) & 15u) - base;
                        }
                        step_kmers<true>(P, seg_seq, npos, base, lane, carry, km, probe);
                    } else {
                        step_kmers(P, seq, npos, base, lane, carry, km, probe);
                    }
                    uint2 br[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) br[i] = probe[i] ? __ldg(P.bitrank + (km[i] >> 5)) : make_uint2(0u, 0u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t bit = km[i] & 31u;
                        const bool has = (br[i].x >> bit) & 1u;      // br is zero where the position does not take part
                        const unsigned m = __ballot_sync(FULL, has);
                        if (has) {
                            const uint32_t slot = (r_head + r_n + __popc(m & lt_mask)) & (kRankRing - 1u);
                            ring_r[slot] = br[i].y + __popc(br[i].x & ((1u << bit) - 1u));
                            if (!P.no_replay) ring_p[slot] = (unsigned short)(base + 4 * lane + i);
                        }
                        r_n += __popc(m);
                    }
                    __syncwarp();
                }
                while (r_n >= 32u || (drain && r_n > 0u)) take(r_n < 32u ? r_n : 32u);
                if (drain) {
                    if (m_n > 0u) flush_buckets(m_n);
                    break;
                }
            }
            if (kStage) { __syncwarp(); ++st_cons; }          // the sequence's last segment is consumed
            const uint32_t T = __reduce_add_sync(FULL, nh);
            if (T == 0) continue;
            const bool any_crossed = __any_sync(FULL, crossed);
            const uint32_t id = P.id_base + (uint32_t)s;
            uint32_t nsurv = 0;
            bool defer = false;
            unsigned long long* wb = buf;        // where the survivors are: shared buffer or global scratch
            if (T >= 65536u) {
                defer = true;                                 // a 16-bit counter may have wrapped (and the order field holds 22 bits): no filtering
            } else if (any_crossed) {
                // ---- pass B: re-walk, keep the survivors only (any order: the sort key carries the emission order); they go to
                // the warp's global scratch, which holds what a strong homolog of a long query produces
                __syncwarp();
                wb = P.gbuf + (size_t)(blockIdx.x * nwarps + warp) * kGCap;
                // Queries whose counter holds exactly one hit have LIS = 1 and are emitted straight from the walk -- but only
                // when the other survivors are sure to fit the scratch (all hits do, or all hits in counters >= 2), because a
                // sequence that overflows it is handed to the deferred path as a whole and must not have emitted anything yet.
                bool direct = T <= (uint32_t)kGCap;
                if (!direct) {
                    uint32_t m2 = 0;
                    for (int i = lane; i < cnt_words; i += 32) {
                        const uint32_t w = cnt[i], c0 = w & 0xffffu, c1 = w >> 16;
                        m2 += (c0 >= 2u ? c0 : 0u) + (c1 >= 2u ? c1 : 0u);
                    }
                    direct = __reduce_add_sync(FULL, m2) <= (uint32_t)kGCap;
                }
                // The k-mers found in pass A are still in the rank queue when the sequence queued at most kRankRing of them (r_head counts
                // from 0 per sequence, nothing was overwritten): the survivors are then collected by REPLAYING the queue -- entry load,
                // counter test, and the query positions of the few hits that stay (counters >= 2) -- instead of walking the residues and
                // probing the index a second time.  Short sequences reach a cut-off with one or two hits (1 / len, 2 / len), so on a
                // length-sorted scan the first third of the database is flagged almost sequence by sequence.  Order field of a hit =
                // (position << 6 | index in its bucket): the emission order of the walk.
                const uint32_t F = r_head;
                const bool replay = !P.no_replay && F <= kRankRing && len < 65536 && !__any_sync(FULL, wide_bucket);
                if (P.trace && lane == 0) atomicAdd(P.counters + (replay ? 6 : 5), 1ull);
                for (uint32_t r0 = 0; replay && r0 < F && !defer; r0 += 32) {
                    const uint32_t idx = r0 + lane;
                    unsigned long long e = 0ull;
                    uint32_t rank = 0, pos = 0;
                    if (idx < F) { rank = ring_r[idx]; pos = ring_p[idx]; e = __ldg(P.entry + rank); }
                    const uint32_t inl = (uint32_t)(e >> 62);
                    const uint32_t c = inl ? inl : (uint32_t)(e >> 32), st = (uint32_t)e;       // hits of this lane's k-mer (a bucket: first entry st)
                    const uint32_t cmax = __reduce_max_sync(FULL, c);
                    for (uint32_t j = 0; j < cmax && !defer; ++j) {
                        unsigned long long ee = 0;
                        bool keep = false;
                        if (j < c) {
                            unsigned long long h = 0ull;
                            uint32_t q;
                            if (inl) q = (uint32_t)(e >> (20u * j)) & 0xfffffu;
                            else { h = __ldg(P.hits + st + j); q = (uint32_t)(h >> 32); }
                            const uint32_t cq = count_of(cnt, q);
                            keep = may_pass(qthr, cq, q, flen);
                            if (direct && keep && cq == 1) { emit(P, q, 1, len, id); keep = false; }
                            if (keep) {
                                if (inl) h = __ldg(P.hits + __ldg(P.bucket_start + rank) + j);
                                ee = ((unsigned long long)q << 44) | ((unsigned long long)((pos << 6) | j) << 22) | (h & 0x3fffffu);
                            }
                        }
                        const uint32_t bal = __ballot_sync(FULL, keep);
                        if (nsurv + __popc(bal) > (uint32_t)kGCap) { defer = true; continue; }
                        if (keep) wb[nsurv + __popc(bal & lt_mask)] = ee;
                        nsurv += __popc(bal);
                    }
                }
                uint32_t ordbase = 0;
                carry = 0xffffffffu;
                for (int base = 0; !replay && base < npos && !defer; base += 128) {
                    uint32_t hb[4], hc[4];
                    scan_step(P, seq, npos, base, lane, carry, hb, hc);
                    const uint32_t c = hc[0] + hc[1] + hc[2] + hc[3];
                    uint32_t total;
                    const uint32_t excl = warp_excl_scan(c, lane, total);
                    if (total == 0) continue;
                    const uint32_t cmax = __reduce_max_sync(FULL, c);
                    if (lane_owned(cmax, total)) {
                        for (uint32_t j0 = 0; j0 < cmax && !defer; j0 += kHitUnroll) {
                            unsigned long long h[kHitUnroll];
#pragma unroll
                            for (int u = 0; u < kHitUnroll; ++u) h[u] = j0 + u < c ? __ldg(P.hits + lane_hit_index(hb, hc, j0 + u)) : 0ull;
#pragma unroll
                            for (int u = 0; u < kHitUnroll; ++u) {
                                if (defer || j0 + u >= cmax) continue;                  // warp-uniform
                                unsigned long long e = 0;
                                bool keep = false;
                                if (j0 + u < c) {
                                    const uint32_t q = (uint32_t)(h[u] >> 32);
                                    const uint32_t cq = count_of(cnt, q);
                                    keep = may_pass(qthr, cq, q, flen);
                                    if (direct && keep && cq == 1) { emit(P, q, 1, len, id); keep = false; }
                                    e = ((unsigned long long)q << 44) | ((unsigned long long)(ordbase + excl + j0 + u) << 22) | (h[u] & 0x3fffffu);
                                }
                                const uint32_t bal = __ballot_sync(FULL, keep);
                                if (nsurv + __popc(bal) > (uint32_t)kGCap) { defer = true; continue; }
                                if (keep) wb[nsurv + __popc(bal & ((1u << lane) - 1u))] = e;
                                nsurv += __popc(bal);
                            }
                        }
                        __syncwarp();
                        ordbase += total;
                        continue;
                    }
                    publish_step(off_s, hb_s, lane, excl, hb, hc);
                    for (uint32_t x0 = 0; x0 < total && !defer; x0 += 32 * kHitUnroll) {
                        unsigned long long h[kHitUnroll];
#pragma unroll
                        for (int u = 0; u < kHitUnroll; ++u) {
                            h[u] = 0;
                            if (x0 + 32 * u < total) {                              // warp-uniform
                                const uint32_t x = x0 + 32 * u + lane;
                                if (x < total) h[u] = __ldg(P.hits + step_hit_index(off_s, hb_s, x));
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kHitUnroll; ++u) {
                            if (defer || x0 + 32 * u >= total) continue;                // warp-uniform
                            const uint32_t x = x0 + 32 * u + lane;
                            unsigned long long e = 0;
                            bool keep = false;
                            if (x < total) {
                                const uint32_t q = (uint32_t)(h[u] >> 32);
                                const uint32_t cq = count_of(cnt, q);
                                keep = may_pass(qthr, cq, q, flen);
                                if (direct && keep && cq == 1) { emit(P, q, 1, len, id); keep = false; }
                                e = ((unsigned long long)q << 44) | ((unsigned long long)(ordbase + x) << 22) | (h[u] & 0x3fffffu);
                            }
                            const uint32_t bal = __ballot_sync(FULL, keep);
                            if (nsurv + __popc(bal) > (uint32_t)kGCap) { defer = true; continue; }
                            if (keep) wb[nsurv + __popc(bal & ((1u << lane) - 1u))] = e;
                            nsurv += __popc(bal);
                        }
                    }
                    __syncwarp();
                    ordbase += total;
                }
            }
            __syncwarp();
            for (int i = 4 * lane; i < cnt_words; i += 128) *reinterpret_cast<uint4*>(cnt + i) = make_uint4(0u, 0u, 0u, 0u);
            // Up to 32 survivors (the usual flagged sequence: a handful of hits of one or two queries): rank sort in registers -- the keys are
            // distinct (the order field numbers the hits), a survivor's place is the number of smaller keys -- written straight into the
            // shared buffer.  The bitonic network below costs 15 shared-memory passes however few entries there are: 17 % of the
            // stall samples of a short-sequence chunk (profiles/r04u_pf_short_chunk_lines.txt).
            bool sorted_small = false;
            if (!defer && wb != buf && nsurv > 0u && nsurv <= 32u) {
                const unsigned long long e = (uint32_t)lane < nsurv ? wb[lane] : ~0ull;
                uint32_t rank = 0;
                for (uint32_t x = 0; x < nsurv; ++x) rank += __shfl_sync(FULL, e, x) < e ? 1u : 0u;
                if ((uint32_t)lane < nsurv) buf[rank] = e;
                wb = buf;
                sorted_small = true;
            } else if (!defer && wb != buf && nsurv <= (1u << (31 - __clz(scap)))) {       // few survivors: sort them in shared memory (the bitonic sort pads to a power of two)
                for (int i = lane; i < (int)nsurv; i += 32) buf[i] = wb[i];
                wb = buf;
            }
            __syncwarp();
            if (defer) {
                if (lane == 0) {
                    const unsigned long long slot = atomicAdd(P.counters + 1, 1ull);
                    const unsigned long long off = atomicAdd(P.counters + 2, (unsigned long long)T);
                    if (slot < P.max_deferred && off + T <= P.pool_cap) { P.def_seq[slot] = (uint32_t)s; P.def_off[slot] = off; }
                    else atomicOr(P.counters + 3, 2ull);
                }
                continue;
            }
            if (nsurv == 0) continue;
            const int S = (int)nsurv;
            // bitonic sort of wb[0..P2) by (query, emission order)
            int P2 = 32;
            while (P2 < S) P2 <<= 1;
            if (!sorted_small) for (int i = S + lane; i < P2; i += 32) wb[i] = ~0ull;
            __syncwarp();
            for (int size = 2; !sorted_small && size <= P2; size <<= 1) {
                for (int stride = size >> 1, lg = 31 - __clz(size >> 1); stride > 0; stride >>= 1, --lg) {
                    for (int t = lane; t < (P2 >> 1); t += 32) {
                        const int lo = ((t >> lg) << (lg + 1)) | (t & (stride - 1));
                        const int hi = lo + stride;
                        const bool up = ((lo & size) == 0);
                        const unsigned long long x = wb[lo], y = wb[hi];
                        if ((x > y) == up) { wb[lo] = y; wb[hi] = x; }
                    }
                    __syncwarp();
                }
            }
            // run starts first (one ballot word per 32 entries, kept by lane i / 32; kGCap <= 1024), because the in-place
            // LIS below overwrites the entries of finished runs
            uint32_t my_starts = 0;
            for (int base = 0; base < S; base += 32) {
                const int i = base + lane;
                const bool st = i < S && (i == 0 || (wb[i] >> 44) != (wb[i - 1] >> 44));
                const uint32_t bal = __ballot_sync(FULL, st);
                if (lane == (base >> 5)) my_starts = bal;
            }
            __syncwarp();
            for (int base = 0; base < S; base += 32) {
                const uint32_t bal = __shfl_sync(FULL, my_starts, base >> 5);
                const bool st = (bal >> lane) & 1u;
                const int i = base + lane;
                uint32_t q = 0;
                int n = 0;
                if (st) {
                    q = (uint32_t)(wb[i] >> 44);
                    int e = i + 1;
                    while (e < S && (uint32_t)(wb[e] >> 44) == q) ++e;
                    n = e - i;
                }
                __syncwarp();          // all run lengths of this block are known before any of its runs is overwritten
                if (st) emit(P, q, n == 1 ? 1 : lis_inplace(wb, i, n), len, id);
                __syncwarp();
            }
        }
    }
}

// Builds of the scan by CTA size (the register budget follows the resident warps).
template <int kThreads, bool kStage>
__global__ void __launch_bounds__(kThreads, 1) pf_scan_kernel(PfParams P, int scap, int cnt_words) {
    pf_scan_body<kStage>(P, scap, cnt_words);
}

// deferred path, step 1: re-walk the sequence and write its hits into the pool
__global__ void __launch_bounds__(kWarps * 32) pf_fill_deferred_kernel(PfParams P, uint32_t n_def) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int k = P.k;
    for (uint32_t slot = blockIdx.x * kWarps + (threadIdx.x >> 5); slot < n_def; slot += gridDim.x * kWarps) {
        const long long s = P.def_seq[slot];
        const int64_t a = P.db_off[s];
        const int len = (int)(P.db_off[s + 1] - a);
        const uint8_t* seq = P.db_codes + a;
        const int npos = len - k + 1;
        unsigned long long* keys = P.pool_keys + P.def_off[slot];
        uint32_t* vals = P.pool_vals + P.def_off[slot];
        uint32_t T = 0, carry = 0xffffffffu;
        for (int base = 0; base < npos; base += 32) {
            const int j = base + lane;
            const bool valid = j < npos;
            const uint32_t kmer = valid ? kmer_at(seq, j, k) : 0xfffffffeu;
            uint32_t prevk = __shfl_up_sync(FULL, kmer, 1);
            if (lane == 0) prevk = carry;
            carry = __shfl_sync(FULL, kmer, 31);
            uint32_t b = 0, c = 0;
            if (valid && !(j > 0 && kmer == prevk)) lookup(P, kmer, b, c);
            uint32_t total;
            const uint32_t excl = warp_excl_scan(c, lane, total);
            for (uint32_t t = 0; t < c; ++t) {
                const unsigned long long h = __ldg(P.hits + b + t);
                const uint32_t ord = T + excl + t;
                keys[ord] = ((unsigned long long)slot << 48) | ((h >> 32) << 28) | (unsigned long long)(ord & 0xfffffffu);
                vals[ord] = (uint32_t)h;
            }
            T += total;
        }
    }
}

// deferred path, step 2 (after a radix sort of the pool): one thread per run of equal (slot, query)
__global__ void pf_lis_deferred_kernel(PfParams P, const unsigned long long* keys, const uint32_t* vals, uint32_t* tails,
                                       unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long g = keys[i] >> 28;
    if (i > 0 && (keys[i - 1] >> 28) == g) return;
    unsigned long long e = i + 1;
    int len = 0;
    uint32_t* tl = tails + i;
    for (unsigned long long x = i; x < n && (keys[x] >> 28) == g; ++x) {
        const uint32_t v = vals[x];
        int lo = 0, hi = len;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (tl[mid] < v) lo = mid + 1; else hi = mid; }
        tl[lo] = v;
        if (lo == len) ++len;
        e = x + 1;
    }
    (void)e;
    const uint32_t slot = (uint32_t)(g >> 20), q = (uint32_t)(g & 0xfffffu);
    const long long s = P.def_seq[slot];
    const int slen = (int)(P.db_off[s + 1] - P.db_off[s]);
    emit(P, q, len, slen, P.id_base + (uint32_t)s);
}

// ---- index build -----------------------------------------------------------------------------------------

__global__ void ix_kmers_kernel(const uint8_t* q_codes, const int64_t* q_off, int nq, int k, uint32_t mask,
                                const int64_t* hit_start, uint32_t* keys, unsigned long long* vals, uint32_t* bits) {
    const int q = blockIdx.x;
    const int64_t a = q_off[q];
    const int len = (int)(q_off[q + 1] - a);
    const int npos = len - k + 1;
    for (int j = threadIdx.x; j < npos; j += blockDim.x) {
        uint32_t v = 0;
        for (int i = 0; i < k; ++i) v = (v << 5) | q_codes[a + j + i];
        keys[hit_start[q] + j] = v;
        vals[hit_start[q] + j] = ((unsigned long long)q << 32) | (unsigned)j;
        atomicOr(bits + (v >> 5), 1u << (v & 31u));
    }
}

__global__ void ix_count_kernel(const int64_t* q_off, int nq, int k, int64_t* cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > nq) return;
    int64_t c = 0;
    if (q < nq) { const int64_t len = q_off[q + 1] - q_off[q]; c = len >= k ? len - k + 1 : 0; }
    cnt[q] = c;
}

__global__ void ix_popc_kernel(const uint32_t* bits, uint32_t n_words, uint32_t* pc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) pc[i] = __popc(bits[i]);
}

__global__ void ix_bitrank_kernel(const uint32_t* bits, const uint32_t* prefix, uint32_t n_words, uint2* bitrank) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) bitrank[i] = make_uint2(bits[i], prefix[i]);
}

__global__ void ix_bucket_kernel(const uint32_t* sorted_keys, int64_t n, const uint2* bitrank, uint32_t* bucket_start, uint32_t n_distinct) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) bucket_start[n_distinct] = (uint32_t)n;
    if (i >= n) return;
    const uint32_t kmer = sorted_keys[i];
    if (i == 0 || sorted_keys[i - 1] != kmer) {
        const uint2 br = bitrank[kmer >> 5];
        const uint32_t r = br.y + __popc(br.x & ((1u << (kmer & 31u)) - 1u));
        bucket_start[r] = (uint32_t)i;
    }
}

// bucket entries of the scan's pass A, which only needs the QUERY of every hit: a k-mer with up to three hits carries their
// queries in its entry (bits 63..62 = number of hits, three 20-bit query ids below: query ids stay under 2^20), the others
// their bucket (size << 32 | first entry in `hits`; sizes stay under 2^30)
__global__ void ix_entry_kernel(const uint32_t* bucket_start, const unsigned long long* hits, uint32_t n_distinct, unsigned long long* entry) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_distinct) return;
    const uint32_t b = bucket_start[r], c = bucket_start[r + 1] - b;
    unsigned long long e = ((unsigned long long)c << 32) | b;
    if (c <= 3) {
        e = (unsigned long long)c << 62;
        for (uint32_t i = 0; i < c; ++i) e |= (hits[b + i] >> 32) << (20 * i);
    }
    entry[r] = e;
}

// ---- candidate buffers -------------------------------------------------------------------------------------

__global__ void cb_init_kernel(unsigned long long* thr, uint32_t* count, int nq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) { thr[q] = kNoThr; count[q] = 0; }
}

// list the queries whose buffer could overflow during the next chunk (or all non-empty ones when `all`).
// sel_cap > 0: buffers holding at most sel_cap keys go to a second list (n_seg[1], filled from the end of the arrays)
// that cb_topn_kernel reduces by selection in shared memory; the rest is sorted.
__global__ void cb_select_kernel(const uint32_t* count, int nq, uint32_t cap, uint32_t limit, int all, int64_t* seg_begin,
                                 int64_t* seg_end, uint32_t* seg_q, unsigned long long* n_seg, uint32_t sel_cap,
                                 const unsigned long long* thr, uint32_t N) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t c = count[q] < cap ? count[q] : cap;
    // a query without a cut-off gets its first one as soon as it holds more than N candidates
    if ((all && c > 0) || c > limit || (thr && c > N && thr[q] == kNoThr)) {
        if (c <= sel_cap) {
            const unsigned long long s = atomicAdd(n_seg + 1, 1ull);
            seg_q[nq - 1 - s] = (uint32_t)q;
            return;
        }
        const unsigned long long s = atomicAdd(n_seg, 1ull);
        seg_begin[s] = (int64_t)q * cap;
        seg_end[s] = (int64_t)q * cap + c;
        seg_q[s] = (uint32_t)q;
    }
}

// Keep the best N keys of a buffer WITHOUT sorting it (the scan only needs the set and the cut-off): one CTA per listed
// query, keys in shared memory, MSB-first radix select (8 passes of 8 bits, warp-aggregated histogram updates because
// score bytes are heavily repeated), then the keys <= the N-th smallest are written back in arbitrary order.
constexpr int kSelCap = 12288;           // keys per CTA (96 KB)
constexpr int kSelThreads = 512;

__global__ void __launch_bounds__(kSelThreads) cb_topn_kernel(unsigned long long* cand, uint32_t* count, unsigned long long* thr,
                                                               uint32_t cap, uint32_t N, const uint32_t* seg_q_end) {
    extern __shared__ unsigned long long sel_keys[];
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_remaining, s_out;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t q = seg_q_end[-(int)blockIdx.x];
    const uint32_t c = count[q] < cap ? count[q] : cap;
    unsigned long long* src = cand + (size_t)q * cap;
    if (c <= N) {                                    // nothing to drop
        if (tid == 0) count[q] = c;
        return;
    }
    for (uint32_t i = tid; i < c; i += kSelThreads) sel_keys[i] = src[i];
    if (tid == 0) { s_prefix = 0; s_remaining = N; s_out = 0; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (uint32_t i0 = 0; i0 < c; i0 += kSelThreads) {
            const uint32_t i = i0 + tid;
            unsigned long long k = 0;
            bool act = i < c;
            if (act) { k = sel_keys[i]; act = (k & himask) == prefix; }
            const unsigned m = __ballot_sync(FULL, act);
            if (act) {
                const uint32_t d = (uint32_t)(k >> shift) & 255u;
                const unsigned peers = __match_any_sync(m, d);
                if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&hist[d], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = hist[tid * 8 + j]; sum += loc[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
            const uint32_t excl = incl - sum;
            const uint32_t rem = s_remaining;
            __syncwarp();
            if (excl < rem && rem <= incl) {
                uint32_t r = rem - excl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (r <= loc[j]) { s_prefix = prefix | ((unsigned long long)(tid * 8 + j) << shift); s_remaining = r; break; }
                    r -= loc[j];
                }
            }
        }
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    for (uint32_t i0 = 0; i0 < c; i0 += kSelThreads) {
        const uint32_t i = i0 + tid;
        unsigned long long k = 0;
        bool keep = false;
        if (i < c) { k = sel_keys[i]; keep = k <= kth; }
        const unsigned m = __ballot_sync(FULL, keep);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_out, (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (keep) src[base + __popc(m & ((1u << lane) - 1u))] = k;
    }
    if (tid == 0) { count[q] = N; thr[q] = kth; }
}

// one CTA per compacted query: copy the best min(count, N) keys back from the sort output, set count and cut-off
__global__ void cb_truncate_kernel(unsigned long long* cand, const unsigned long long* sorted, uint32_t* count, unsigned long long* thr,
                                   uint32_t cap, uint32_t N, const uint32_t* seg_q) {
    const uint32_t q = seg_q[blockIdx.x];
    uint32_t c = count[q] < cap ? count[q] : cap;
    if (c > N) c = N;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) cand[(size_t)q * cap + i] = sorted[(size_t)q * cap + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        if (c >= N && N != 0xffffffffu) thr[q] = sorted[(size_t)q * cap + N - 1];
        count[q] = c;
    }
}

// final: sorted candidate keys -> (ids, scores) rows; optionally re-ordered by ascending id
__global__ void cb_output_kernel(const unsigned long long* cand, const uint32_t* count, uint32_t cap, uint32_t N, int nq,
                                 uint32_t* out_ids, float* out_scores, uint32_t* out_counts) {
    const int q = blockIdx.x;
    const uint32_t c = count[q] < N ? count[q] : N;
    if (threadIdx.x == 0) out_counts[q] = c;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        const unsigned long long key = cand[(size_t)q * cap + i];
        out_ids[(size_t)q * N + i] = (uint32_t)key;
        if (out_scores) out_scores[(size_t)q * N + i] = __uint_as_float(~(uint32_t)(key >> 32));
    }
}

// keys for the id-order pass: (id << 32 | score bits), sorted ascending per query
__global__ void cb_idkeys_kernel(unsigned long long* cand, const uint32_t* count, uint32_t cap, uint32_t N, int nq) {
    const int q = blockIdx.x;
    const uint32_t c = count[q] < N ? count[q] : N;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        const unsigned long long key = cand[(size_t)q * cap + i];
        cand[(size_t)q * cap + i] = (key << 32) | (key >> 32);
    }
}

__global__ void cb_output_byid_kernel(const unsigned long long* cand, const uint32_t* count, uint32_t cap, uint32_t N, int nq,
                                      uint32_t* out_ids, float* out_scores, uint32_t* out_counts) {
    const int q = blockIdx.x;
    const uint32_t c = count[q] < N ? count[q] : N;
    if (threadIdx.x == 0) out_counts[q] = c;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        const unsigned long long key = cand[(size_t)q * cap + i];
        out_ids[(size_t)q * N + i] = (uint32_t)(key >> 32);
        if (out_scores) out_scores[(size_t)q * N + i] = __uint_as_float(~(uint32_t)key);
    }
}

__global__ void mg_keys_kernel(int n_ranks, int nq, uint32_t N, const uint32_t* ids, const float* scores, const uint32_t* counts,
                               unsigned long long* cand, uint32_t* count, uint32_t cap) {
    const int q = blockIdx.x;
    uint32_t base = 0;
    for (int r = 0; r < n_ranks; ++r) {
        const uint32_t c = counts[(size_t)r * nq + q];
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
            const size_t src = ((size_t)r * nq + q) * N + i;
            cand[(size_t)q * cap + base + i] = cand_key(scores[src], ids[src]);
        }
        base += c;
    }
    if (threadIdx.x == 0) count[q] = base;
}

int segmented_sort(s4g_ctx* ctx, const unsigned long long* keys_in, unsigned long long* keys_out, int64_t n_items_bound, int n_seg,
                   const int64_t* d_begin, const int64_t* d_end) {
    size_t tmp = 0;
    if (n_items_bound >= ((int64_t)1 << 31)) { s4g_set_error(ctx, "segmented sort over %lld items (>= 2^31)", (long long)n_items_bound); return S4G_ERR_CAPACITY; }
    cub::DeviceSegmentedSort::SortKeys(nullptr, tmp, keys_in, keys_out, n_items_bound, (int64_t)n_seg, d_begin, d_end, ctx->stream);
    void* d_tmp = s4g_scratch(ctx, SLOT_PF_CUB, tmp);
    if (!d_tmp) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cub::DeviceSegmentedSort::SortKeys(d_tmp, tmp, keys_in, keys_out, n_items_bound, (int64_t)n_seg, d_begin, d_end, ctx->stream));
    ctx->launches += 3;
    return S4G_OK;
}

}  // namespace

// Compact (sort + keep best N) the listed queries.  `all` = every non-empty query (final pass).
// `select` = the order inside the buffers does not matter afterwards (intermediate compactions, and the final one
// when the output is re-sorted by id): buffers that fit in shared memory are reduced by cb_topn_kernel.
static int compact(s4g_ctx* ctx, int nq, uint32_t cap, uint32_t N, uint32_t limit, int all, unsigned long long* d_cand,
                   unsigned long long* d_cand_alt, uint32_t* d_count, unsigned long long* d_thr, int64_t* d_seg, uint32_t* d_seg_q,
                   unsigned long long* d_nseg, bool* did, bool select = false) {
    cudaStream_t st = ctx->stream;
    S4G_CUDA(ctx, cudaMemsetAsync(d_nseg, 0, 16, st));
    cb_select_kernel<<<(nq + 255) / 256, 256, 0, st>>>(d_count, nq, cap, limit, all, d_seg, d_seg + nq, d_seg_q, d_nseg,
                                                       select ? (uint32_t)kSelCap : 0u, (select && !all) ? d_thr : nullptr, N);
    S4G_CHECK_LAUNCH(ctx);
    unsigned long long h_n[2] = {0, 0};
    S4G_CUDA(ctx, cudaMemcpyAsync(h_n, d_nseg, 16, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    const unsigned long long h_nseg = h_n[0];
    if (did) *did = h_nseg + h_n[1] > 0;
    if (h_n[1] > 0) {
        const size_t smem = sizeof(unsigned long long) * kSelCap;
        S4G_CUDA(ctx, cudaFuncSetAttribute(cb_topn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cb_topn_kernel<<<(unsigned)h_n[1], kSelThreads, smem, st>>>(d_cand, d_count, d_thr, cap, N, d_seg_q + nq - 1);
        S4G_CHECK_LAUNCH(ctx);
    }
    if (h_nseg == 0) return S4G_OK;
    int rc = segmented_sort(ctx, d_cand, d_cand_alt, (int64_t)nq * cap, (int)h_nseg, d_seg, d_seg + nq);
    if (rc != S4G_OK) return rc;
    cb_truncate_kernel<<<(unsigned)h_nseg, 256, 0, st>>>(d_cand, d_cand_alt, d_count, d_thr, cap, N, d_seg_q);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}

namespace {
__global__ void ord_keys_kernel(const int64_t* off, int64_t n, uint32_t* keys, uint32_t* vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (uint32_t)(off[i + 1] - off[i]); vals[i] = (uint32_t)i; }
}
}  // namespace

// Scan order of a shard: sequences by ascending length (stable).  score = LIS / len, so the short sequences hold the best
// candidates of every query: scanning them first settles the cut-offs after a few percent of the residues, and the count
// filter then removes nearly every hit of the long sequences.  Built once per database (it is resident).
static int build_scan_order(s4g_ctx* ctx, s4g_db* db) {
    if (db->d_order || db->n == 0) return S4G_OK;
    const int64_t n = db->n;
    S4G_CUDA(ctx, cudaMalloc(&db->d_order, sizeof(uint32_t) * (size_t)n));
    uint32_t* tmp = (uint32_t*)s4g_scratch(ctx, SLOT_PF_TMP2, sizeof(uint32_t) * 3 * (size_t)n);
    if (!tmp) return S4G_ERR_NOMEM;
    uint32_t* keys = tmp, *keys2 = tmp + n, *vals = tmp + 2 * n;
    ord_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(db->d_off, n, keys, vals);
    S4G_CHECK_LAUNCH(ctx);
    int bits = 1;
    while (bits < 32 && (1ll << bits) <= (long long)db->max_len) ++bits;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys2, vals, db->d_order, n, 0, bits, ctx->stream);
    void* d_tmp = s4g_scratch(ctx, SLOT_PF_CUB, bytes);
    if (!d_tmp) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, bytes, keys, keys2, vals, db->d_order, n, 0, bits, ctx->stream));
    ctx->launches += 3;
    return S4G_OK;
}

// Shared memory of the scan kernel: what a CTA may use, the per-warp part for `nq` queries, and the largest batch that still
// leaves kMinScanWarps warps per SM beside the 2-byte cut-off table (larger batches are scanned in groups of queries: the
// candidate lists of different queries are independent, the database streams once per group).
constexpr size_t kScanSmemCap = 227 * 1024;
constexpr int kScanSortCap = 192;          // 8-byte entries of a warp's buffer: the queues of pass A (1 KB of ranks + 512 B of buckets; + 512 B of positions = kScanReplayCap when flagged sequences are replayed), step tables / sort buffer of pass B; every KB here is a KB less L1 for the index probes
constexpr int kMinScanWarps = 4;
static int scan_cnt_words(int nq) { return ((nq + 1) / 2 + 127) / 128 * 128; }
constexpr int kScanReplayCap = 256;
static int scan_sort_cap(int) { return kScanSortCap; }
static size_t scan_warp_smem(int nq) { return sizeof(uint32_t) * (size_t)scan_cnt_words(nq) + sizeof(unsigned long long) * scan_sort_cap(nq); }
static size_t scan_qthr_smem(int nq) { return sizeof(unsigned short) * (size_t)((nq + 7) & ~7); }
// Queries per scan: the exact counters cost 2 B per query and warp, and the scan lives on resident warps (one random index
// probe per position and warp instruction: it is bound by issue and latency, not by bandwidth).  Large batches are therefore
// scanned in groups of queries -- the candidate lists of different queries are independent, and one more pass over the
// resident database costs far less than running the whole batch at a few warps per SM.
static int scan_max_queries() {
    int group = 4096;
    if (const char* e = getenv("S4G_PF_GROUP")) group = atoi(e);
    int cap = 1024;
    while (scan_qthr_smem(cap + 1024) + kMinScanWarps * scan_warp_smem(cap + 1024) <= kScanSmemCap) cap += 1024;
    return std::max(256, std::min(group, cap));
}

static int prefilter_group(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int k, int max_candidates, int sorted_by_id,
                           uint32_t* d_ids, float* d_scores, uint32_t* d_counts);

int s4g_prefilter_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int k, int max_candidates, int sorted_by_id,
                         uint32_t* d_ids, float* d_scores, uint32_t* d_counts) {
    const int nq = q->n;
    if (nq >= (1 << 20)) { s4g_set_error(ctx, "at most 2^20-1 queries per batch"); return S4G_ERR_ARG; }
    if (q->max_len >= (1 << 22)) { s4g_set_error(ctx, "query longer than 2^22"); return S4G_ERR_ARG; }
    const int group = scan_max_queries();
    if (nq <= group) return prefilter_group(ctx, db, q, k, max_candidates, sorted_by_id, d_ids, d_scores, d_counts);
    // equal groups of queries, each a view of the resident batch
    const int n_groups = (nq + group - 1) / group;
    for (int g = 0; g < n_groups; ++g) {
        const int g0 = (int)((int64_t)nq * g / n_groups), g1 = (int)((int64_t)nq * (g + 1) / n_groups);
        s4g_queries view;
        view.ctx = q->ctx; view.d_codes = q->d_codes; view.d_off = q->d_off + g0; view.n = g1 - g0; view.max_len = q->max_len;
        view.h_off.assign(q->h_off.begin() + g0, q->h_off.begin() + g1 + 1);
        const int rc = prefilter_group(ctx, db, &view, k, max_candidates, sorted_by_id, d_ids + (size_t)g0 * max_candidates,
                                       d_scores ? d_scores + (size_t)g0 * max_candidates : nullptr, d_counts + g0);
        if (rc != S4G_OK) return rc;
    }
    return S4G_OK;
}

static int prefilter_group(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int k, int max_candidates, int sorted_by_id,
                           uint32_t* d_ids, float* d_scores, uint32_t* d_counts) {
    cudaStream_t st = ctx->stream;
    const int nq = q->n;
    const uint32_t N = (uint32_t)max_candidates;

    {
        int rc = build_scan_order(ctx, db);
        if (rc != S4G_OK) return rc;
    }
    s4g_trace_start(ctx);
    // ---- index over all queries ----
    int64_t n_hits = 0;
    for (int i = 0; i < nq; ++i) { int64_t len = q->h_off[i + 1] - q->h_off[i]; if (len >= k) n_hits += len - k + 1; }
    if (n_hits >= ((int64_t)1 << 31)) {      // index entries are addressed in 32 bits (bucket starts, cub item counts)
        s4g_set_error(ctx, "prefilter: %lld query k-mers exceed 2^31 index entries; split the query batch", (long long)n_hits);
        return S4G_ERR_CAPACITY;
    }
    const uint32_t n_kmer_space = 1u << (5 * k);
    const uint32_t n_words = n_kmer_space / 32;
    const uint32_t mask = n_kmer_space - 1;

    size_t ix_bytes = sizeof(int64_t) * 2 * (nq + 1) + sizeof(uint32_t) * 2 * (size_t)(n_hits + 1) + sizeof(unsigned long long) * 2 * (size_t)(n_hits + 1) + 64;
    char* ix = (char*)s4g_scratch(ctx, SLOT_PF_INDEX, ix_bytes);
    uint32_t* d_bits = (uint32_t*)s4g_scratch(ctx, SLOT_PF_BITMAP, sizeof(uint32_t) * 3 * (size_t)n_words);
    uint2* d_bitrank = (uint2*)s4g_scratch(ctx, SLOT_PF_RANK, sizeof(uint2) * (size_t)n_words);
    uint32_t* d_bucket = (uint32_t*)s4g_scratch(ctx, SLOT_PF_BUCKET, sizeof(uint32_t) * (size_t)(n_hits + 2));
    if (!ix || !d_bits || !d_bitrank || !d_bucket) return S4G_ERR_NOMEM;
    int64_t* d_cnt = (int64_t*)ix;
    int64_t* d_start = d_cnt + (nq + 1);
    unsigned long long* d_vals = (unsigned long long*)(d_start + (nq + 1));
    unsigned long long* d_vals2 = d_vals + (n_hits + 1);
    uint32_t* d_keys = (uint32_t*)(d_vals2 + (n_hits + 1));
    uint32_t* d_keys2 = d_keys + (n_hits + 1);
    uint32_t* d_pc = d_bits + n_words;
    uint32_t* d_prefix = d_pc + n_words;

    S4G_CUDA(ctx, cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * n_words, st));
    ix_count_kernel<<<(nq + 1 + 255) / 256, 256, 0, st>>>(q->d_off, nq, k, d_cnt);
    S4G_CHECK_LAUNCH(ctx);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_cnt, d_start, nq + 1, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_PF_CUB, tmp);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_cnt, d_start, nq + 1, st));
        ctx->launches += 1;
    }
    uint32_t n_distinct = 0;
    if (n_hits > 0) {
        ix_kmers_kernel<<<nq, 128, 0, st>>>(q->d_codes, q->d_off, nq, k, mask, d_start, d_keys, d_vals, d_bits);
        S4G_CHECK_LAUNCH(ctx);
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_keys, d_keys2, d_vals, d_vals2, (int)n_hits, 0, 5 * k, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_PF_CUB, tmp);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp, d_keys, d_keys2, d_vals, d_vals2, (int)n_hits, 0, 5 * k, st));
        ctx->launches += 3;
    }
    ix_popc_kernel<<<(n_words + 255) / 256, 256, 0, st>>>(d_bits, n_words, d_pc);
    S4G_CHECK_LAUNCH(ctx);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_pc, d_prefix, (int)n_words, st);
        void* d_tmp = s4g_scratch(ctx, SLOT_PF_CUB, tmp);
        if (!d_tmp) return S4G_ERR_NOMEM;
        S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_pc, d_prefix, (int)n_words, st));
        ctx->launches += 1;
    }
    ix_bitrank_kernel<<<(n_words + 255) / 256, 256, 0, st>>>(d_bits, d_prefix, n_words, d_bitrank);
    S4G_CHECK_LAUNCH(ctx);
    {
        uint32_t last_pc = 0, last_prefix = 0;
        S4G_CUDA(ctx, cudaMemcpyAsync(&last_pc, d_pc + n_words - 1, 4, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(&last_prefix, d_prefix + n_words - 1, 4, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaStreamSynchronize(st));
        n_distinct = last_pc + last_prefix;
    }
    if (n_hits > 0) {
        ix_bucket_kernel<<<(unsigned)((n_hits + 255) / 256), 256, 0, st>>>(d_keys2, n_hits, d_bitrank, d_bucket, n_distinct);
        S4G_CHECK_LAUNCH(ctx);
    }
    // ---- scan configuration: warps per SM from the exact counters (2 B per query and warp) ----
    const int cnt_words = scan_cnt_words(nq);
    // Flagged sequences are replayed from the rank queue (pass B without a second walk) where that pays: the index holds a minority
    // of the k-mer space in small buckets -- a replay round takes 32 found k-mers, a re-walk step 128 positions, so the replay wins
    // while fewer than ~1 position in 4 hits the index (configs[1]: 1 in 7; measured at configs[2]'s groups of 4 000 human-sized
    // queries, 1 in 2: replay 262 ms against 254 ms; configs[3]: every position hits) -- and where the 512 B of positions per warp cost no
    // resident warp.  S4G_PF_REPLAY=0/1 overrides the density rule.
    bool replay = n_hits <= 2 * (int64_t)n_distinct && n_distinct <= (1u << 20);
    { const char* e = getenv("S4G_PF_REPLAY"); if (e) replay = e[0] != '0'; }
    const size_t qthr_bytes = scan_qthr_smem(nq);
    {
        const size_t base_w = scan_warp_smem(nq), replay_w = base_w + sizeof(unsigned long long) * (kScanReplayCap - kScanSortCap);
        if (std::min<size_t>(32, (kScanSmemCap - qthr_bytes) / replay_w) < std::min<size_t>(32, (kScanSmemCap - qthr_bytes) / base_w)) replay = false;
    }
    const int scap = replay ? kScanReplayCap : scan_sort_cap(nq);
    const size_t per_warp_smem = scan_warp_smem(nq) + sizeof(unsigned long long) * (scap - scan_sort_cap(nq));
    int scan_warps = (int)std::min<size_t>(32, (kScanSmemCap - qthr_bytes) / per_warp_smem);
    if (const char* e = getenv("S4G_PF_WARPS")) scan_warps = std::max(1, std::min(scan_warps, atoi(e)));
    // TMA residue staging (pf_scan_body<true>): 0.6 KB per warp.  For an NVLink-striped view (the copies hide the peer latency).
    // A resident database is scanned without: measured at configs[1], staging costs 27.8 -> 30.8 ms there (and 41.9 ms with a
    // 2.3 KB ring: the shared memory comes out of the L1 that serves the index probes) -- profiles/r02_stage_tma.md.
    // S4G_PF_STAGE=0/1 overrides.
    const size_t stage_pad = 16;
    int stage_warps = (int)std::min<size_t>(32, (kScanSmemCap - qthr_bytes - stage_pad) / (per_warp_smem + kStageWarpBytes));
    if (const char* e = getenv("S4G_PF_WARPS")) stage_warps = std::max(1, std::min(stage_warps, atoi(e)));
    bool stage = stage_warps >= 1 && db->borrowed_codes;
    if (const char* e = getenv("S4G_PF_STAGE")) stage = atoi(e) != 0 && stage_warps >= 1;
    if (stage) scan_warps = stage_warps;
    unsigned long long* d_entry = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_ENTRY, sizeof(unsigned long long) * ((size_t)n_distinct + 1));
    if (!d_entry) return S4G_ERR_NOMEM;
    if (n_distinct > 0) {
        ix_entry_kernel<<<(n_distinct + 255) / 256, 256, 0, st>>>(d_bucket, d_vals2, n_distinct, d_entry);
        S4G_CHECK_LAUNCH(ctx);
    }

    s4g_trace_mark(ctx, "index");
    // ---- candidate buffers ----
    // chunk of sequences per scan launch; every buffer can take a whole chunk on top of N + slack
    const uint32_t slack = N / 4 > 256 ? N / 4 : 256;
    // the scan runs in chunks of sequences; chunk sizes grow geometrically (cut-offs settle on the first small
    // chunks, later chunks append little) up to what the candidate-buffer budget allows
    int64_t chunk = 1 << 20;
    // bytes for both candidate buffers: a third of the free HBM, 12 .. 48 GiB.  The driver is only asked when the
    // buffers this context already holds cannot take full-size chunks (cudaMemGetInfo is a device-wide query that
    // can stall for milliseconds; a steady-state call must not pay for it).
    size_t budget = ctx->slot_bytes[SLOT_PF_CAND] + ctx->slot_bytes[SLOT_PF_TMP];
    if ((size_t)nq * (size_t)(N + slack + chunk) * 16 > budget && !(ctx->pf_budget_nq == nq && ctx->pf_budget_n == N)) {
        budget = (size_t)12 << 30;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += ctx->slot_bytes[SLOT_PF_CAND] + ctx->slot_bytes[SLOT_PF_TMP];      // what this call may reuse
            budget = std::min<size_t>(std::max<size_t>(free_b / 3, (size_t)12 << 30), (size_t)48 << 30);
            if (budget > free_b / 2) budget = free_b / 2;
        }
        ctx->pf_budget = budget; ctx->pf_budget_nq = nq; ctx->pf_budget_n = N;
    } else if (ctx->pf_budget_nq == nq && ctx->pf_budget_n == N) {
        budget = std::max(budget, ctx->pf_budget);     // same batch shape as the last query of the driver: same answer
    }
    // (nq x cap keys must also stay below 2^31: the segmented sort of the compactions counts items in 32 bits)
    const size_t max_items = ((size_t)1 << 31) - 1;
    while (chunk > 4096 && ((size_t)nq * (size_t)(N + slack + chunk) * 16 > budget || (size_t)nq * (size_t)(N + slack + chunk) > max_items)) chunk >>= 1;
    if ((size_t)nq * (size_t)(N + slack + chunk) > max_items) {
        s4g_set_error(ctx, "prefilter: %d queries x %u candidates exceed 2^31 candidate-buffer entries; split the query batch", nq, N);
        return S4G_ERR_CAPACITY;
    }
    if (chunk > db->n) chunk = db->n > 0 ? db->n : 1;
    const uint32_t cap = (uint32_t)(N + slack + chunk);
    unsigned long long* d_cand = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_CAND, sizeof(unsigned long long) * (size_t)nq * cap);
    unsigned long long* d_cand_alt = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_TMP, sizeof(unsigned long long) * (size_t)nq * cap);
    uint32_t* d_count = (uint32_t*)s4g_scratch(ctx, SLOT_PF_COUNT, sizeof(uint32_t) * 2 * nq);
    uint32_t* d_count_saved = d_count ? d_count + nq : nullptr;
    unsigned long long* d_thr = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_THR, sizeof(unsigned long long) * nq + 64);
    const uint32_t max_deferred = 1u << 16;
    const unsigned long long pool_cap = 1ull << 27;   // 128 M hits per chunk through the deferred path (28-bit order field)
    char* spill = (char*)s4g_scratch(ctx, SLOT_PF_SPILL, 64 + sizeof(int64_t) * 2 * nq + sizeof(uint32_t) * nq + (sizeof(uint32_t) + sizeof(unsigned long long)) * max_deferred);
    if (!d_cand || !d_cand_alt || !d_count || !d_thr || !spill) return S4G_ERR_NOMEM;
    unsigned long long* d_counters = (unsigned long long*)spill;          // 8 counters
    int64_t* d_seg = (int64_t*)(spill + 64);
    unsigned long long* d_def_off = (unsigned long long*)(d_seg + 2 * nq);
    uint32_t* d_seg_q = (uint32_t*)(d_def_off + max_deferred);
    uint32_t* d_def_seq = d_seg_q + nq;
    unsigned long long* d_nseg = d_counters + 4;

    cb_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(d_thr, d_count, nq);
    S4G_CHECK_LAUNCH(ctx);

    PfParams P;
    P.db_codes = db->d_codes; P.local_lo = db->local_lo; P.local_hi = db->local_hi; P.db_off = db->d_off; P.order = db->d_order; P.id_base = db->id_base;
    P.k = k; P.mask = mask; P.bitrank = d_bitrank; P.bucket_start = d_bucket; P.hits = d_vals2; P.nq = nq;
    P.thr = d_thr; P.count = d_count; P.cand = d_cand; P.cap = cap; P.counters = d_counters;
    P.def_seq = d_def_seq; P.def_off = d_def_off; P.max_deferred = max_deferred;
    P.pool_keys = nullptr; P.pool_vals = nullptr; P.pool_cap = pool_cap;

    P.entry = d_entry; P.q_hit_start = d_start;
    P.trace = ctx->trace ? 1 : 0;
    P.no_replay = replay ? 0 : 1;
    // one CTA per SM: its warps share the cut-off table and the filter; the build follows the CTA size (register budget)
    const size_t scan_smem = per_warp_smem * scan_warps + qthr_bytes + (stage ? stage_pad + (size_t)kStageWarpBytes * scan_warps : 0);
    void (*scan_kernel)(PfParams, int, int) = nullptr;
    if (stage) scan_kernel = scan_warps > 24 ? pf_scan_kernel<1024, true> : (scan_warps > 16 ? pf_scan_kernel<768, true> : (scan_warps > 8 ? pf_scan_kernel<512, true> : pf_scan_kernel<256, true>));
    else scan_kernel = scan_warps > 24 ? pf_scan_kernel<1024, false> : (scan_warps > 16 ? pf_scan_kernel<768, false> : (scan_warps > 8 ? pf_scan_kernel<512, false> : pf_scan_kernel<256, false>));
    if (const char* e = getenv("S4G_PF_BUILD")) { if (atoi(e) == 1024) scan_kernel = stage ? pf_scan_kernel<1024, true> : pf_scan_kernel<1024, false>; }   // experiment: 64-register build at any CTA size
    S4G_CUDA(ctx, cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
    const int grid = ctx->sm_count;
    P.gbuf = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_GBUF, sizeof(unsigned long long) * (size_t)grid * scan_warps * kGCap);
    if (!P.gbuf) return S4G_ERR_NOMEM;
    if (ctx->trace) fprintf(stderr, "[s4g trace] scan: %d queries, %u index k-mers (%lld hits), %d warps per SM, counters %d B per warp, %zu B of shared memory, residues %s\n",
                            nq, n_distinct, (long long)n_hits, scan_warps, cnt_words * 4, scan_smem, stage ? "staged by TMA bulk copies" : "loaded by the lanes");

    s4g_trace_mark(ctx, "setup");
    int64_t this_chunk = chunk < 16384 ? chunk : 16384;
    int64_t ceiling = chunk;
    int ceiling_hold = 0;
    for (int64_t s0 = 0; s0 < db->n && n_hits > 0; ) {
        P.seq_begin = s0;
        P.seq_end = s0 + this_chunk < db->n ? s0 + this_chunk : db->n;
        S4G_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 64, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(d_count_saved, d_count, sizeof(uint32_t) * nq, cudaMemcpyDeviceToDevice, st));
        // the pool is only allocated once a chunk needs it (first pass counts; see below)
        P.pool_keys = (unsigned long long*)ctx->slot_ptr[SLOT_PF_HITS];
        const auto t_chunk = std::chrono::steady_clock::now();
        scan_kernel<<<grid, scan_warps * 32, scan_smem, st>>>(P, scap, cnt_words);
        S4G_CHECK_LAUNCH(ctx);
        unsigned long long h_c[8];
        S4G_CUDA(ctx, cudaMemcpyAsync(h_c, d_counters, 64, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaStreamSynchronize(st));
        s4g_trace_mark(ctx, "scan");
        if (ctx->trace) fprintf(stderr, "[s4g trace] chunk [%lld,%lld): scan %.3f ms, deferred %llu sequences, %llu hits; %llu flagged sequences re-walked, %llu replayed from the rank queue\n", (long long)P.seq_begin, (long long)P.seq_end,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_chunk).count(), h_c[1], h_c[2], h_c[5], h_c[6]);
        if (h_c[3] & 1ull) { s4g_set_error(ctx, "prefilter: candidate buffer overflow (internal)"); return S4G_ERR_INTERNAL; }
        if (h_c[3] & 2ull) {
            if (P.seq_end - P.seq_begin <= 256) { s4g_set_error(ctx, "prefilter: more than %u deferred sequences or %llu deferred hits in a chunk of %lld sequences", max_deferred, pool_cap, (long long)(P.seq_end - P.seq_begin)); return S4G_ERR_CAPACITY; }
            S4G_CUDA(ctx, cudaMemcpyAsync(d_count, d_count_saved, sizeof(uint32_t) * nq, cudaMemcpyDeviceToDevice, st));
            this_chunk = (P.seq_end - P.seq_begin) / 2;
            ceiling = this_chunk;
            ceiling_hold = 8;
            continue;
        }
        s0 = P.seq_end;
        if (ceiling_hold > 0 && --ceiling_hold == 0) ceiling = chunk;
        this_chunk = std::min<int64_t>(this_chunk * 2, std::min(chunk, ceiling));
        if (h_c[1] > 0) {
            const unsigned long long n_pool = h_c[2];
            const uint32_t n_def = (uint32_t)h_c[1];
            char* pool = (char*)s4g_scratch(ctx, SLOT_PF_HITS, (sizeof(unsigned long long) * 2 + sizeof(uint32_t) * 3) * (size_t)n_pool + 256);
            if (!pool) return S4G_ERR_NOMEM;
            unsigned long long* pk = (unsigned long long*)pool;
            unsigned long long* pk2 = pk + n_pool;
            uint32_t* pv = (uint32_t*)(pk2 + n_pool);
            uint32_t* pv2 = pv + n_pool;
            uint32_t* tails = pv2 + n_pool;
            P.pool_keys = pk; P.pool_vals = pv;
            pf_fill_deferred_kernel<<<ctx->sm_count * 2, kWarps * 32, 0, st>>>(P, n_def);
            S4G_CHECK_LAUNCH(ctx);
            size_t tmp = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp, pk, pk2, pv, pv2, (int64_t)n_pool, 0, 64, st);
            void* d_tmp = s4g_scratch(ctx, SLOT_PF_TMP2, tmp);
            if (!d_tmp) return S4G_ERR_NOMEM;
            S4G_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tmp, pk, pk2, pv, pv2, (int64_t)n_pool, 0, 64, st));
            ctx->launches += 8;
            pf_lis_deferred_kernel<<<(unsigned)((n_pool + 127) / 128), 128, 0, st>>>(P, pk2, pv2, tails, n_pool);
            S4G_CHECK_LAUNCH(ctx);
        }
        s4g_trace_mark(ctx, "deferred");
        // compact buffers that could overflow in the next chunk
        const bool last = P.seq_end >= db->n;
        if (!last) {
            int rc = compact(ctx, nq, cap, N, N + slack, 0, d_cand, d_cand_alt, d_count, d_thr, d_seg, d_seg_q, d_nseg, nullptr, true);
            if (rc != S4G_OK) return rc;
        }
        s4g_trace_mark(ctx, "compact");
    }
    s4g_trace_mark(ctx, "compact");
    if (ctx->trace) {
        std::vector<unsigned long long> h_thr(nq);
        cudaMemcpy(h_thr.data(), d_thr, sizeof(unsigned long long) * nq, cudaMemcpyDeviceToHost);
        std::vector<float> sc;
        int none = 0;
        for (int i = 0; i < nq; ++i) { if (h_thr[i] == kNoThr) ++none; else { uint32_t b = ~(uint32_t)(h_thr[i] >> 32); float f; memcpy(&f, &b, 4); sc.push_back(f); } }
        std::sort(sc.begin(), sc.end());
        if (!sc.empty()) fprintf(stderr, "[s4g trace] cut-off scores: %d queries without, min %.4f p10 %.4f median %.4f max %.4f\n", none, sc.front(), sc[sc.size() / 10], sc[sc.size() / 2], sc.back());
    }
    // ---- final top-N, output ----
    {
        int rc = compact(ctx, nq, cap, N, 0, 1, d_cand, d_cand_alt, d_count, d_thr, d_seg, d_seg_q, d_nseg, nullptr, sorted_by_id != 0);
        if (rc != S4G_OK) return rc;
    }
    if (!sorted_by_id) {
        cb_output_kernel<<<nq, 128, 0, st>>>(d_cand, d_count, cap, N, nq, d_ids, d_scores, d_counts);
        S4G_CHECK_LAUNCH(ctx);
    } else {
        cb_idkeys_kernel<<<nq, 128, 0, st>>>(d_cand, d_count, cap, N, nq);
        S4G_CHECK_LAUNCH(ctx);
        bool did = false;
        // sort every non-empty query by the swapped key (id major)
        int rc = compact(ctx, nq, cap, 0xffffffffu, 0, 1, d_cand, d_cand_alt, d_count, d_thr, d_seg, d_seg_q, d_nseg, &did);
        if (rc != S4G_OK) return rc;
        cb_output_byid_kernel<<<nq, 128, 0, st>>>(d_cand, d_count, cap, N, nq, d_ids, d_scores, d_counts);
        S4G_CHECK_LAUNCH(ctx);
    }
    s4g_trace_mark(ctx, "final");
    s4g_trace_report(ctx, "prefilter");
    return S4G_OK;
}

extern "C" int s4g_merge_candidates(s4g_ctx* ctx, int n_ranks, int n_queries, int max_candidates, const uint32_t* gathered_ids,
                                    const float* gathered_scores, const uint32_t* gathered_counts, uint32_t* out_ids,
                                    float* out_scores, uint32_t* out_counts) {
    if (!ctx || n_ranks < 1 || n_queries < 1 || max_candidates < 1 || !gathered_ids || !gathered_scores || !gathered_counts || !out_ids || !out_counts)
        return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int nq = n_queries;
    const uint32_t N = (uint32_t)max_candidates, cap = N * (uint32_t)n_ranks;
    unsigned long long* d_cand = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_CAND, sizeof(unsigned long long) * (size_t)nq * cap);
    unsigned long long* d_cand_alt = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_TMP, sizeof(unsigned long long) * (size_t)nq * cap);
    uint32_t* d_count = (uint32_t*)s4g_scratch(ctx, SLOT_PF_COUNT, sizeof(uint32_t) * nq);
    unsigned long long* d_thr = (unsigned long long*)s4g_scratch(ctx, SLOT_PF_THR, sizeof(unsigned long long) * nq + 64);
    char* spill = (char*)s4g_scratch(ctx, SLOT_PF_SPILL, 64 + sizeof(int64_t) * 2 * nq + sizeof(uint32_t) * nq);
    if (!d_cand || !d_cand_alt || !d_count || !d_thr || !spill) return S4G_ERR_NOMEM;
    unsigned long long* d_nseg = (unsigned long long*)spill;
    int64_t* d_seg = (int64_t*)(spill + 64);
    uint32_t* d_seg_q = (uint32_t*)(d_seg + 2 * nq);
    mg_keys_kernel<<<nq, 128, 0, st>>>(n_ranks, nq, N, gathered_ids, gathered_scores, gathered_counts, d_cand, d_count, cap);
    S4G_CHECK_LAUNCH(ctx);
    int rc = compact(ctx, nq, cap, N, 0, 1, d_cand, d_cand_alt, d_count, d_thr, d_seg, d_seg_q, d_nseg, nullptr);
    if (rc != S4G_OK) return rc;
    cb_idkeys_kernel<<<nq, 128, 0, st>>>(d_cand, d_count, cap, N, nq);
    S4G_CHECK_LAUNCH(ctx);
    rc = compact(ctx, nq, cap, 0xffffffffu, 0, 1, d_cand, d_cand_alt, d_count, d_thr, d_seg, d_seg_q, d_nseg, nullptr);
    if (rc != S4G_OK) return rc;
    cb_output_byid_kernel<<<nq, 128, 0, st>>>(d_cand, d_count, cap, N, nq, out_ids, out_scores, out_counts);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}

// ---- multi-GPU, query-owner protocol ------------------------------------------------------------------------
// Every rank owns a slice of the queries.  After an all-to-all the owner holds, for each of its queries, the
// best-first rows of all shards; it only has to find the global cut-off key (the max_candidates-th smallest
// (score desc, id asc) key of the union) -- database_search.cpp:132-154 keeps exactly the rows up to it.  Rows stay
// where they are: each shard then keeps the prefix of its own row that lies within the cut-off.
namespace {

__device__ __forceinline__ uint32_t rows_upper_bound(const uint32_t* ids, const float* sc, uint32_t a, uint32_t b, unsigned long long key) {
    while (a < b) {
        const uint32_t m = (a + b) >> 1;
        if (cand_key(sc[m], ids[m]) <= key) a = m + 1; else b = m;
    }
    return a;
}

// one warp per query, lane r walks shard r's row.  Bisection on the key value; every lane keeps the window of its
// row that can still contain the boundary, so the searches shrink with the key range.
__global__ void __launch_bounds__(kWarps * 32) mg_cutoff_kernel(int n_ranks, int nq, uint32_t N, const uint32_t* ids, const float* scores,
                                                                const uint32_t* counts, unsigned long long* cutoff) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (q >= nq) return;
    const unsigned FULL = 0xffffffffu;
    uint32_t c = 0;
    const uint32_t* my_ids = ids;
    const float* my_sc = scores;
    if (lane < n_ranks) {
        c = counts[(size_t)lane * nq + q];
        if (c > N) c = N;
        my_ids = ids + ((size_t)lane * nq + q) * N;
        my_sc = scores + ((size_t)lane * nq + q) * N;
    }
    if (__reduce_add_sync(FULL, c) <= N) {                 // nothing to drop
        if (lane == 0) cutoff[q] = ~0ull;
        return;
    }
    unsigned long long lo = 0, hi = ~0ull;
    uint32_t a = 0, b = c;
    while (lo < hi) {
        const unsigned long long mid = lo + ((hi - lo) >> 1);
        const uint32_t p = rows_upper_bound(my_ids, my_sc, a, b, mid);
        if (__reduce_add_sync(FULL, p) >= N) { hi = mid; b = p; } else { lo = mid + 1; a = p; }
    }
    if (lane == 0) cutoff[q] = lo;
}

__global__ void mg_within_kernel(int nq, uint32_t N, const uint32_t* ids, const float* scores, const uint32_t* counts,
                                 const unsigned long long* cutoff, uint32_t* out_counts) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t c = counts[q];
    if (c > N) c = N;
    out_counts[q] = rows_upper_bound(ids + (size_t)q * N, scores + (size_t)q * N, 0, c, cutoff[q]);
}

}  // namespace

extern "C" int s4g_topn_cutoff(s4g_ctx* ctx, int n_ranks, int n_queries, int max_candidates, const uint32_t* gathered_ids,
                               const float* gathered_scores, const uint32_t* gathered_counts, uint64_t* out_cutoff) {
    if (!ctx || n_ranks < 1 || n_ranks > 32 || n_queries < 0 || max_candidates < 1 || !gathered_ids || !gathered_scores || !gathered_counts || !out_cutoff)
        return S4G_ERR_ARG;
    if (n_queries == 0) return S4G_OK;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    mg_cutoff_kernel<<<(n_queries + kWarps - 1) / kWarps, kWarps * 32, 0, ctx->stream>>>(
        n_ranks, n_queries, (uint32_t)max_candidates, gathered_ids, gathered_scores, gathered_counts, (unsigned long long*)out_cutoff);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}

extern "C" int s4g_cutoff_counts(s4g_ctx* ctx, int n_queries, int max_candidates, const uint32_t* ids, const float* scores,
                                 const uint32_t* counts, const uint64_t* cutoff, uint32_t* out_counts) {
    if (!ctx || n_queries < 0 || max_candidates < 1 || !ids || !scores || !counts || !cutoff || !out_counts) return S4G_ERR_ARG;
    if (n_queries == 0) return S4G_OK;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    mg_within_kernel<<<(n_queries + 127) / 128, 128, 0, ctx->stream>>>(n_queries, (uint32_t)max_candidates, ids, scores, counts,
                                                                      (const unsigned long long*)cutoff, out_counts);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}
