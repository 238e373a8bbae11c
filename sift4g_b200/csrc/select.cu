// Between stage 2 and 3: E-values and hit selection, on the host side of the boundary (threaded C++).
// Restates createEValueParams / calculateEValueProt (vendor/swsharp/swsharp/src/evalue.cu:73-88,148-220,
// 436-489) and extractThread's ordering (database.c:847-869,1043-1059).  The C++ shims that link the
// reference's host objects call the reference's own eValues() instead (sift4g_b200/host/database_alignment.cpp);
// this entry point serves every other caller (Python binding, bench, multi-GPU pipeline).
// Doubles: same operation order as the reference, libm erf/exp/sqrt, no FMA contraction (x86-64 baseline ISA).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

struct EvRow { int go, ge; double lambda, K, H, a, C, alpha, sigma; };
const EvRow kB62[] = {
    {-1, -1, 0.3176, 0.134, 0.4012, 0.7916, 0.623757, 4.964660, 4.964660},
    {11, 2, 0.297, 0.082, 0.27, 1.1, 0.641766, 12.673800, 12.757600},
    {10, 2, 0.291, 0.075, 0.23, 1.3, 0.649362, 16.474000, 16.602600},
    {9, 2, 0.279, 0.058, 0.19, 1.5, 0.659245, 22.751900, 22.950000},
    {8, 2, 0.264, 0.045, 0.15, 1.8, 0.672692, 35.483800, 35.821300},
    {7, 2, 0.239, 0.027, 0.10, 2.5, 0.702056, 61.238300, 61.886000},
    {6, 2, 0.201, 0.012, 0.061, 3.3, 0.740802, 140.417000, 141.882000},
    {13, 1, 0.292, 0.071, 0.23, 1.2, 0.647715, 19.506300, 19.893100},
    {12, 1, 0.283, 0.059, 0.19, 1.5, 0.656391, 27.856200, 28.469900},
    {11, 1, 0.267, 0.041, 0.14, 1.9, 0.669720, 42.602800, 43.636200},
    {10, 1, 0.243, 0.024, 0.10, 2.5, 0.693267, 83.178700, 85.065600},
    {9, 1, 0.206, 0.010, 0.052, 4.0, 0.731887, 210.333000, 214.842000},
};

struct EvParams { double lambda, K, a, b, alpha, beta, sigma, tau; double length; };

// createEValueParams (sw/evalue.cu:148-220): the table holds rows for BLOSUM_62 (protein) and EDNA_FULL (DNA) only.  BLOSUM_62
// with listed gap penalties takes its row; BLOSUM_62 with other penalties and EVERY other protein matrix fall back to
// row 0 (the ungapped BLOSUM_62 constants, with a warning) -- so row 0 is what any name but "BLOSUM_62" selects here.
// EDNA_FULL selects the DNA formula (calculateEValueDna), which is not on this path: rejected by the callers.
bool matrix_supported(const char* matrix_name) { return !(matrix_name && strcmp(matrix_name, "EDNA_FULL") == 0); }

EvParams make_params(const char* matrix_name, uint64_t db_residues, int go, int ge) {
    int idx = 0;
    if (!matrix_name || strcmp(matrix_name, "BLOSUM_62") == 0)
        for (int i = 0; i < (int)(sizeof(kB62) / sizeof(kB62[0])); ++i)
            if (kB62[i].go == go && kB62[i].ge == ge) { idx = i; break; }
    const EvRow& r = kB62[idx];
    const double G = go + ge;
    EvParams p;
    p.lambda = r.lambda; p.K = r.K; p.a = r.a; p.alpha = r.alpha; p.sigma = r.sigma;
    p.b = 2.0 * G * (kB62[0].a - r.a);
    p.beta = 2.0 * G * (kB62[0].alpha - r.alpha);
    p.tau = 2.0 * G * (kB62[0].alpha - r.sigma);
    p.length = (double)(long long)db_residues;
    return p;
}

inline double evalue(const EvParams& p, int score, int qlen, int tlen) {
    const double y = score, m = qlen, n = tlen;
    const double scale = p.length / (double)tlen;
    const double c0 = 0.39894228040143267793994605993438;
    double lm = m - (p.a * y + p.b);
    double vm = std::max(2.0 * p.alpha / p.lambda, p.alpha * y + p.beta);
    double sm = sqrt(vm);
    double fm = lm / sm;
    double pm = 0.5 + 0.5 * erf(fm);
    double p1 = lm * pm + sm * c0 * exp(-0.5 * fm * fm);
    double ln = n - (p.a * y + p.b);
    double vn = std::max(2.0 * p.alpha / p.lambda, p.alpha * y + p.beta);
    double sn = sqrt(vn);
    double fn = ln / sn;
    double pn = 0.5 + 0.5 * erf(fn);
    double p2 = ln * pn + sn * c0 * exp(-0.5 * fn * fn);
    double c = std::max(2.0 * p.sigma / p.lambda, p.sigma * y + p.tau);
    double area = p1 * p2 + c * pm * pn;
    return area * p.K * exp(-p.lambda * y) * scale;
}

// The factors of evalue() that depend on the score and the query length only, memoised per (query, score): a query's
// survivors share few distinct scores, and the m-side costs one erf, one sqrt and two exp of the formula's 2 + 2 + 3.
// Same operations in the same order as evalue() -> the same doubles bit for bit.
struct EvMemo { int gen; double ab, sm, pm, p1, cpm, ex; };

// ... and the factors that depend on the score alone (one sqrt, one exp), computed once per score value and worker thread:
// the survivors of a query mostly carry distinct scores, so this is the memo that is hit in practice.
struct EvScore { int set; double ab, sm, c, ex; };

inline void evalue_score_side(const EvParams& p, int score, EvScore& o) {
    const double y = score;
    o.ab = p.a * y + p.b;
    o.sm = sqrt(std::max(2.0 * p.alpha / p.lambda, p.alpha * y + p.beta));
    o.c = std::max(2.0 * p.sigma / p.lambda, p.sigma * y + p.tau);
    o.ex = exp(-p.lambda * y);
    o.set = 1;
}

inline void evalue_query_side(const EvScore& s, int qlen, EvMemo& o) {
    const double m = qlen;
    const double c0 = 0.39894228040143267793994605993438;
    o.ab = s.ab;
    double lm = m - o.ab;
    o.sm = s.sm;
    double fm = lm / o.sm;
    o.pm = 0.5 + 0.5 * erf(fm);
    o.p1 = lm * o.pm + o.sm * c0 * exp(-0.5 * fm * fm);
    o.cpm = s.c * o.pm;
    o.ex = s.ex;
}

inline double evalue_target_side(const EvParams& p, const EvMemo& o, int tlen) {
    const double n = tlen;
    const double scale = p.length / (double)tlen;
    const double c0 = 0.39894228040143267793994605993438;
    double ln = n - o.ab;
    double fn = ln / o.sm;                       // vn == vm, so sn == sm
    double pn = 0.5 + 0.5 * erf(fn);
    double p2 = ln * pn + o.sm * c0 * exp(-0.5 * fn * fn);
    double area = o.p1 * p2 + o.cpm * pn;
    return area * p.K * o.ex * scale;
}

struct Row { int64_t i; int score; double value; };

__global__ void ev_flag_kernel(EvParams P, const uint32_t* cand_ids, const int64_t* cand_off, int nq, int64_t n, const int32_t* scores,
                               const int64_t* q_off, const int64_t* db_off, uint32_t id_base, double limit, uint8_t* flags,
                               uint32_t* qidx, int32_t* tlens) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nq;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cand_off[mid] <= i) lo = mid; else hi = mid; }
    const int qlen = (int)(q_off[lo + 1] - q_off[lo]);
    const uint32_t t = cand_ids[i] - id_base;
    const int tlen = (int)(db_off[t + 1] - db_off[t]);
    const double y = scores[i], m = qlen, nn = tlen;
    const double c0 = 0.39894228040143267793994605993438;
    const double lm = m - (P.a * y + P.b);
    const double vm = fmax(2.0 * P.alpha / P.lambda, P.alpha * y + P.beta);
    const double sm = sqrt(vm), fm = lm / sm;
    const double pm = 0.5 + 0.5 * erf(fm);
    const double p1 = lm * pm + sm * c0 * exp(-0.5 * fm * fm);
    const double ln = nn - (P.a * y + P.b);
    const double fn = ln / sm;
    const double pn = 0.5 + 0.5 * erf(fn);
    const double p2 = ln * pn + sm * c0 * exp(-0.5 * fn * fn);
    const double c = fmax(2.0 * P.sigma / P.lambda, P.sigma * y + P.tau);
    const double e = (p1 * p2 + c * pm * pn) * P.K * exp(-P.lambda * y) * (P.length / nn);
    flags[i] = (e <= limit) ? 1 : 0;
    qidx[i] = (uint32_t)lo;
    tlens[i] = tlen;
}

__global__ void ev_gather_kernel(const uint32_t* sel, const uint32_t* n_sel, const uint32_t* cand_ids, const int32_t* scores,
                                 const uint32_t* qidx, const int32_t* tlens, uint32_t* out_q, uint32_t* out_id, int32_t* out_score,
                                 int32_t* out_tlen) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_sel) return;
    const uint32_t s = sel[i];
    out_q[i] = qidx[s]; out_id[i] = cand_ids[s]; out_score[i] = scores[s]; out_tlen[i] = tlens[s];
}

}  // namespace

extern "C" int s4g_select_hits(s4g_ctx* ctx, int32_t nq, const int32_t* query_lens, const uint32_t* cand_ids,
                               const int64_t* cand_offsets, const int32_t* cand_scores, const int32_t* cand_lens,
                               const char* const* cand_names, const char* matrix_name, uint64_t db_residues, int gap_open,
                               int gap_extend, double max_evalue, int max_alignments, int n_threads, uint32_t* out_q,
                               uint32_t* out_t, int32_t* out_score, double* out_evalue, int64_t* out_offsets) {
    return s4g_select_hits_indexed(ctx, nq, query_lens, cand_ids, cand_offsets, cand_scores, cand_lens, cand_names, matrix_name, db_residues, gap_open,
                                   gap_extend, max_evalue, max_alignments, n_threads, out_q, out_t, out_score, out_evalue, out_offsets, nullptr);
}

// s4g_select_hits + (optional) the position of every kept hit in the candidate arrays (out_index, capacity like out_q)
int s4g_select_hits_indexed(s4g_ctx* ctx, int32_t nq, const int32_t* query_lens, const uint32_t* cand_ids,
                            const int64_t* cand_offsets, const int32_t* cand_scores, const int32_t* cand_lens,
                            const char* const* cand_names, const char* matrix_name, uint64_t db_residues, int gap_open,
                            int gap_extend, double max_evalue, int max_alignments, int n_threads, uint32_t* out_q,
                            uint32_t* out_t, int32_t* out_score, double* out_evalue, int64_t* out_offsets, uint32_t* out_index) {
    if (nq < 0 || !query_lens || !cand_offsets || !out_offsets || max_alignments < 0) return S4G_ERR_ARG;
    if (!matrix_supported(matrix_name)) { s4g_set_error(ctx, "s4g_select_hits: %s selects the DNA E-value formula, which this path does not provide", matrix_name); return S4G_ERR_ARG; }
    if (cand_offsets[nq] > 0 && (!cand_ids || !cand_scores || !cand_lens || !out_q || !out_t || !out_score || !out_evalue)) return S4G_ERR_ARG;
    const EvParams P = make_params(matrix_name, db_residues, gap_open, gap_extend);
    if (n_threads <= 0) n_threads = std::min(32, (int)std::thread::hardware_concurrency());   // more threads cost more to start than they save
    if (n_threads < 1) n_threads = 1;
    if (n_threads > nq) n_threads = nq > 0 ? nq : 1;
    std::vector<int32_t> kept(nq, 0);
    auto work = [&](int tid) {
        std::vector<Row> rows;
        std::vector<EvMemo> memo;
        std::vector<EvScore> by_score;
        for (int q = tid; q < nq; q += n_threads) {
            const int64_t b = cand_offsets[q], e = cand_offsets[q + 1];
            rows.resize(e - b);
            int pass = 0;
            for (int64_t i = b; i < e; ++i) {
                const int sc = cand_scores[i];
                double v;
                if (sc >= 0 && sc < (1 << 20)) {
                    if ((size_t)sc >= memo.size()) { memo.resize((size_t)sc + 256, EvMemo{-1, 0, 0, 0, 0, 0, 0}); by_score.resize(memo.size(), EvScore{0, 0, 0, 0, 0}); }
                    EvMemo& mm = memo[sc];
                    if (mm.gen != q) {
                        EvScore& ss = by_score[sc];
                        if (!ss.set) evalue_score_side(P, sc, ss);
                        evalue_query_side(ss, query_lens[q], mm);
                        mm.gen = q;
                    }
                    v = evalue_target_side(P, mm, cand_lens[i]);
                } else {
                    v = evalue(P, sc, query_lens[q], cand_lens[i]);
                }
                rows[i - b] = {i, sc, v};
                if (v <= max_evalue) ++pass;
            }
            const int k = std::min<int64_t>(std::min(pass, max_alignments), e - b);
            auto less = [&](const Row& x, const Row& y) {
                if (x.value == y.value) {
                    if (x.score == y.score) {
                        if (cand_names) { int c = strcmp(cand_names[x.i], cand_names[y.i]); if (c != 0) return c < 0; }
                        return cand_ids[x.i] < cand_ids[y.i];
                    }
                    return x.score > y.score;
                }
                return x.value < y.value;
            };
            std::partial_sort(rows.begin(), rows.begin() + k, rows.end(), less);
            uint32_t* oq = out_q + (size_t)q * max_alignments; uint32_t* ot = out_t + (size_t)q * max_alignments;
            int32_t* os = out_score + (size_t)q * max_alignments; double* oe = out_evalue + (size_t)q * max_alignments;
            for (int j = 0; j < k; ++j) { oq[j] = (uint32_t)q; ot[j] = cand_ids[rows[j].i]; os[j] = rows[j].score; oe[j] = rows[j].value; }
            if (out_index) for (int j = 0; j < k; ++j) out_index[(size_t)q * max_alignments + j] = (uint32_t)rows[j].i;
            kept[q] = k;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    // compact the per-query blocks (stride max_alignments) into contiguous output
    int64_t w = 0;
    out_offsets[0] = 0;
    for (int q = 0; q < nq; ++q) {
        const size_t src = (size_t)q * max_alignments;
        if ((size_t)w != src)
            for (int j = 0; j < kept[q]; ++j) {
                out_q[w + j] = out_q[src + j]; out_t[w + j] = out_t[src + j]; out_score[w + j] = out_score[src + j]; out_evalue[w + j] = out_evalue[src + j];
                if (out_index) out_index[w + j] = out_index[src + j];
            }
        w += kept[q];
        out_offsets[q + 1] = w;
    }
    return S4G_OK;
}

extern "C" int s4g_merge_candidates_host(int n_ranks, int n_queries, int max_candidates, const uint32_t* const* ids,
                                         const float* const* scores, const uint32_t* const* counts, int n_threads,
                                         uint32_t* out_ids, uint32_t* out_counts) {
    if (n_ranks < 1 || n_queries < 0 || max_candidates < 1 || !ids || !scores || !counts || !out_ids || !out_counts) return S4G_ERR_ARG;
    for (int r = 0; r < n_ranks; ++r) if (!ids[r] || !scores[r] || !counts[r]) return S4G_ERR_ARG;
    if (n_threads <= 0) n_threads = std::min(32, (int)std::thread::hardware_concurrency());
    n_threads = std::max(1, std::min(n_threads, std::max(n_queries, 1)));
    const size_t row = (size_t)max_candidates;
    auto work = [&](int tid) {
        std::vector<unsigned long long> keys;
        for (int q = tid; q < n_queries; q += n_threads) {
            keys.clear();
            for (int r = 0; r < n_ranks; ++r) {
                const uint32_t* id = ids[r] + (size_t)q * row;
                const float* sc = scores[r] + (size_t)q * row;
                const uint32_t c = std::min<uint32_t>(counts[r][q], (uint32_t)max_candidates);
                for (uint32_t j = 0; j < c; ++j) {
                    uint32_t bits;
                    memcpy(&bits, sc + j, 4);
                    keys.push_back(((unsigned long long)(~bits) << 32) | id[j]);      // ascending = (score desc, id asc): the device's cand_key
                }
            }
            const size_t keep = std::min(keys.size(), row);
            if (keep < keys.size()) std::nth_element(keys.begin(), keys.begin() + keep, keys.end());
            uint32_t* out = out_ids + (size_t)q * row;
            for (size_t j = 0; j < keep; ++j) out[j] = (uint32_t)keys[j];
            std::sort(out, out + keep);                                                // database_search.cpp:173-180
            out_counts[q] = (uint32_t)keep;
        }
    };
    if (n_threads == 1) { work(0); return S4G_OK; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    return S4G_OK;
}

extern "C" int s4g_merge_hits(s4g_ctx* ctx, int n_ranks, int32_t nq, int max_alignments, const double* gathered, int64_t rank_stride,
                              const int64_t* gathered_counts, uint32_t own_lo, uint32_t own_hi, int n_threads, uint32_t* out_q,
                              uint32_t* out_t, int32_t* out_score, double* out_evalue, int64_t* out_offsets) {
    (void)ctx;
    if (n_ranks < 1 || nq < 0 || max_alignments < 0 || !gathered || !gathered_counts || !out_offsets) return S4G_ERR_ARG;
    const int M = max_alignments;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 16) n_threads = 16;
    if (n_threads > nq) n_threads = nq > 0 ? nq : 1;
    // start of every (rank, query) run inside the rank's block
    std::vector<int64_t> start((size_t)n_ranks * (nq + 1), 0);
    for (int r = 0; r < n_ranks; ++r) {
        int64_t* st = start.data() + (size_t)r * (nq + 1);
        for (int q = 0; q < nq; ++q) st[q + 1] = st[q] + gathered_counts[(size_t)r * nq + q];
        if (st[nq] > rank_stride) return S4G_ERR_ARG;
    }
    std::vector<int32_t> kept(nq, 0);
    struct G { double value, score, id; };
    auto work = [&](int tid) {
        std::vector<G> rows;
        for (int q = tid; q < nq; q += n_threads) {
            rows.clear();
            for (int r = 0; r < n_ranks; ++r) {
                const int64_t* st = start.data() + (size_t)r * (nq + 1);
                const double* src = gathered + ((size_t)r * rank_stride + st[q]) * 3;
                for (int64_t j = 0; j < st[q + 1] - st[q]; ++j) rows.push_back({src[3 * j], src[3 * j + 1], src[3 * j + 2]});
            }
            const int k = std::min<int>((int)rows.size(), M);
            std::partial_sort(rows.begin(), rows.begin() + k, rows.end(), [](const G& x, const G& y) {
                if (x.value != y.value) return x.value < y.value;
                if (x.score != y.score) return x.score > y.score;
                return x.id < y.id;
            });
            int n = 0;
            const size_t dst = (size_t)q * M;
            for (int j = 0; j < k; ++j) {
                const uint32_t id = (uint32_t)rows[j].id;
                if (id < own_lo || id >= own_hi) continue;
                out_q[dst + n] = (uint32_t)q; out_t[dst + n] = id; out_score[dst + n] = (int32_t)rows[j].score; out_evalue[dst + n] = rows[j].value;
                ++n;
            }
            kept[q] = n;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    int64_t w = 0;
    out_offsets[0] = 0;
    for (int q = 0; q < nq; ++q) {
        const size_t src = (size_t)q * M;
        if ((size_t)w != src)
            for (int j = 0; j < kept[q]; ++j) { out_q[w + j] = out_q[src + j]; out_t[w + j] = out_t[src + j]; out_score[w + j] = out_score[src + j]; out_evalue[w + j] = out_evalue[src + j]; }
        w += kept[q];
        out_offsets[q + 1] = w;
    }
    return S4G_OK;
}

extern "C" int s4g_evalue_screen(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* cand_ids, const int64_t* cand_offsets,
                                 int64_t n_pairs, const int32_t* scores, const char* matrix_name, uint64_t db_residues, int gap_open,
                                 int gap_extend, double max_evalue, uint32_t* out_query, uint32_t* out_id, int32_t* out_score, int32_t* out_tlen,
                                 uint32_t* out_count) {
    if (!ctx || !db || !q || !cand_offsets || !out_count || n_pairs < 0) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (n_pairs == 0) { S4G_CUDA(ctx, cudaMemsetAsync(out_count, 0, 4, st)); return S4G_OK; }
    if (n_pairs >= (1ll << 31)) { s4g_set_error(ctx, "s4g_evalue_screen: more than 2^31 pairs"); return S4G_ERR_CAPACITY; }
    if (!matrix_supported(matrix_name)) { s4g_set_error(ctx, "s4g_evalue_screen: %s selects the DNA E-value formula, which this path does not provide", matrix_name); return S4G_ERR_ARG; }
    const EvParams P = make_params(matrix_name, db_residues, gap_open, gap_extend);
    char* buf = (char*)s4g_scratch(ctx, SLOT_AL_WORK, (size_t)n_pairs * (1 + 4 + 4 + 4) + 256);
    if (!buf) return S4G_ERR_NOMEM;
    uint32_t* qidx = (uint32_t*)buf;
    int32_t* tlens = (int32_t*)(qidx + n_pairs);
    uint32_t* sel = (uint32_t*)(tlens + n_pairs);
    uint8_t* flags = (uint8_t*)(sel + n_pairs);
    const int threads = 256;
    const unsigned blocks = (unsigned)((n_pairs + threads - 1) / threads);
    ev_flag_kernel<<<blocks, threads, 0, st>>>(P, cand_ids, cand_offsets, q->n, n_pairs, scores, q->d_off, db->d_off, db->id_base,
                                              max_evalue * (1.0 + 1e-6), flags, qidx, tlens);
    S4G_CHECK_LAUNCH(ctx);
    size_t tmp = 0;
    cub::CountingInputIterator<uint32_t> it(0);
    cub::DeviceSelect::Flagged(nullptr, tmp, it, flags, sel, out_count, (int)n_pairs, st);
    void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp);
    if (!d_tmp) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cub::DeviceSelect::Flagged(d_tmp, tmp, it, flags, sel, out_count, (int)n_pairs, st));
    ctx->launches += 2;
    ev_gather_kernel<<<blocks, threads, 0, st>>>(sel, out_count, cand_ids, scores, qidx, tlens, out_query, out_id, out_score, out_tlen);
    S4G_CHECK_LAUNCH(ctx);
    return S4G_OK;
}
