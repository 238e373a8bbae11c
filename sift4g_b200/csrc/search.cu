// The hot path as ONE call with host buffers: what a host program (the CLI's seams, a service, bench.py's `e2e`) binds when
// it does not want to sequence the stages itself.
//
//   s4g_score_screen   stage 2 + E-value pre-screen: scores every (query, candidate) pair on the device and hands back only
//                      the pairs whose E-value can pass max_evalue (a few per cent) -- replaces the per-query
//                      scoreDatabase + eValues sweep of sw/database.c:402-646 for callers that select over several shards
//   s4g_search         prefilter -> scores -> exact host selection (libm doubles, threaded) -> traceback; the candidate
//                      lists, scores and survivors never leave the device unless asked for
//                      (= searchDatabase + alignDatabase, sift4g/src/main.cpp:203-220, for one resident shard)
//
// Results live in pinned host buffers owned by the context (valid until the next call of either function on it).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

enum { PIN_SURV_Q = 0, PIN_SURV_ID, PIN_SURV_SC, PIN_SURV_TL, PIN_MISC, PIN_CAND_IDS, PIN_CAND_OFF, PIN_HIT_Q, PIN_HIT_T, PIN_HIT_S,
       PIN_HIT_E, PIN_HIT_OFF, PIN_COORDS, PIN_PATHS, PIN_PATH_OFF, PIN_COUNTS, PIN_HIT_IDX };

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void sr_compact_rows_kernel(const uint32_t* rows, const uint32_t* counts, const int64_t* off, uint32_t N, uint32_t* out) {
    const int q = blockIdx.x;
    const uint32_t c = counts[q];
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) out[off[q] + i] = rows[(size_t)q * N + i];
}

// algorithmic SW cells of a batch: sum over pairs of len(query) * len(target)
__global__ void sr_cells_kernel(const uint32_t* cand_ids, const int64_t* cand_off, int nq, int64_t n, const int64_t* q_off,
                                const int64_t* db_off, uint32_t id_base, unsigned long long* acc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (i < n) {
        int lo = 0, hi = nq;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cand_off[mid] <= i) lo = mid; else hi = mid; }
        const uint32_t t = cand_ids[i] - id_base;
        v = (unsigned long long)(q_off[lo + 1] - q_off[lo]) * (unsigned long long)(db_off[t + 1] - db_off[t]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(acc, v);
}

// kept hit h sits at position idx[h] of the survivor arrays: its coords, and the length of its path (slot n: 0, for the scan)
__global__ void sr_gather_coords_kernel(const uint32_t* idx, int64_t n, const int32_t* coords, const int64_t* path_off, int32_t* out_coords, int64_t* out_len) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h > n) return;
    if (h == n) { out_len[h] = 0; return; }
    const uint32_t i = idx[h];
    reinterpret_cast<int4*>(out_coords)[h] = reinterpret_cast<const int4*>(coords)[i];
    out_len[h] = path_off[i + 1] - path_off[i];
}

__global__ void sr_gather_paths_kernel(const uint32_t* idx, const uint8_t* paths, const int64_t* path_off, const int64_t* out_off, uint8_t* out) {
    const int64_t h = blockIdx.x;
    const uint32_t i = idx[h];
    const uint8_t* src = paths + path_off[i];
    uint8_t* dst = out + out_off[h];
    const int64_t len = out_off[h + 1] - out_off[h];
    for (int64_t x = threadIdx.x; x < len; x += blockDim.x) dst[x] = src[x];
}

// scores + screen of device-resident candidate lists; survivors to the context's pinned buffers
int score_screen_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* d_ids, const int64_t* d_off, int64_t n_pairs,
                        const int32_t* matrix, const char* matrix_name, uint64_t db_residues, int go, int ge, double max_evalue,
                        s4g_survivors* out) {
    cudaStream_t st = ctx->stream;
    memset(out, 0, sizeof(*out));
    out->n_pairs = n_pairs;
    if (n_pairs == 0) return S4G_OK;
    int32_t* d_scores = (int32_t*)s4g_scratch(ctx, SLOT_SR_SCORES, sizeof(int32_t) * (size_t)n_pairs);
    uint32_t* d_surv = (uint32_t*)s4g_scratch(ctx, SLOT_SR_SURV, sizeof(uint32_t) * 4 * (size_t)n_pairs + 64);
    if (!d_scores || !d_surv) return S4G_ERR_NOMEM;
    unsigned long long* d_acc = (unsigned long long*)(d_surv + 4 * (size_t)n_pairs);          // [0] cells, then the survivor count
    uint32_t* d_count = (uint32_t*)(d_acc + 1);
    S4G_CUDA(ctx, cudaMemsetAsync(d_acc, 0, 16, st));
    int rc = s4g_sw_score_device(ctx, db, q, d_ids, d_off, n_pairs, matrix, go, ge, d_scores);
    if (rc != S4G_OK) return rc;
    sr_cells_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(d_ids, d_off, q->n, n_pairs, q->d_off, db->d_off, db->id_base, d_acc);
    S4G_CHECK_LAUNCH(ctx);
    uint32_t* s_q = d_surv, *s_id = d_surv + n_pairs;
    int32_t* s_sc = (int32_t*)(d_surv + 2 * n_pairs), *s_tl = (int32_t*)(d_surv + 3 * n_pairs);
    rc = s4g_evalue_screen(ctx, db, q, d_ids, d_off, n_pairs, d_scores, matrix_name, db_residues, go, ge, max_evalue, s_q, s_id, s_sc, s_tl, d_count);
    if (rc != S4G_OK) return rc;
    unsigned long long* h_misc = (unsigned long long*)s4g_pinned(ctx, PIN_MISC, 64);
    if (!h_misc) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cudaMemcpyAsync(h_misc, d_acc, 16, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t n_s = (int64_t)(uint32_t)h_misc[1];
    out->sw_cells = h_misc[0];
    out->n = n_s;
    uint32_t* h_q = (uint32_t*)s4g_pinned(ctx, PIN_SURV_Q, 4 * (size_t)n_s), *h_id = (uint32_t*)s4g_pinned(ctx, PIN_SURV_ID, 4 * (size_t)n_s);
    int32_t* h_sc = (int32_t*)s4g_pinned(ctx, PIN_SURV_SC, 4 * (size_t)n_s), *h_tl = (int32_t*)s4g_pinned(ctx, PIN_SURV_TL, 4 * (size_t)n_s);
    if (!h_q || !h_id || !h_sc || !h_tl) return S4G_ERR_NOMEM;
    if (n_s) {
        S4G_CUDA(ctx, cudaMemcpyAsync(h_q, s_q, 4 * n_s, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(h_id, s_id, 4 * n_s, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(h_sc, s_sc, 4 * n_s, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaMemcpyAsync(h_tl, s_tl, 4 * n_s, cudaMemcpyDeviceToHost, st));
        S4G_CUDA(ctx, cudaStreamSynchronize(st));
    }
    out->query = h_q; out->id = h_id; out->score = h_sc; out->tlen = h_tl;
    if (s4g_last_sw_kernel_ms(ctx, &out->sw_kernel_ms) != S4G_OK) out->sw_kernel_ms = 0.f;
    return S4G_OK;
}

}  // namespace

extern "C" int s4g_score_screen(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* cand_ids, const int64_t* cand_offsets,
                                int64_t n_pairs, int where, const int32_t* matrix, const char* matrix_name, uint64_t db_residues,
                                int gap_open, int gap_extend, double max_evalue, s4g_survivors* out) {
    if (!ctx || !db || !q || !cand_offsets || !matrix || !out || n_pairs < 0 || (n_pairs > 0 && !cand_ids)) return S4G_ERR_ARG;
    if (n_pairs >= ((int64_t)1 << 31)) { s4g_set_error(ctx, "s4g_score_screen: %lld pairs in one call (limit 2^31 - 1); split the batch", (long long)n_pairs); return S4G_ERR_CAPACITY; }
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    if (db_residues == 0) db_residues = s4g_db_total_residues(db);
    if (where == S4G_DEVICE || n_pairs == 0)
        return score_screen_device(ctx, db, q, cand_ids, cand_offsets, n_pairs, matrix, matrix_name, db_residues, gap_open, gap_extend, max_evalue, out);
    uint32_t* d_ids = (uint32_t*)s4g_scratch(ctx, SLOT_SR_CAND, sizeof(uint32_t) * (size_t)n_pairs);
    int64_t* d_off = (int64_t*)s4g_scratch(ctx, SLOT_SR_OFF, sizeof(int64_t) * (q->n + 1));
    if (!d_ids || !d_off) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cudaMemcpyAsync(d_ids, cand_ids, sizeof(uint32_t) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_off, cand_offsets, sizeof(int64_t) * (q->n + 1), cudaMemcpyHostToDevice, ctx->stream));
    return score_screen_device(ctx, db, q, d_ids, d_off, n_pairs, matrix, matrix_name, db_residues, gap_open, gap_extend, max_evalue, out);
}

extern "C" int s4g_search(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const s4g_search_params* prm, s4g_search_result* out) {
    if (!ctx || !db || !q || !prm || !out || !prm->matrix) return S4G_ERR_ARG;
    if (prm->kmer_length < 3 || prm->kmer_length > 5) { s4g_set_error(ctx, "kmer_length possible values = 3,4,5"); return S4G_ERR_ARG; }
    if (prm->max_candidates <= 0 || prm->max_alignments < 0) { s4g_set_error(ctx, "invalid max candidates / max alignments number"); return S4G_ERR_ARG; }
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    memset(out, 0, sizeof(*out));
    const int nq = q->n;
    const uint32_t N = (uint32_t)prm->max_candidates;
    const uint64_t db_residues = s4g_db_total_residues(db);
    out->n_queries = nq;
    out->db_residues = db_residues;
    double t0 = now_ms();
    // ---- stage 1: candidate rows (ascending ids), resident
    uint32_t* d_rows = (uint32_t*)s4g_scratch(ctx, SLOT_SR_ROWS, sizeof(uint32_t) * (size_t)nq * N);
    uint32_t* d_cnt = (uint32_t*)s4g_scratch(ctx, SLOT_SR_CNT, sizeof(uint32_t) * (size_t)nq);
    int64_t* d_off = (int64_t*)s4g_scratch(ctx, SLOT_SR_OFF, sizeof(int64_t) * ((size_t)nq + 1));
    uint32_t* h_cnt = (uint32_t*)s4g_pinned(ctx, PIN_COUNTS, sizeof(uint32_t) * (size_t)nq);
    int64_t* h_off = (int64_t*)s4g_pinned(ctx, PIN_CAND_OFF, sizeof(int64_t) * ((size_t)nq + 1));
    if (!d_rows || !d_cnt || !d_off || !h_cnt || !h_off) return S4G_ERR_NOMEM;
    int rc = s4g_prefilter_device(ctx, db, q, prm->kmer_length, prm->max_candidates, /*sorted_by_id=*/1, d_rows, nullptr, d_cnt);
    if (rc != S4G_OK) return rc;
    S4G_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    h_off[0] = 0;
    bool full = true;
    for (int i = 0; i < nq; ++i) { h_off[i + 1] = h_off[i] + h_cnt[i]; full = full && h_cnt[i] == N; }
    const int64_t n_pairs = h_off[nq];
    out->n_pairs = n_pairs;
    out->d2h_bytes += (int64_t)sizeof(uint32_t) * nq;
    if (n_pairs >= ((int64_t)1 << 31)) { s4g_set_error(ctx, "s4g_search: %lld (query, candidate) pairs (limit 2^31 - 1); split the query batch", (long long)n_pairs); return S4G_ERR_CAPACITY; }
    S4G_CUDA(ctx, cudaMemcpyAsync(d_off, h_off, sizeof(int64_t) * (nq + 1), cudaMemcpyHostToDevice, st));
    out->h2d_bytes += (int64_t)sizeof(int64_t) * (nq + 1);
    const uint32_t* d_cand = d_rows;                     // full rows are the ragged list already
    if (!full && n_pairs > 0) {
        uint32_t* d_c = (uint32_t*)s4g_scratch(ctx, SLOT_SR_CAND, sizeof(uint32_t) * (size_t)n_pairs);
        if (!d_c) return S4G_ERR_NOMEM;
        sr_compact_rows_kernel<<<nq, 128, 0, st>>>(d_rows, d_cnt, d_off, N, d_c);
        S4G_CHECK_LAUNCH(ctx);
        d_cand = d_c;
    }
    if (prm->want_candidates) {
        uint32_t* h_ids = (uint32_t*)s4g_pinned(ctx, PIN_CAND_IDS, sizeof(uint32_t) * (size_t)n_pairs);
        if (!h_ids) return S4G_ERR_NOMEM;
        if (n_pairs) S4G_CUDA(ctx, cudaMemcpyAsync(h_ids, d_cand, sizeof(uint32_t) * n_pairs, cudaMemcpyDeviceToHost, st));
        out->cand_ids = h_ids;
        out->cand_offsets = h_off;
        out->d2h_bytes += (int64_t)sizeof(uint32_t) * n_pairs;
    }
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    double t1 = now_ms();
    out->ms_prefilter = (float)(t1 - t0);
    // ---- stage 2 + screen
    s4g_survivors sv;
    rc = score_screen_device(ctx, db, q, d_cand, d_off, n_pairs, prm->matrix, prm->matrix_name, db_residues, prm->gap_open, prm->gap_extend,
                             prm->max_evalue, &sv);
    if (rc != S4G_OK) return rc;
    out->sw_cells = sv.sw_cells;
    out->sw_kernel_ms = sv.sw_kernel_ms;
    out->n_survivors = sv.n;
    out->d2h_bytes += 16 * sv.n + 16;
    double t2 = now_ms();
    out->ms_score = (float)(t2 - t1);
    // ---- exact selection on the host (reference arithmetic and order) and stage 3.
    // The survivors of the device screen are, up to the screen's 1e-6 margin and the top-max_alignments cut, the hits that
    // will be kept -- so when at least 90 % of them are sure to stay (sum_q min(survivors_q, max_alignments)), their
    // traceback starts right away from the device-resident survivor arrays while the host threads compute the exact
    // E-values and the order; the kept hits are then gathered out of the survivors' results (coords, paths) on the device.
    // Otherwise (few hits kept per query out of many survivors) the hits are selected first and only they are traced back.
    const size_t hit_cap = (size_t)nq * (size_t)prm->max_alignments;       // s4g_select_hits works in per-query blocks of max_alignments
    uint32_t* h_hq = (uint32_t*)s4g_pinned(ctx, PIN_HIT_Q, 4 * hit_cap), *h_ht = (uint32_t*)s4g_pinned(ctx, PIN_HIT_T, 4 * hit_cap);
    int32_t* h_hs = (int32_t*)s4g_pinned(ctx, PIN_HIT_S, 4 * hit_cap);
    double* h_he = (double*)s4g_pinned(ctx, PIN_HIT_E, 8 * hit_cap);
    int64_t* h_hoff = (int64_t*)s4g_pinned(ctx, PIN_HIT_OFF, 8 * ((size_t)nq + 1));
    uint32_t* h_idx = (uint32_t*)s4g_pinned(ctx, PIN_HIT_IDX, 4 * hit_cap);
    if (!h_hq || !h_ht || !h_hs || !h_he || !h_hoff || !h_idx) return S4G_ERR_NOMEM;
    std::vector<int64_t> s_off((size_t)nq + 1, 0);
    int64_t sure = 0;
    {
        // survivors come in candidate order: grouped by ascending query
        int64_t i = 0;
        for (int qq = 0; qq < nq; ++qq) {
            while (i < sv.n && sv.query[i] == (uint32_t)qq) ++i;
            s_off[qq + 1] = i;
            sure += std::min<int64_t>(i - s_off[qq], prm->max_alignments);
        }
    }
    std::vector<int32_t> q_lens(nq);
    for (int i = 0; i < nq; ++i) q_lens[i] = (int32_t)(q->h_off[i + 1] - q->h_off[i]);
    std::vector<const char*> names;
    if (!db->names.empty()) {                             // the reference's tie key: strcmp of the target names
        names.resize((size_t)sv.n);
        for (int64_t i = 0; i < sv.n; ++i) names[i] = db->names[sv.id[i] - db->id_base].c_str();
    }
    int rc_sel = S4G_OK;
    double sel_ms = 0.0;
    auto select = [&]() {
        const double a = now_ms();
        rc_sel = s4g_select_hits_indexed(ctx, nq, q_lens.data(), sv.id, s_off.data(), sv.score, sv.tlen, names.empty() ? nullptr : names.data(), prm->matrix_name,
                                         db_residues, prm->gap_open, prm->gap_extend, prm->max_evalue, prm->max_alignments, prm->n_threads, h_hq, h_ht, h_hs, h_he,
                                         h_hoff, h_idx);
        sel_ms = now_ms() - a;
    };
    const char* no_spec = getenv("S4G_NO_SPECULATE");
    const bool speculate = prm->want_alignments && sv.n > 0 && sure * 10 >= sv.n * 9 && !(no_spec && no_spec[0] && no_spec[0] != '0');
    // device-resident survivor arrays (score_screen_device left them in SLOT_SR_SURV: query | id | score | tlen, n_pairs each)
    const uint32_t* d_sq = (const uint32_t*)ctx->slot_ptr[SLOT_SR_SURV];
    const uint32_t* d_sid = d_sq ? d_sq + n_pairs : nullptr;
    const int32_t* d_ssc = d_sq ? (const int32_t*)(d_sq + 2 * n_pairs) : nullptr;
    std::thread sel_thread;
    if (speculate) sel_thread = std::thread(select);
    else select();
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{sel_thread};
    if (!speculate && rc_sel != S4G_OK) return rc_sel;
    double t3 = now_ms();
    if (!speculate) out->ms_select = (float)(t3 - t2);
    int64_t n_hits = speculate ? 0 : h_hoff[nq];
    // ---- stage 3
    if (prm->want_alignments) {
        // path bytes of a hit <= its query + target lengths; the survivors' target residues bound those of the kept hits (a
        // subset) without a random gather from the shard's offset table
        const int64_t n_al = speculate ? sv.n : n_hits;
        int64_t cap = 16;
        for (int64_t i = 0; i < sv.n; ++i) cap += sv.tlen[i];
        if (speculate) for (int64_t i = 0; i < sv.n; ++i) cap += q_lens[sv.query[i]];
        else for (int64_t h = 0; h < n_hits; ++h) cap += q_lens[h_hq[h]];
        int32_t* d_co = (int32_t*)s4g_scratch(ctx, SLOT_SR_AL_COORDS, 16 * (size_t)n_al + 16);
        uint8_t* d_pa = (uint8_t*)s4g_scratch(ctx, SLOT_SR_AL_PATHS, (size_t)cap + 64);
        int64_t* d_po = (int64_t*)s4g_scratch(ctx, SLOT_SR_AL_POFF, 8 * ((size_t)n_al + 1));
        uint32_t* d_hits = (uint32_t*)s4g_scratch(ctx, SLOT_SR_HITS, 4 * 4 * std::max<size_t>(hit_cap, 1) + 64);
        if (!d_co || !d_pa || !d_po || !d_hits) return S4G_ERR_NOMEM;
        const uint32_t *d_aq = d_sq, *d_at = d_sid;
        const int32_t* d_as = d_ssc;
        if (!speculate) {
            if (n_hits) {
                S4G_CUDA(ctx, cudaMemcpyAsync(d_hits, h_hq, 4 * n_hits, cudaMemcpyHostToDevice, st));
                S4G_CUDA(ctx, cudaMemcpyAsync(d_hits + hit_cap, h_ht, 4 * n_hits, cudaMemcpyHostToDevice, st));
                S4G_CUDA(ctx, cudaMemcpyAsync(d_hits + 2 * hit_cap, h_hs, 4 * n_hits, cudaMemcpyHostToDevice, st));
            }
            d_aq = d_hits; d_at = d_hits + hit_cap; d_as = (const int32_t*)(d_hits + 2 * hit_cap);
            out->h2d_bytes += 12 * n_hits;
        }
        rc = s4g_sw_align(ctx, db, q, n_al, d_aq, d_at, d_as, prm->matrix, prm->gap_open, prm->gap_extend, d_co, d_pa, cap, d_po, S4G_DEVICE);
        const int32_t* d_rco = d_co; const uint8_t* d_rpa = d_pa; const int64_t* d_rpo = d_po;
        int64_t n_path = -1;
        if (speculate) {
            sel_thread.join();
            out->ms_select = (float)sel_ms;                     // ran beside the traceback
            if (rc_sel != S4G_OK) return rc_sel;
            if (rc != S4G_OK) return rc;
            n_hits = h_hoff[nq];
            // gather the kept hits (positions h_idx in the survivor arrays) out of the survivors' results
            int32_t* d_co2 = (int32_t*)s4g_scratch(ctx, SLOT_SR_OUT_COORDS, 16 * (size_t)n_hits + 16);
            int64_t* d_po2 = (int64_t*)s4g_scratch(ctx, SLOT_SR_OUT_POFF, 8 * 2 * ((size_t)n_hits + 1));
            if (!d_co2 || !d_po2) return S4G_ERR_NOMEM;
            uint32_t* d_idx = d_hits + 3 * hit_cap;
            if (n_hits) S4G_CUDA(ctx, cudaMemcpyAsync(d_idx, h_idx, 4 * n_hits, cudaMemcpyHostToDevice, st));
            out->h2d_bytes += 4 * n_hits;
            int64_t* d_len = d_po2 + (n_hits + 1);
            sr_gather_coords_kernel<<<(unsigned)((n_hits + 1 + 255) / 256), 256, 0, st>>>(d_idx, n_hits, d_co, d_po, d_co2, d_len);
            S4G_CHECK_LAUNCH(ctx);
            size_t tmp = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_len, d_po2, (int)(n_hits + 1), st);
            void* d_tmp = s4g_scratch(ctx, SLOT_SW_CUB, tmp);
            if (!d_tmp) return S4G_ERR_NOMEM;
            S4G_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_len, d_po2, (int)(n_hits + 1), st));
            ctx->launches += 1;
            uint8_t* d_pa2 = (uint8_t*)s4g_scratch(ctx, SLOT_SR_OUT_PATHS, (size_t)cap + 64);     // a subset of the survivors' paths
            if (!d_pa2) return S4G_ERR_NOMEM;
            if (n_hits) {
                sr_gather_paths_kernel<<<(unsigned)n_hits, 64, 0, st>>>(d_idx, d_pa, d_po, d_po2, d_pa2);
                S4G_CHECK_LAUNCH(ctx);
            }
            d_rco = d_co2; d_rpa = d_pa2; d_rpo = d_po2;
        } else if (rc != S4G_OK) {
            return rc;
        }
        out->n_hits = n_hits;
        if (prm->device_results) {
            S4G_CUDA(ctx, cudaStreamSynchronize(st));
            out->coords = d_rco; out->paths = d_rpa; out->path_offsets = d_rpo;
        } else {
            int32_t* h_co = (int32_t*)s4g_pinned(ctx, PIN_COORDS, 16 * (size_t)n_hits);
            int64_t* h_po = (int64_t*)s4g_pinned(ctx, PIN_PATH_OFF, 8 * ((size_t)n_hits + 1));
            if (!h_co || !h_po) return S4G_ERR_NOMEM;
            if (n_hits) S4G_CUDA(ctx, cudaMemcpyAsync(h_co, d_rco, 16 * n_hits, cudaMemcpyDeviceToHost, st));
            S4G_CUDA(ctx, cudaMemcpyAsync(h_po, d_rpo, 8 * (n_hits + 1), cudaMemcpyDeviceToHost, st));
            S4G_CUDA(ctx, cudaStreamSynchronize(st));
            n_path = h_po[n_hits];
            uint8_t* h_pa = (uint8_t*)s4g_pinned(ctx, PIN_PATHS, (size_t)n_path + 16);
            if (!h_pa) return S4G_ERR_NOMEM;
            if (n_path) S4G_CUDA(ctx, cudaMemcpyAsync(h_pa, d_rpa, (size_t)n_path, cudaMemcpyDeviceToHost, st));
            S4G_CUDA(ctx, cudaStreamSynchronize(st));
            out->coords = h_co; out->paths = h_pa; out->path_offsets = h_po;
            out->d2h_bytes += 16 * n_hits + n_path + 8 * (n_hits + 1);
        }
    }
    out->n_hits = n_hits;
    out->hit_query = h_hq; out->hit_target = h_ht; out->hit_score = h_hs; out->hit_evalue = h_he; out->hit_offsets = h_hoff;
    out->ms_align = (float)(now_ms() - t3);
    return S4G_OK;
}
