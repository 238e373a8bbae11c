// C ABI entry points: context, database shard, query batch, host<->device marshalling.
// The stage kernels live in sw_score.cu / prefilter.cu / align.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cctype>

#include <algorithm>
#include "common.cuh"

static thread_local std::string g_tls_error;

void s4g_set_error(s4g_ctx* ctx, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_tls_error = buf;
    if (ctx) ctx->err = buf;
}

#include <chrono>
static double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void s4g_trace_start(s4g_ctx* ctx) {
    if (!ctx->trace) return;
    cudaStreamSynchronize(ctx->stream);
    ctx->trace_last = wall_ms();
}
void s4g_trace_mark(s4g_ctx* ctx, const char* label) {
    if (!ctx->trace) return;
    cudaStreamSynchronize(ctx->stream);
    const double now = wall_ms();
    for (auto& kv : ctx->trace_acc) if (kv.first == label) { kv.second += now - ctx->trace_last; ctx->trace_last = now; return; }
    ctx->trace_acc.emplace_back(label, now - ctx->trace_last);
    ctx->trace_last = now;
}
void s4g_trace_report(s4g_ctx* ctx, const char* header) {
    if (!ctx->trace) return;
    fprintf(stderr, "[s4g trace] %s:", header);
    for (auto& kv : ctx->trace_acc) fprintf(stderr, " %s=%.3fms", kv.first.c_str(), kv.second);
    fprintf(stderr, "\n");
    ctx->trace_acc.clear();
}

void* s4g_scratch(s4g_ctx* ctx, int slot, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (ctx->slot_bytes[slot] >= bytes) return ctx->slot_ptr[slot];
    if (ctx->slot_ptr[slot]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->slot_ptr[slot]);
        ctx->slot_ptr[slot] = nullptr;
        ctx->slot_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 4 + 256;   // grow with slack so repeated calls settle
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        s4g_set_error(ctx, "cudaMalloc(%zu bytes, scratch slot %d): %s", want, slot, cudaGetErrorString(e));
        return nullptr;
    }
    ctx->slot_ptr[slot] = p;
    ctx->slot_bytes[slot] = want;
    return p;
}

void* s4g_pinned(s4g_ctx* ctx, int slot, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (ctx->pin_bytes[slot] >= bytes) return ctx->pin_ptr[slot];
    if (ctx->pin_ptr[slot]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(ctx->pin_ptr[slot]);
        ctx->pin_ptr[slot] = nullptr;
        ctx->pin_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
        s4g_set_error(ctx, "cudaMallocHost(%zu bytes): %s", want, cudaGetErrorString(e));
        return nullptr;
    }
    ctx->pin_ptr[slot] = p;
    ctx->pin_bytes[slot] = want;
    return p;
}

extern "C" {

int s4g_version(void) { return 100; }

const char* s4g_last_error(const s4g_ctx* ctx) { return ctx ? ctx->err.c_str() : g_tls_error.c_str(); }

int s4g_init(int device, s4g_ctx** out) {
    if (!out) return S4G_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        s4g_set_error(nullptr, "no CUDA device available (%s); sift4g_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return S4G_ERR_CUDA;
    }
    if (device < 0 || device >= n) { s4g_set_error(nullptr, "device %d out of range (0..%d)", device, n - 1); return S4G_ERR_ARG; }
    s4g_ctx* ctx = new s4g_ctx();
    ctx->device = device;
    // every failure below releases the context (s4g_shutdown destroys what was created so far) and leaves its text in the
    // calling thread's error slot, which is what s4g_last_error(NULL) reads
    const int rc = [&]() -> int {
        S4G_CUDA(ctx, cudaSetDevice(device));
        cudaDeviceProp prop;
        S4G_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            s4g_set_error(nullptr, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
            return S4G_ERR_CUDA;
        }
        ctx->sm_count = prop.multiProcessorCount;
        { const char* t = getenv("S4G_TRACE"); ctx->trace = t && t[0] && t[0] != '0'; }
        S4G_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        S4G_CUDA(ctx, cudaEventCreate(&ctx->ev_sw0));
        S4G_CUDA(ctx, cudaEventCreate(&ctx->ev_sw1));
        for (int i = 0; i < 4; ++i) S4G_CUDA(ctx, cudaEventCreate(&ctx->ev_al[i]));
        return S4G_OK;
    }();
    if (rc != S4G_OK) { s4g_shutdown(ctx); return rc; }
    *out = ctx;
    return S4G_OK;
}

int s4g_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void s4g_shutdown(s4g_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < s4g_ctx::kSlots; ++i) if (ctx->slot_ptr[i]) cudaFree(ctx->slot_ptr[i]);
    for (int i = 0; i < s4g_ctx::kPins; ++i) if (ctx->pin_ptr[i]) cudaFreeHost(ctx->pin_ptr[i]);
    if (ctx->ev_sw0) cudaEventDestroy(ctx->ev_sw0);
    if (ctx->ev_sw1) cudaEventDestroy(ctx->ev_sw1);
    for (int i = 0; i < 4; ++i) if (ctx->ev_al[i]) cudaEventDestroy(ctx->ev_al[i]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int s4g_set_stream(s4g_ctx* ctx, void* cuda_stream) {
    if (!ctx) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return S4G_OK;
}

int s4g_sync(s4g_ctx* ctx) {
    if (!ctx) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return S4G_OK;
}

int64_t s4g_launch_count(const s4g_ctx* ctx) { return ctx ? ctx->launches : 0; }
void s4g_launch_count_reset(s4g_ctx* ctx) { if (ctx) ctx->launches = 0; }

// ---- database -------------------------------------------------------------------------------------

static int db_finish(s4g_ctx* ctx, s4g_db* db, const uint8_t* codes, const int64_t* offsets, int where) {
    const int64_t n = db->n;
    // offsets are needed on the host in every case (metadata for the shims and for tiling)
    db->h_off.resize(n + 1);
    if (where == S4G_HOST) memcpy(db->h_off.data(), offsets, sizeof(int64_t) * (n + 1));
    else S4G_CUDA(ctx, cudaMemcpy(db->h_off.data(), offsets, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) {
        int64_t len = db->h_off[i + 1] - db->h_off[i];
        if (len <= 0 || len > 0x7fffff00) { s4g_set_error(ctx, "sequence %lld has invalid length %lld", (long long)i, (long long)len); return S4G_ERR_ARG; }
        if (len > db->max_len) db->max_len = (int32_t)len;
    }
    if (n > 0 && db->h_off[0] != 0) { s4g_set_error(ctx, "offsets[0] must be 0"); return S4G_ERR_ARG; }
    db->residues = n > 0 ? (uint64_t)db->h_off[n] : 0;
    S4G_CUDA(ctx, cudaMalloc(&db->d_codes, db->residues + S4G_DB_TAIL_PAD));
    S4G_CUDA(ctx, cudaMalloc(&db->d_off, sizeof(int64_t) * (n + 1)));
    cudaMemcpyKind kind = where == S4G_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (db->residues) S4G_CUDA(ctx, cudaMemcpyAsync(db->d_codes, codes, db->residues, kind, ctx->stream));
    S4G_CUDA(ctx, cudaMemsetAsync(db->d_codes + db->residues, S4G_PAD_CODE, S4G_DB_TAIL_PAD, ctx->stream));
    S4G_CUDA(ctx, cudaMemcpyAsync(db->d_off, db->h_off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
    S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return S4G_OK;
}

int s4g_db_create(s4g_ctx* ctx, const uint8_t* codes, const int64_t* offsets, int64_t n_seqs, uint32_t id_base,
                  int where, s4g_db** out) {
    if (!ctx || !out || n_seqs < 0 || (n_seqs > 0 && (!codes || !offsets))) return S4G_ERR_ARG;
    *out = nullptr;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    s4g_db* db = new s4g_db();
    db->ctx = ctx; db->n = n_seqs; db->id_base = id_base;
    int64_t zero = 0;
    int rc = db_finish(ctx, db, codes, n_seqs ? offsets : &zero, n_seqs ? where : S4G_HOST);
    if (rc != S4G_OK) { s4g_db_close(db); return rc; }
    if (where == S4G_HOST && db->residues) db->h_codes.assign(codes, codes + db->residues);
    *out = db;
    return S4G_OK;
}

int s4g_db_create_view(s4g_ctx* ctx, s4g_view* view, const int64_t* offsets, int64_t n_seqs, int where, s4g_db** out) {
    if (!ctx || !view || !out || n_seqs <= 0 || !offsets) return S4G_ERR_ARG;
    *out = nullptr;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    s4g_db* db = new s4g_db();
    db->ctx = ctx; db->n = n_seqs; db->id_base = 0;
    db->d_codes = (uint8_t*)s4g_view_ptr(view);
    db->borrowed_codes = true;
    { uint64_t a = 0, b = 0; s4g_view_local_range(view, &a, &b); db->local_lo = (int64_t)a; db->local_hi = (int64_t)b; }
    const int rc = [&]() -> int {
        db->h_off.resize(n_seqs + 1);
        if (where == S4G_HOST) memcpy(db->h_off.data(), offsets, sizeof(int64_t) * (n_seqs + 1));
        else S4G_CUDA(ctx, cudaMemcpy(db->h_off.data(), offsets, sizeof(int64_t) * (n_seqs + 1), cudaMemcpyDeviceToHost));
        if (db->h_off[0] != 0) { s4g_set_error(ctx, "offsets[0] must be 0"); return S4G_ERR_ARG; }
        for (int64_t i = 0; i < n_seqs; ++i) {
            const int64_t len = db->h_off[i + 1] - db->h_off[i];
            if (len <= 0 || len > 0x7fffff00) { s4g_set_error(ctx, "sequence %lld has invalid length %lld", (long long)i, (long long)len); return S4G_ERR_ARG; }
            if (len > db->max_len) db->max_len = (int32_t)len;
        }
        db->residues = (uint64_t)db->h_off[n_seqs];
        if (db->residues + S4G_DB_TAIL_PAD > s4g_view_bytes(view)) {
            s4g_set_error(ctx, "s4g_db_create_view: %llu residues + %d pad bytes do not fit the view's %llu bytes", (unsigned long long)db->residues, S4G_DB_TAIL_PAD, (unsigned long long)s4g_view_bytes(view));
            return S4G_ERR_ARG;
        }
        S4G_CUDA(ctx, cudaMalloc(&db->d_off, sizeof(int64_t) * (n_seqs + 1)));
        S4G_CUDA(ctx, cudaMemcpyAsync(db->d_off, db->h_off.data(), sizeof(int64_t) * (n_seqs + 1), cudaMemcpyHostToDevice, ctx->stream));
        // the readable pad behind the last residue (every rank of a striped database writes the same bytes)
        int r = s4g_view_fill(view, db->residues, S4G_PAD_CODE, S4G_DB_TAIL_PAD);
        if (r != S4G_OK) return r;
        S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return S4G_OK;
    }();
    if (rc != S4G_OK) { s4g_db_close(db); return rc; }
    *out = db;
    return S4G_OK;
}

void s4g_db_close(s4g_db* db) {
    if (!db) return;
    cudaSetDevice(db->ctx->device);
    cudaStreamSynchronize(db->ctx->stream);
    if (db->d_codes && !db->borrowed_codes) cudaFree(db->d_codes);
    if (db->d_off) cudaFree(db->d_off);
    if (db->d_order) cudaFree(db->d_order);
    delete db;
}

int64_t s4g_db_num_seqs(const s4g_db* db) { return db ? db->n : 0; }
uint64_t s4g_db_num_residues(const s4g_db* db) { return db ? db->residues : 0; }
uint32_t s4g_db_id_base(const s4g_db* db) { return db ? db->id_base : 0; }
int64_t s4g_db_total_seqs(const s4g_db* db) { return db ? (db->total_seqs ? db->total_seqs : db->n) : 0; }
uint64_t s4g_db_total_residues(const s4g_db* db) { return db ? (db->total_residues ? db->total_residues : db->residues) : 0; }
const int64_t* s4g_db_host_offsets(const s4g_db* db) { return db ? db->h_off.data() : nullptr; }
const uint8_t* s4g_db_host_codes(const s4g_db* db) { return (db && !db->h_codes.empty()) ? db->h_codes.data() : nullptr; }
const char* s4g_db_name(const s4g_db* db, int64_t i) {
    if (!db || i < 0 || i >= (int64_t)db->names.size()) return nullptr;
    return db->names[i].c_str();
}

// ---- queries ---------------------------------------------------------------------------------------

int s4g_queries_create(s4g_ctx* ctx, const uint8_t* codes, const int64_t* offsets, int32_t nq, int where,
                       s4g_queries** out) {
    if (!ctx || !out || nq <= 0 || !codes || !offsets) return S4G_ERR_ARG;
    *out = nullptr;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    s4g_queries* q = new s4g_queries();
    q->ctx = ctx; q->n = nq;
    q->h_off.resize(nq + 1);
    if (where == S4G_HOST) memcpy(q->h_off.data(), offsets, sizeof(int64_t) * (nq + 1));
    else cudaMemcpy(q->h_off.data(), offsets, sizeof(int64_t) * (nq + 1), cudaMemcpyDeviceToHost);
    for (int32_t i = 0; i < nq; ++i) {
        int64_t len = q->h_off[i + 1] - q->h_off[i];
        if (len <= 0 || len > (1 << 24)) { s4g_set_error(ctx, "query %d has invalid length %lld", i, (long long)len); delete q; return S4G_ERR_ARG; }
        if (len > q->max_len) q->max_len = (int32_t)len;
    }
    int64_t total = q->h_off[nq];
    q->h_codes.resize(total);
    if (where == S4G_HOST) memcpy(q->h_codes.data(), codes, total);
    else cudaMemcpy(q->h_codes.data(), codes, total, cudaMemcpyDeviceToHost);
    for (int64_t i = 0; i < total; ++i) if (q->h_codes[i] >= S4G_NLET) { s4g_set_error(ctx, "query code %d out of range", q->h_codes[i]); delete q; return S4G_ERR_ARG; }
    std::vector<int32_t> order(nq);
    for (int32_t i = 0; i < nq; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return q->h_off[a + 1] - q->h_off[a] > q->h_off[b + 1] - q->h_off[b]; });
    if (cudaMalloc(&q->d_codes, total + S4G_DB_TAIL_PAD) != cudaSuccess || cudaMalloc(&q->d_off, sizeof(int64_t) * (nq + 1)) != cudaSuccess ||
        cudaMalloc(&q->d_len_order, sizeof(int32_t) * nq) != cudaSuccess || cudaMalloc(&q->d_len_rank, sizeof(int32_t) * nq) != cudaSuccess) {
        s4g_set_error(ctx, "cudaMalloc failed for the query batch");
        s4g_queries_free(q);
        return S4G_ERR_NOMEM;
    }
    cudaMemcpyAsync(q->d_codes, q->h_codes.data(), total, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(q->d_codes + total, S4G_PAD_CODE, S4G_DB_TAIL_PAD, ctx->stream);
    cudaMemcpyAsync(q->d_off, q->h_off.data(), sizeof(int64_t) * (nq + 1), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(q->d_len_order, order.data(), sizeof(int32_t) * nq, cudaMemcpyHostToDevice, ctx->stream);
    std::vector<int32_t> rank(nq);
    for (int32_t i = 0; i < nq; ++i) rank[order[i]] = i;
    cudaMemcpyAsync(q->d_len_rank, rank.data(), sizeof(int32_t) * nq, cudaMemcpyHostToDevice, ctx->stream);
    S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = q;
    return S4G_OK;
}

void s4g_queries_free(s4g_queries* q) {
    if (!q) return;
    cudaSetDevice(q->ctx->device);
    cudaStreamSynchronize(q->ctx->stream);
    if (q->d_codes) cudaFree(q->d_codes);
    if (q->d_off) cudaFree(q->d_off);
    if (q->d_len_order) cudaFree(q->d_len_order);
    if (q->d_len_rank) cudaFree(q->d_len_rank);
    delete q;
}

// ---- stage wrappers: host <-> device marshalling -----------------------------------------------------

int s4g_sw_score(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* cand_ids, const int64_t* cand_offsets,
                 int64_t n_pairs, const int32_t* matrix, int gap_open, int gap_extend, int32_t* out_scores, int where) {
    if (!ctx || !db || !q || !cand_offsets || !matrix || n_pairs < 0 || (n_pairs > 0 && (!cand_ids || !out_scores))) return S4G_ERR_ARG;
    if (gap_open < 0 || gap_extend < 0 || gap_open > 4096 || gap_extend > 4096) { s4g_set_error(ctx, "gap penalties out of range"); return S4G_ERR_ARG; }
    // pair positions travel as uint32 (sort values, overflow list) and cub counts items in int
    if (n_pairs >= ((int64_t)1 << 31)) { s4g_set_error(ctx, "s4g_sw_score: %lld pairs in one call (limit 2^31 - 1); split the batch", (long long)n_pairs); return S4G_ERR_CAPACITY; }
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_pairs == 0) return S4G_OK;
    if (where == S4G_DEVICE)
        return s4g_sw_score_device(ctx, db, q, cand_ids, cand_offsets, n_pairs, matrix, gap_open, gap_extend, out_scores);
    uint32_t* d_ids = (uint32_t*)s4g_scratch(ctx, SLOT_IO_A, sizeof(uint32_t) * n_pairs);
    int64_t* d_off = (int64_t*)s4g_scratch(ctx, SLOT_IO_B, sizeof(int64_t) * (q->n + 1));
    int32_t* d_out = (int32_t*)s4g_scratch(ctx, SLOT_IO_C, sizeof(int32_t) * n_pairs);
    if (!d_ids || !d_off || !d_out) return S4G_ERR_NOMEM;
    S4G_CUDA(ctx, cudaMemcpyAsync(d_ids, cand_ids, sizeof(uint32_t) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_off, cand_offsets, sizeof(int64_t) * (q->n + 1), cudaMemcpyHostToDevice, ctx->stream));
    int rc = s4g_sw_score_device(ctx, db, q, d_ids, d_off, n_pairs, matrix, gap_open, gap_extend, d_out);
    if (rc != S4G_OK) return rc;
    S4G_CUDA(ctx, cudaMemcpyAsync(out_scores, d_out, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return S4G_OK;
}

int s4g_prefilter(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int k, int max_candidates, int sorted_by_id,
                  uint32_t* out_ids, float* out_scores, uint32_t* out_counts, int where) {
    if (!ctx || !db || !q || !out_ids || !out_counts) return S4G_ERR_ARG;
    if (k < 3 || k > 5) { s4g_set_error(ctx, "kmer_length possible values = 3,4,5"); return S4G_ERR_ARG; }
    if (max_candidates <= 0) { s4g_set_error(ctx, "invalid max candidates number"); return S4G_ERR_ARG; }
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    if (where == S4G_DEVICE)
        return s4g_prefilter_device(ctx, db, q, k, max_candidates, sorted_by_id, out_ids, out_scores, out_counts);
    size_t cells = (size_t)q->n * (size_t)max_candidates;
    uint32_t* d_ids = (uint32_t*)s4g_scratch(ctx, SLOT_IO_A, sizeof(uint32_t) * cells);
    float* d_sc = (float*)s4g_scratch(ctx, SLOT_IO_B, sizeof(float) * cells);
    uint32_t* d_cnt = (uint32_t*)s4g_scratch(ctx, SLOT_IO_C, sizeof(uint32_t) * q->n);
    if (!d_ids || !d_sc || !d_cnt) return S4G_ERR_NOMEM;
    int rc = s4g_prefilter_device(ctx, db, q, k, max_candidates, sorted_by_id, d_ids, d_sc, d_cnt);
    if (rc != S4G_OK) return rc;
    S4G_CUDA(ctx, cudaMemcpyAsync(out_ids, d_ids, sizeof(uint32_t) * cells, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_scores) S4G_CUDA(ctx, cudaMemcpyAsync(out_scores, d_sc, sizeof(float) * cells, cudaMemcpyDeviceToHost, ctx->stream));
    S4G_CUDA(ctx, cudaMemcpyAsync(out_counts, d_cnt, sizeof(uint32_t) * q->n, cudaMemcpyDeviceToHost, ctx->stream));
    S4G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return S4G_OK;
}

int s4g_last_align_profile(s4g_ctx* ctx, float* ms3, uint64_t* cells3) {
    if (!ctx || !ms3 || !cells3) return S4G_ERR_ARG;
    if (!ctx->al_timed) { s4g_set_error(ctx, "no s4g_sw_align call has been timed yet"); return S4G_ERR_ARG; }
    S4G_CUDA(ctx, cudaEventSynchronize(ctx->ev_al[3]));
    for (int i = 0; i < 3; ++i) { S4G_CUDA(ctx, cudaEventElapsedTime(ms3 + i, ctx->ev_al[i], ctx->ev_al[i + 1])); cells3[i] = ctx->al_cells[i]; }
    return S4G_OK;
}

int s4g_last_sw_kernel_ms(s4g_ctx* ctx, float* ms) {
    if (!ctx || !ms) return S4G_ERR_ARG;
    if (!ctx->sw_timed) { s4g_set_error(ctx, "no s4g_sw_score call has been timed yet"); return S4G_ERR_ARG; }
    S4G_CUDA(ctx, cudaEventSynchronize(ctx->ev_sw1));
    S4G_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_sw0, ctx->ev_sw1));
    return S4G_OK;
}

}  // extern "C"
