// the step right behind the hot path -- alignmentsExtract + alignmentsSelect of
// sift4g/src/select_alignments.cpp:127-242 -- for a whole query batch on the GPU.
//
//   s4g_alignment_strings   every kept hit as a query-anchored string (one letter per query position: the target residue
//                           aligned to it, 'X' where the hit does not cover the position or gaps it): aligmentStr +
//                           alignmentsExtract (:127-180, 244-299)
//   s4g_alignments_select   how many of a query's hits (in their order) the SIFT stage keeps: strings are added one at a time
//                           until the median over the query positions of  log2(20) + sum_a p_a log2 p_a  drops to the
//                           threshold (:182-242, getMedian constants.hpp:77-86)
//
// Bit-identical selection needs the reference's float arithmetic:
//   * p log2 p  is  (n / (float) valid) * log2f(n / (float) valid)  with glibc's log2f.  n <= valid <= hits of a query, so all
//     values the loop can meet come from a table built ON THE HOST with the same expression (libm is not re-implemented on the
//     device); the kernel only adds table entries, in the reference's order (letters A..Z, zero counts skipped), with IEEE float
//     adds, then adds log2(20) in double and rounds once (pos_freq[j] += kLog_2_20 with a double constant);
//   * getMedian sorts all but the LAST element (std::sort(&a[0], &a[len - 1])) and averages the two middle entries of an even
//     length as float sum / 2.0 -- both kept.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kSelThreads = 256;
constexpr double kLog2_20 = 4.321928095;          // sift4g/src/constants.hpp:10

// one warp per hit: walk the path 32 moves at a time
__global__ void __launch_bounds__(256) f3_extract_kernel(const uint8_t* db_codes, const int64_t* db_off, uint32_t id_base, const int64_t* q_off,
                                                         int64_t n_hits, const uint32_t* hit_q, const uint32_t* hit_t, const int32_t* coords,
                                                         const uint8_t* paths, const int64_t* path_off, const int64_t* str_off, uint8_t* out) {
    const int lane = threadIdx.x & 31;
    const int64_t h = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= n_hits) return;
    const unsigned FULL = 0xffffffffu;
    const uint32_t q = hit_q[h];
    const int qlen = (int)(q_off[q + 1] - q_off[q]);
    const uint8_t* t = db_codes + db_off[hit_t[h] - id_base];
    uint8_t* s = out + str_off[h];
    const int qs = coords[4 * h + 0], ts = coords[4 * h + 2];
    const int64_t p0 = path_off[h];
    const int plen = (int)(path_off[h + 1] - p0);
    for (int j = lane; j < qs && j < qlen; j += 32) s[j] = 'X';
    int qi = qs, ti = ts;
    for (int base = 0; base < plen; base += 32) {
        const int k = base + lane;
        const int op = k < plen ? paths[p0 + k] : 0;
        const bool qadv = op == 1 || op == 3, tadv = op == 1 || op == 2;         // DIAG 1, LEFT 2 (target only), UP 3 (query only)
        const unsigned qm = __ballot_sync(FULL, qadv), tm = __ballot_sync(FULL, tadv);
        const unsigned lt = (1u << lane) - 1u;
        const int qpos = qi + __popc(qm & lt), tpos = ti + __popc(tm & lt);
        if (qadv && qpos < qlen) s[qpos] = op == 1 ? (uint8_t)('A' + t[tpos]) : (uint8_t)'X';
        qi += __popc(qm); ti += __popc(tm);
    }
    for (int j = qi + lane; j < qlen; j += 32) s[j] = 'X';
}

__device__ __forceinline__ uint32_t f2key(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// k-th smallest (0-based) of keys[0..n): MSB-first radix select, whole CTA
__device__ uint32_t cta_kth(const uint32_t* keys, int n, int k, uint32_t* hist, uint32_t* sh) {
    uint32_t prefix = 0, mask = 0;
    int rem = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int r = rem, d = 0;
            for (; d < 256; ++d) { if (r < (int)hist[d]) break; r -= (int)hist[d]; }
            sh[0] = (uint32_t)d; sh[1] = (uint32_t)r;
        }
        __syncthreads();
        prefix |= sh[0] << shift;
        mask |= 255u << shift;
        rem = (int)sh[1];
        __syncthreads();
    }
    return prefix;
}

// one CTA per query
__global__ void __launch_bounds__(kSelThreads) f3_select_kernel(const int64_t* q_off, const int64_t* hit_off, const uint8_t* strings, const int64_t* str_off,
                                                                const float* table, int table_dim, float threshold, unsigned short* counts,
                                                                unsigned short* valid, uint32_t* pos_keys, int32_t* out_selected) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sh[2];
    const int q = blockIdx.x;
    const int qlen = (int)(q_off[q + 1] - q_off[q]);
    const int64_t h0 = hit_off[q];
    const int n = (int)(hit_off[q + 1] - h0);
    unsigned short* cnt = counts + q_off[q] * 26;
    unsigned short* val = valid + q_off[q];
    uint32_t* keys = pos_keys + q_off[q];
    for (int64_t i = threadIdx.x; i < (int64_t)qlen * 26; i += blockDim.x) cnt[i] = 0;
    for (int i = threadIdx.x; i < qlen; i += blockDim.x) val[i] = 0;
    __syncthreads();
    float median = (float)kLog2_20;
    int i = 1;
    for (; median > threshold && i <= n; ++i) {
        const uint8_t* s = strings + str_off[h0 + i - 1];
        for (int j = threadIdx.x; j < qlen; j += blockDim.x) {
            const uint8_t c = s[j];
            unsigned short* cj = cnt + (int64_t)j * 26;
            int v = val[j];
            if (c != 'X') { cj[c - 'A']++; val[j] = (unsigned short)++v; }
            float acc = 0.0f;
            const float* row = table + (int64_t)v * table_dim;
#pragma unroll 2
            for (int a = 0; a < 26; ++a) {
                const int na = cj[a];
                if (na != 0) acc = __fadd_rn(acc, row[na]);
            }
            keys[j] = f2key((float)((double)acc + kLog2_20));
        }
        __syncthreads();
        // getMedian: entries 0 .. len-2 sorted, the last one left where it is
        const int len = qlen;
        auto at = [&](int idx) -> float {
            if (idx == len - 1) return key2f(keys[len - 1]);
            return key2f(cta_kth(keys, len - 1, idx, hist, sh));
        };
        if (len % 2 == 0) {
            const float a = at(len / 2 - 1), b = at(len / 2);
            median = (float)((double)__fadd_rn(a, b) / 2.0);
        } else {
            median = at(len / 2);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_selected[q] = i - 1;
}

}  // namespace

extern "C" int s4g_alignment_strings(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_hits, const uint32_t* hit_query, const uint32_t* hit_target,
                                     const int32_t* coords, const uint8_t* paths, const int64_t* path_offsets, uint8_t* out_strings,
                                     int64_t* out_string_offsets) {
    if (!ctx || !db || !q || n_hits < 0 || !out_string_offsets) return S4G_ERR_ARG;
    if (n_hits > 0 && (!hit_query || !hit_target || !coords || !paths || !path_offsets || !out_strings)) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    out_string_offsets[0] = 0;
    for (int64_t h = 0; h < n_hits; ++h) {
        if (hit_query[h] >= (uint32_t)q->n || hit_target[h] < db->id_base || hit_target[h] - db->id_base >= (uint64_t)db->n) {
            s4g_set_error(ctx, "s4g_alignment_strings: hit %lld references an unknown query or target", (long long)h);
            return S4G_ERR_ARG;
        }
        out_string_offsets[h + 1] = out_string_offsets[h] + (q->h_off[hit_query[h] + 1] - q->h_off[hit_query[h]]);
    }
    if (n_hits == 0) return S4G_OK;
    const int64_t total = out_string_offsets[n_hits], n_path = path_offsets[n_hits];
    char* buf = (char*)s4g_scratch(ctx, SLOT_AL_WORK, (size_t)n_hits * (4 + 4 + 16 + 8 + 8) + (size_t)n_path + 256);
    uint8_t* d_str = (uint8_t*)s4g_scratch(ctx, SLOT_AL_OUT, (size_t)total + 64);
    if (!buf || !d_str) return S4G_ERR_NOMEM;
    int64_t* d_poff = (int64_t*)buf;
    int64_t* d_soff = d_poff + n_hits + 1;
    int32_t* d_co = (int32_t*)(d_soff + n_hits + 1);
    uint32_t* d_hq = (uint32_t*)(d_co + 4 * n_hits);
    uint32_t* d_ht = d_hq + n_hits;
    uint8_t* d_paths = (uint8_t*)(d_ht + n_hits);
    S4G_CUDA(ctx, cudaMemcpyAsync(d_poff, path_offsets, 8 * (n_hits + 1), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_soff, out_string_offsets, 8 * (n_hits + 1), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_co, coords, 16 * n_hits, cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_hq, hit_query, 4 * n_hits, cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_ht, hit_target, 4 * n_hits, cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_paths, paths, n_path, cudaMemcpyHostToDevice, st));
    f3_extract_kernel<<<(unsigned)((n_hits + 7) / 8), 256, 0, st>>>(db->d_codes, db->d_off, db->id_base, q->d_off, n_hits, d_hq, d_ht, d_co, d_paths, d_poff,
                                                                    d_soff, d_str);
    S4G_CHECK_LAUNCH(ctx);
    S4G_CUDA(ctx, cudaMemcpyAsync(out_strings, d_str, total, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    return S4G_OK;
}

extern "C" int s4g_alignments_select(s4g_ctx* ctx, int32_t n_queries, const int32_t* query_lens, const int64_t* hit_offsets, const uint8_t* strings,
                                     float threshold, int32_t* out_selected) {
    if (!ctx || n_queries < 0 || !query_lens || !hit_offsets || !out_selected) return S4G_ERR_ARG;
    if (n_queries == 0) return S4G_OK;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n_hits = hit_offsets[n_queries];
    if (n_hits > 0 && !strings) return S4G_ERR_ARG;
    std::vector<int64_t> q_off((size_t)n_queries + 1, 0), s_off((size_t)n_hits + 1, 0);
    int max_hits = 0;
    for (int i = 0; i < n_queries; ++i) {
        if (query_lens[i] <= 0) { s4g_set_error(ctx, "s4g_alignments_select: query %d has no residues", i); return S4G_ERR_ARG; }
        q_off[i + 1] = q_off[i] + query_lens[i];
        const int64_t c = hit_offsets[i + 1] - hit_offsets[i];
        if (c < 0) return S4G_ERR_ARG;
        max_hits = (int)std::max<int64_t>(max_hits, c);
        for (int64_t h = hit_offsets[i]; h < hit_offsets[i + 1]; ++h) s_off[h + 1] = s_off[h] + query_lens[i];
    }
    if (max_hits > 2047) {       // the p log2 p table is quadratic in the hits per query; the CLI's --max-aligns default is 400
        s4g_set_error(ctx, "s4g_alignments_select: %d hits for one query (limit 2047)", max_hits);
        return S4G_ERR_CAPACITY;
    }
    // (n / (float) valid) * log2f(n / (float) valid), the host's libm: row `valid`, column n
    const int dim = max_hits + 1;
    std::vector<float> table((size_t)dim * dim, 0.0f);
    for (int v = 1; v < dim; ++v)
        for (int n = 1; n <= v; ++n) {
            volatile float p = n / (float)v;                    // the reference's operands, each rounded to float
            volatile float l = log2f(n / (float)v);
            table[(size_t)v * dim + n] = p * l;
        }
    const int64_t total_q = q_off[n_queries], total_s = s_off[n_hits];
    char* buf = (char*)s4g_scratch(ctx, SLOT_AL_WORK, (size_t)total_q * (26 * 2 + 2 + 4) + sizeof(float) * table.size() + 8 * ((size_t)n_queries + 1) * 2 +
                                                          8 * ((size_t)n_hits + 1) + 4 * (size_t)n_queries + 1024);
    uint8_t* d_str = (uint8_t*)s4g_scratch(ctx, SLOT_AL_OUT, (size_t)total_s + 64);
    if (!buf || !d_str) return S4G_ERR_NOMEM;
    int64_t* d_qoff = (int64_t*)buf;
    int64_t* d_hoff = d_qoff + n_queries + 1;
    int64_t* d_soff = d_hoff + n_queries + 1;
    float* d_table = (float*)(d_soff + n_hits + 1);
    uint32_t* d_keys = (uint32_t*)(d_table + table.size());
    int32_t* d_sel = (int32_t*)(d_keys + total_q);
    unsigned short* d_cnt = (unsigned short*)(d_sel + n_queries);
    unsigned short* d_val = d_cnt + total_q * 26;
    S4G_CUDA(ctx, cudaMemcpyAsync(d_qoff, q_off.data(), 8 * (n_queries + 1), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_hoff, hit_offsets, 8 * (n_queries + 1), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_soff, s_off.data(), 8 * (n_hits + 1), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_table, table.data(), sizeof(float) * table.size(), cudaMemcpyHostToDevice, st));
    if (total_s) S4G_CUDA(ctx, cudaMemcpyAsync(d_str, strings, total_s, cudaMemcpyHostToDevice, st));
    f3_select_kernel<<<n_queries, kSelThreads, 0, st>>>(d_qoff, d_hoff, d_str, d_soff, d_table, dim, threshold, d_cnt, d_val, d_keys, d_sel);
    S4G_CHECK_LAUNCH(ctx);
    S4G_CUDA(ctx, cudaMemcpyAsync(out_selected, d_sel, 4 * n_queries, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    return S4G_OK;
}

// ---- the --sub-results alignment table fed from the result buffers -------------------------------------------
// outputShotgunDatabase -> outputDatabaseBlastM8/M9 (sw/post_proc.c:253-266,962-1049): per hit identities, mismatches and gap
// openings of the alignment (counted on the GPU from the path and the residues), then one tab-separated line per hit.
namespace {

// one thread per hit: the reference's counting loop (post_proc.c:982-1003) over the path.  Its state machine is kept as
// written: only an identical pair closes an open gap -- a mismatch between two gaps of the same sequence does not.
__global__ void f4_stats_kernel(const uint8_t* db_codes, const int64_t* db_off, uint32_t id_base, const uint8_t* q_codes, const int64_t* q_off,
                                int64_t n_hits, const uint32_t* hit_q, const uint32_t* hit_t, const int32_t* coords, const uint8_t* paths,
                                const int64_t* path_off, int32_t* stats) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hits) return;
    const uint8_t* q = q_codes + q_off[hit_q[h]] + coords[4 * h + 0];
    const uint8_t* t = db_codes + db_off[hit_t[h] - id_base] + coords[4 * h + 2];
    const uint8_t* p = paths + path_off[h];
    const int len = (int)(path_off[h + 1] - path_off[h]);
    int identity = 0, mismatches = 0, openings = 0, open_q = 0, open_t = 0, qi = 0, ti = 0;
    for (int k = 0; k < len; ++k) {
        const int op = p[k];
        if (op == 1) {
            if (q[qi] == t[ti]) { ++identity; open_q = 0; open_t = 0; } else ++mismatches;
            ++qi; ++ti;
        } else if (op == 2) {                       // MOVE_LEFT: gap in the query string
            if (!open_q) ++openings;
            open_q = 1; open_t = 0; ++ti;
        } else {                                    // MOVE_UP: gap in the target string
            if (!open_t) ++openings;
            open_q = 0; open_t = 1; ++qi;
        }
    }
    stats[4 * h + 0] = identity; stats[4 * h + 1] = mismatches; stats[4 * h + 2] = openings; stats[4 * h + 3] = len;
}

}  // namespace

extern "C" int s4g_alignment_stats(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_hits, const uint32_t* hit_query, const uint32_t* hit_target,
                                   const int32_t* coords, const uint8_t* paths, const int64_t* path_offsets, int32_t* out_stats) {
    if (!ctx || !db || !q || n_hits < 0) return S4G_ERR_ARG;
    if (n_hits == 0) return S4G_OK;
    if (!hit_query || !hit_target || !coords || !paths || !path_offsets || !out_stats) return S4G_ERR_ARG;
    S4G_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    for (int64_t h = 0; h < n_hits; ++h)
        if (hit_query[h] >= (uint32_t)q->n || hit_target[h] < db->id_base || hit_target[h] - db->id_base >= (uint64_t)db->n) {
            s4g_set_error(ctx, "s4g_alignment_stats: hit %lld references an unknown query or target", (long long)h);
            return S4G_ERR_ARG;
        }
    const int64_t n_path = path_offsets[n_hits];
    char* buf = (char*)s4g_scratch(ctx, SLOT_AL_WORK, (size_t)n_hits * (8 + 16 + 16 + 4 + 4) + (size_t)n_path + 256);
    if (!buf) return S4G_ERR_NOMEM;
    int64_t* d_poff = (int64_t*)buf;
    int32_t* d_co = (int32_t*)(d_poff + n_hits + 1);
    int32_t* d_stats = d_co + 4 * n_hits;
    uint32_t* d_hq = (uint32_t*)(d_stats + 4 * n_hits);
    uint32_t* d_ht = d_hq + n_hits;
    uint8_t* d_paths = (uint8_t*)(d_ht + n_hits);
    S4G_CUDA(ctx, cudaMemcpyAsync(d_poff, path_offsets, 8 * (n_hits + 1), cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_co, coords, 16 * n_hits, cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_hq, hit_query, 4 * n_hits, cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_ht, hit_target, 4 * n_hits, cudaMemcpyHostToDevice, st));
    S4G_CUDA(ctx, cudaMemcpyAsync(d_paths, paths, n_path, cudaMemcpyHostToDevice, st));
    f4_stats_kernel<<<(unsigned)((n_hits + 127) / 128), 128, 0, st>>>(db->d_codes, db->d_off, db->id_base, q->d_codes, q->d_off, n_hits, d_hq, d_ht, d_co,
                                                                      d_paths, d_poff, d_stats);
    S4G_CHECK_LAUNCH(ctx);
    S4G_CUDA(ctx, cudaMemcpyAsync(out_stats, d_stats, 16 * n_hits, cudaMemcpyDeviceToHost, st));
    S4G_CUDA(ctx, cudaStreamSynchronize(st));
    return S4G_OK;
}

extern "C" int s4g_write_blast_tab(const char* path, int with_header, int32_t n_queries, const int64_t* hit_offsets, const char* const* query_names,
                                   const char* const* target_names, const int32_t* stats, const int32_t* coords, const double* evalues,
                                   const int32_t* scores) {
    if (n_queries < 0 || !hit_offsets) return S4G_ERR_ARG;
    if (hit_offsets[n_queries] > 0 && (!query_names || !target_names || !stats || !coords || !evalues || !scores)) return S4G_ERR_ARG;
    FILE* f = path ? fopen(path, "w") : stdout;
    if (!f) { s4g_set_error(nullptr, "cannot write '%s'", path); return S4G_ERR_IO; }
    auto short_len = [](const char* name) {          // names are cut at the first blank, at most 30 characters (post_proc.c:1008-1016)
        const char* sp = strchr(name, ' ');
        const long n = sp ? sp - name : 30;
        return (int)(n < 30 ? n : 30);
    };
    for (int32_t i = 0; i < n_queries; ++i) {
        if (with_header)
            fprintf(f, "# Fields:\n"
                       "Query id,Subject id,%% identity,alignment length,mismatches,"
                       "gap openings,q. start,q. end,s. start,s. end,e-value,score\n");
        for (int64_t h = hit_offsets[i]; h < hit_offsets[i + 1]; ++h) {
            const int length = stats[4 * h + 3];
            fprintf(f, "%.*s\t", short_len(query_names[i]), query_names[i]);
            fprintf(f, "%.*s\t", short_len(target_names[h]), target_names[h]);
            fprintf(f, "%.2f\t", (100.f * stats[4 * h + 0]) / length);
            fprintf(f, "%d\t%d\t%d\t", length, stats[4 * h + 1], stats[4 * h + 2]);
            fprintf(f, "%d\t%d\t%d\t%d\t", coords[4 * h + 0] + 1, coords[4 * h + 1] + 1, coords[4 * h + 2] + 1, coords[4 * h + 3] + 1);
            const double value = evalues[h];
            if (value > 10e-3 && value < 100) fprintf(f, "%.2f\t", value); else fprintf(f, "%.2e\t", value);
            fprintf(f, "%-d ", scores[h]);
            fprintf(f, "\n");
        }
    }
    const bool ok = !ferror(f);
    if (f != stdout) fclose(f);
    if (!ok) { s4g_set_error(nullptr, "write to '%s' failed", path ? path : "stdout"); return S4G_ERR_IO; }
    return S4G_OK;
}
