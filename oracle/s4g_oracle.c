/* TEST INFRASTRUCTURE -- not product code.  See s4g_oracle.h for the contract and parity status.
 *
 * Plain-C restatement of the reference algorithms on the SIFT4G database-search hot path.
 * Written for clarity, not speed: scalar loops, O(n*m) memory where that is simplest.
 */
#include "s4g_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NLET 26

/* ------------------------------------------------------------------------------------------ */
/* BLOSUM62 over A..Z.  Built from the standard NCBI 23-letter table (order below); J, O, U are
 * the reference's extra rows: -4 against everything, +1 among themselves
 * (sw/constants.c:97,102,108).  tests/test_oracle_vs_reference.py compares all 676 entries with
 * the table the reference's scorerCreateMatrix("BLOSUM_62") returns. */
static const char NCBI_ORDER[] = "ARNDCQEGHILKMFPSTWYVBZX";
static const signed char NCBI_B62[23][23] = {
    { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0,-2,-1, 0},
    {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1, 0,-1},
    {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3, 3, 0,-1},
    {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3, 4, 1,-1},
    { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-3,-2},
    {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2, 0, 3,-1},
    {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1},
    { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,-2,-1},
    {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3, 0, 0,-1},
    {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-3,-3,-1},
    {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-4,-3,-1},
    {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2, 0, 1,-1},
    {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-3,-1,-1},
    {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-3,-3,-1},
    {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,-1,-2},
    { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0, 0, 0},
    { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0,-1,-1, 0},
    {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-4,-3,-2},
    {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-3,-2,-1},
    { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-3,-2,-1},
    {-2,-1, 3, 4,-3, 0, 1,-1, 0,-3,-4, 0,-3,-3,-2, 0,-1,-4,-3,-3, 4, 1,-1},
    {-1, 0, 0, 1,-3, 3, 4,-2, 0,-3,-3, 1,-1,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1},
    { 0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1,-1,-1,-1},
};

void s4g_oracle_blosum62(int32_t* out) {
    int pos[NLET];
    for (int a = 0; a < NLET; ++a) pos[a] = -1;
    for (int i = 0; i < 23; ++i) pos[NCBI_ORDER[i] - 'A'] = i;
    for (int a = 0; a < NLET; ++a)
        for (int b = 0; b < NLET; ++b) {
            int v;
            if (pos[a] >= 0 && pos[b] >= 0) v = NCBI_B62[pos[a]][pos[b]];
            else if (pos[a] < 0 && pos[b] < 0) v = 1;
            else v = -4;
            out[a * NLET + b] = v;
        }
}

int64_t s4g_oracle_encode(const char* str, int64_t n, uint8_t* out) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        unsigned char c = (unsigned char)str[i];
        if (c >= 'A' && c <= 'Z') out[m++] = (uint8_t)(c - 'A');
        else if (c >= 'a' && c <= 'z') out[m++] = (uint8_t)(c - 'a');
    }
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* Prefilter */

int32_t s4g_oracle_lis(const int32_t* src, int32_t n) {
    /* patience sorting, strictly increasing (help[mid] < x moves right): database_search.cpp:268 */
    int32_t* tail = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t best = 0;
    for (int32_t i = 0; i <= n; ++i) tail[i] = INT32_MAX;
    tail[0] = INT32_MIN;
    for (int32_t i = 0; i < n; ++i) {
        int32_t lo = 0, hi = n;
        while (hi > lo) {
            int32_t mid = (lo + hi) / 2;
            if (tail[mid] < src[i]) lo = mid + 1; else hi = mid;
        }
        tail[lo] = src[i];
        if (lo > best) best = lo;
    }
    free(tail);
    return best;
}

typedef struct { uint32_t kmer, q, pos; } QHit;

static int qhit_cmp(const void* a, const void* b) {
    const QHit* x = (const QHit*)a; const QHit* y = (const QHit*)b;
    if (x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
    if (x->q != y->q) return x->q < y->q ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return 0;
}

typedef struct { float score; uint32_t id; } Cand;

static int cand_cmp(const void* a, const void* b) {
    const Cand* x = (const Cand*)a; const Cand* y = (const Cand*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    return 0;
}

static int cand_id_cmp(const void* a, const void* b) {
    const Cand* x = (const Cand*)a; const Cand* y = (const Cand*)b;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

static uint32_t kmer_mask(int k) { return k >= 6 ? 0x3FFFFFFFu : ((1u << (5 * k)) - 1u); }

uint64_t s4g_oracle_prefilter(const uint8_t* db_codes, const int64_t* db_off, int64_t n_db,
                              const uint8_t* q_codes, const int64_t* q_off, int32_t nq, int32_t k,
                              int32_t max_candidates, uint32_t* out_ids, float* out_scores,
                              uint32_t* out_counts, float* all_scores) {
    const uint32_t mask = kmer_mask(k);
    /* index over all queries: hash.cpp:56-85.  Bucket order = (query asc, position asc). */
    int64_t n_hits = 0;
    for (int32_t q = 0; q < nq; ++q) {
        int64_t len = q_off[q + 1] - q_off[q];
        if (len >= k) n_hits += len - k + 1;
    }
    QHit* idx = (QHit*)malloc(sizeof(QHit) * (size_t)(n_hits > 0 ? n_hits : 1));
    int64_t h = 0;
    for (int32_t q = 0; q < nq; ++q) {
        const uint8_t* s = q_codes + q_off[q];
        int64_t len = q_off[q + 1] - q_off[q];
        if (len < k) continue;
        uint32_t km = 0;
        for (int64_t i = 0; i < len; ++i) {
            km = ((km << 5) | s[i]) & mask;
            if (i >= k - 1) { idx[h].kmer = km; idx[h].q = (uint32_t)q; idx[h].pos = (uint32_t)(i - k + 1); ++h; }
        }
    }
    qsort(idx, (size_t)n_hits, sizeof(QHit), qhit_cmp);

    /* per-query candidate lists */
    Cand** cand = (Cand**)calloc((size_t)nq, sizeof(Cand*));
    int64_t* ncand = (int64_t*)calloc((size_t)nq, sizeof(int64_t));
    int64_t* capcand = (int64_t*)calloc((size_t)nq, sizeof(int64_t));
    /* per-sequence scratch: hits[q] lists */
    int32_t** hl = (int32_t**)calloc((size_t)nq, sizeof(int32_t*));
    int32_t* hn = (int32_t*)calloc((size_t)nq, sizeof(int32_t));
    int32_t* hc = (int32_t*)calloc((size_t)nq, sizeof(int32_t));
    int32_t* touched = (int32_t*)malloc(sizeof(int32_t) * (size_t)nq);
    uint64_t cells = 0;
    if (all_scores) memset(all_scores, 0, sizeof(float) * (size_t)nq * (size_t)n_db);

    for (int64_t d = 0; d < n_db; ++d) {
        const uint8_t* s = db_codes + db_off[d];
        int64_t len = db_off[d + 1] - db_off[d];
        cells += (uint64_t)len;
        if (len < k) continue;
        int32_t nt = 0;
        uint32_t km = 0, prev = 0;
        for (int64_t i = 0; i < len; ++i) {
            km = ((km << 5) | s[i]) & mask;
            if (i < k - 1) continue;
            int64_t j = i - k + 1;
            if (j != 0 && km == prev) { prev = km; continue; }   /* database_search.cpp:212-214 */
            prev = km;
            /* lower bound of km in idx */
            int64_t lo = 0, hi = n_hits;
            while (lo < hi) { int64_t mid = (lo + hi) / 2; if (idx[mid].kmer < km) lo = mid + 1; else hi = mid; }
            for (; lo < n_hits && idx[lo].kmer == km; ++lo) {
                uint32_t q = idx[lo].q;
                if (hn[q] == 0) touched[nt++] = (int32_t)q;
                if (hn[q] == hc[q]) {
                    hc[q] = hc[q] ? hc[q] * 2 : 16;
                    hl[q] = (int32_t*)realloc(hl[q], sizeof(int32_t) * (size_t)hc[q]);
                }
                hl[q][hn[q]++] = (int32_t)idx[lo].pos;
            }
        }
        for (int32_t ti = 0; ti < nt; ++ti) {
            int32_t q = touched[ti];
            float score = (float)s4g_oracle_lis(hl[q], hn[q]) / (float)len;   /* :228-229 */
            hn[q] = 0;
            if (all_scores) all_scores[(size_t)q * (size_t)n_db + (size_t)d] = score;
            if (ncand[q] == capcand[q]) {
                capcand[q] = capcand[q] ? capcand[q] * 2 : 64;
                cand[q] = (Cand*)realloc(cand[q], sizeof(Cand) * (size_t)capcand[q]);
            }
            cand[q][ncand[q]].score = score;
            cand[q][ncand[q]].id = (uint32_t)d;
            ++ncand[q];
        }
    }

    for (int32_t q = 0; q < nq; ++q) {
        int64_t n = ncand[q];
        if (n > 0) qsort(cand[q], (size_t)n, sizeof(Cand), cand_cmp);
        if (n > max_candidates) n = max_candidates;
        if (n > 0) qsort(cand[q], (size_t)n, sizeof(Cand), cand_id_cmp);   /* :179 */
        out_counts[q] = (uint32_t)n;
        for (int64_t i = 0; i < n; ++i) {
            out_ids[(size_t)q * (size_t)max_candidates + (size_t)i] = cand[q][i].id;
            if (out_scores) out_scores[(size_t)q * (size_t)max_candidates + (size_t)i] = cand[q][i].score;
        }
        free(cand[q]);
        free(hl[q]);
    }
    free(cand); free(ncand); free(capcand); free(hl); free(hn); free(hc); free(touched); free(idx);
    return cells;
}

/* ------------------------------------------------------------------------------------------ */
/* Smith-Waterman score (Gotoh, affine: first gap residue costs gap_open, each further gap_extend) */

#define NEG_INF (INT_MIN / 4)

static inline int imax(int a, int b) { return a > b ? a : b; }

int32_t s4g_oracle_sw_score(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                            const int32_t* mat, int32_t go, int32_t ge) {
    int32_t* H = (int32_t*)calloc((size_t)qlen + 1, sizeof(int32_t));   /* previous column */
    int32_t* E = (int32_t*)malloc(sizeof(int32_t) * ((size_t)qlen + 1));
    int32_t best = 0;
    for (int32_t i = 0; i <= qlen; ++i) E[i] = NEG_INF;
    for (int32_t j = 0; j < tlen; ++j) {
        int32_t diag = 0, up_h = 0, up_f = NEG_INF;
        const int32_t tj = t[j];
        for (int32_t i = 0; i < qlen; ++i) {
            int32_t e = imax(H[i] - go, E[i] - ge);           /* from the left  */
            int32_t f = imax(up_h - go, up_f - ge);           /* from above     */
            int32_t h = imax(imax(0, diag + mat[q[i] * NLET + tj]), imax(e, f));
            diag = H[i];
            H[i] = h; E[i] = e;
            up_h = h; up_f = f;
            if (h > best) best = h;
        }
    }
    free(H); free(E);
    return best;
}

/* ------------------------------------------------------------------------------------------ */
/* E-value */

typedef struct { int go, ge; double lambda, K, H, a, C, alpha, sigma; } EvRow;
/* BLOSUM_62 rows of sw/evalue.cu:73-88 (row 0 = ungapped) */
static const EvRow EV_B62[] = {
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

double s4g_oracle_evalue(int32_t score, int32_t qlen, int32_t tlen, int64_t db_len, int32_t go,
                         int32_t ge) {
    int idx = 0;   /* no exact row: the reference falls back to the first row of the matrix */
    for (int i = 0; i < (int)(sizeof(EV_B62) / sizeof(EV_B62[0])); ++i)
        if (EV_B62[i].go == go && EV_B62[i].ge == ge) { idx = i; break; }
    const EvRow* r = &EV_B62[idx];
    const double G = (double)(go + ge);
    const double a_un = EV_B62[0].a, alpha_un = EV_B62[0].alpha;
    const double b = 2.0 * G * (a_un - r->a);
    const double beta = 2.0 * G * (alpha_un - r->alpha);
    const double tau = 2.0 * G * (alpha_un - r->sigma);
    const double inv_sqrt_2pi = 0.39894228040143267793994605993438;
    const double y = score, m = qlen, n = tlen;
    const double scale = (double)db_len / (double)tlen;

    double lm = m - (r->a * y + b);
    double vm = fmax(2.0 * r->alpha / r->lambda, r->alpha * y + beta);
    double sm = sqrt(vm);
    double fm = lm / sm;
    double pm = 0.5 + 0.5 * erf(fm);
    double p1 = lm * pm + sm * inv_sqrt_2pi * exp(-0.5 * fm * fm);

    double ln = n - (r->a * y + b);
    double vn = fmax(2.0 * r->alpha / r->lambda, r->alpha * y + beta);
    double sn = sqrt(vn);
    double fn = ln / sn;
    double pn = 0.5 + 0.5 * erf(fn);
    double p2 = ln * pn + sn * inv_sqrt_2pi * exp(-0.5 * fn * fn);

    double c = fmax(2.0 * r->sigma / r->lambda, r->sigma * y + tau);
    double area = p1 * p2 + c * pm * pn;
    return area * r->K * exp(-r->lambda * y) * scale;
}

typedef struct { int32_t idx, score; double value; const char* name; } SelRow;

static int sel_cmp(const void* a, const void* b) {
    const SelRow* x = (const SelRow*)a; const SelRow* y = (const SelRow*)b;
    if (x->value == y->value) {
        if (x->score == y->score) return strcmp(x->name, y->name);
        return y->score - x->score;
    }
    return x->value < y->value ? -1 : 1;
}

int32_t s4g_oracle_select(const double* values, const int32_t* scores, const char* const* names,
                          int32_t n, double threshold, int32_t max_alignments, int32_t* out_idx) {
    SelRow* rows = (SelRow*)malloc(sizeof(SelRow) * (size_t)(n > 0 ? n : 1));
    int32_t pass = 0;
    for (int32_t i = 0; i < n; ++i) {
        rows[i].idx = i; rows[i].score = scores[i]; rows[i].value = values[i]; rows[i].name = names[i];
        if (values[i] <= threshold) ++pass;
    }
    int32_t k = pass < max_alignments ? pass : max_alignments;
    qsort(rows, (size_t)n, sizeof(SelRow), sel_cmp);
    for (int32_t i = 0; i < k; ++i) out_idx[i] = rows[i].idx;
    free(rows);
    return k;
}

/* ------------------------------------------------------------------------------------------ */
/* SSW-rule endpoints: ssw.c:283-308,491-512 (end), :296,500,827-838 (begin) */

/* One Gotoh sweep, target columns visited in the given order, query rows 0..qlen-1.
 * Finds the first visited column whose maximum is (a) a new strict global maximum [stop==0] or
 * (b) equal to `stop` [stop>0, sweep ends there]; reports that column and the smallest row holding
 * the maximum in it. */
static void ssw_sweep(const uint8_t* q, int32_t qlen, int q_rev_from, const uint8_t* t, int32_t t_first,
                      int32_t t_last, int32_t t_step, const int32_t* mat, int32_t go, int32_t ge,
                      int32_t stop, int32_t* out_max, int32_t* out_col, int32_t* out_row) {
    int32_t* H = (int32_t*)calloc((size_t)qlen + 1, sizeof(int32_t));
    int32_t* E = (int32_t*)calloc((size_t)qlen + 1, sizeof(int32_t));
    int32_t best = 0, bcol = -1, brow = -1;
    for (int32_t j = t_first; j != t_last + t_step; j += t_step) {
        int32_t diag = 0, up_h = 0, up_f = 0, colmax = 0, colrow = -1;
        const int32_t tj = t[j];
        for (int32_t i = 0; i < qlen; ++i) {
            int32_t qi = q_rev_from >= 0 ? q[q_rev_from - i] : q[i];
            /* SSW keeps E and F clamped at 0 (unsigned saturating subtract); equivalent for H */
            int32_t e = imax(imax(H[i] - go, E[i] - ge), 0);
            int32_t f = imax(imax(up_h - go, up_f - ge), 0);
            int32_t h = imax(imax(0, diag + mat[tj * NLET + qi]), imax(e, f));
            diag = H[i];
            H[i] = h; E[i] = e; up_h = h; up_f = f;
            if (h > colmax) { colmax = h; colrow = i; }
        }
        if (stop > 0) {
            if (colmax == stop) { best = colmax; bcol = j; brow = colrow; break; }
            if (colmax > best) best = colmax;
        } else if (colmax > best) { best = colmax; bcol = j; brow = colrow; }
    }
    free(H); free(E);
    *out_max = best; *out_col = bcol; *out_row = brow;
}

void s4g_oracle_ssw_endpoints(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                              const int32_t* mat, int32_t go, int32_t ge, int32_t* score,
                              int32_t* coords) {
    int32_t mx, t_end, q_end;
    ssw_sweep(q, qlen, -1, t, 0, tlen - 1, 1, mat, go, ge, 0, &mx, &t_end, &q_end);
    *score = mx;
    coords[0] = coords[1] = coords[2] = coords[3] = -1;
    if (mx <= 0) return;
    int32_t mx2, t_beg, r_row;
    ssw_sweep(q, q_end + 1, q_end, t, t_end, 0, -1, mat, go, ge, mx, &mx2, &t_beg, &r_row);
    coords[1] = q_end; coords[3] = t_end;
    if (t_beg < 0) return;
    coords[0] = q_end - r_row; coords[2] = t_beg;
}

/* ------------------------------------------------------------------------------------------ */
/* SSW banded traceback: ssw.c:549-727.  The reference keeps three band-relative rows (previous H,
 * E, current H) and a direction array with 3 entries per band cell; its band-edge handling has
 * observable quirks (border zeroing by absolute column, buffers carried between band doublings), so
 * the buffers and index formulas are modelled one to one. */

static inline int slot_of(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x + 1; }
static inline int dslot_of(int w, int i, int j, int p) { int x = i - w; if (x < 0) x = 0; return (j - x) * 3 + p; }

int32_t s4g_oracle_ssw_banded(const uint8_t* read, int32_t readLen, const uint8_t* ref, int32_t refLen,
                              const int32_t* mat, int32_t go, int32_t ge, int32_t score, uint8_t* path,
                              int32_t path_cap) {
    int32_t w = abs(refLen - readLen) + 1;
    int32_t best = 0;
    size_t rowcap = 16;
    int32_t *hb = (int32_t*)calloc(rowcap, 4), *eb = (int32_t*)calloc(rowcap, 4), *hc = (int32_t*)calloc(rowcap, 4);
    int8_t* dir = NULL;
    int32_t width_d = 0;
    do {
        int32_t width = 2 * w + 3;
        width_d = 2 * w + 1;
        if ((size_t)width + 2 > rowcap) {
            size_t nc = (size_t)width * 2 + 8;
            hb = (int32_t*)realloc(hb, nc * 4); eb = (int32_t*)realloc(eb, nc * 4); hc = (int32_t*)realloc(hc, nc * 4);
            memset(hb + rowcap, 0, (nc - rowcap) * 4); memset(eb + rowcap, 0, (nc - rowcap) * 4);
            memset(hc + rowcap, 0, (nc - rowcap) * 4);
            rowcap = nc;
        }
        free(dir);
        dir = (int8_t*)calloc((size_t)width_d * 3 * (size_t)readLen + 16, 1);
        for (int32_t j = 1; j < width - 1; ++j) hb[j] = 0;
        for (int32_t i = 0; i < readLen; ++i) {
            int32_t beg = i - w > 0 ? i - w : 0;
            int32_t end = i + w < refLen - 1 ? i + w : refLen - 1;
            int32_t edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int32_t f = 0, u = 0;
            hb[0] = eb[0] = hb[edge] = eb[edge] = hc[0] = 0;
            int8_t* line = dir + (size_t)width_d * 3 * (size_t)i;
            for (int32_t j = beg; j <= end; ++j) {
                u = slot_of(w, i, j);
                int32_t up = slot_of(w, i - 1, j), left = slot_of(w, i, j - 1), dg = slot_of(w, i - 1, j - 1);
                int32_t open = i == 0 ? -go : hb[up] - go;
                int32_t ext = i == 0 ? -ge : eb[up] - ge;
                eb[u] = open > ext ? open : ext;
                line[dslot_of(w, i, j, 0)] = open > ext ? 3 : 2;
                open = hc[left] - go;
                ext = f - ge;
                f = open > ext ? open : ext;
                line[dslot_of(w, i, j, 1)] = open > ext ? 5 : 4;
                int32_t e1 = eb[u] > 0 ? eb[u] : 0;
                int32_t f1 = f > 0 ? f : 0;
                int32_t gap = e1 > f1 ? e1 : f1;
                int32_t dsc = hb[dg] + mat[ref[j] * NLET + read[i]];
                hc[u] = gap > dsc ? gap : dsc;
                if (hc[u] > best) best = hc[u];
                if (gap <= dsc) line[dslot_of(w, i, j, 2)] = 1;
                else line[dslot_of(w, i, j, 2)] = e1 > f1 ? line[dslot_of(w, i, j, 0)] : line[dslot_of(w, i, j, 1)];
            }
            for (int32_t j = 1; j <= u; ++j) hb[j] = hc[j];
        }
        w *= 2;
    } while (best < score && w < (1 << 28));
    w /= 2;

    /* traceback from the bottom-right corner, state H; stops when row 0 is reached */
    int32_t i = readLen - 1, j = refLen - 1, state = 2, n = 0, rc = 0;
    uint8_t* rev = (uint8_t*)malloc((size_t)readLen + (size_t)refLen + 2);
    while (i > 0) {
        int32_t x = i - w > 0 ? i - w : 0;
        int32_t hi = i + w < refLen - 1 ? i + w : refLen - 1;
        if (j < x || j > hi) { rc = -2; break; }
        int8_t d = dir[(size_t)width_d * 3 * (size_t)i + (size_t)dslot_of(w, i, j, state)];
        if (n >= readLen + refLen) { rc = -2; break; }
        switch (d) {
            case 1: --i; --j; state = 2; rev[n++] = 1; break;
            case 2: --i; state = 0; rev[n++] = 3; break;
            case 3: --i; state = 2; rev[n++] = 3; break;
            case 4: --j; state = 1; rev[n++] = 2; break;
            case 5: --j; state = 2; rev[n++] = 2; break;
            default: rc = -2; break;
        }
        if (rc) break;
    }
    if (!rc) {
        rev[n++] = 1;   /* the remaining cell is closed as one more match: ssw.c:689-706 */
        if (n > path_cap) rc = -1;
        else { for (int32_t a = 0; a < n; ++a) path[a] = rev[n - 1 - a]; rc = n; }
    }
    free(rev); free(dir); free(hb); free(eb); free(hc);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* swAlign-rule path (score > 32767 or SSW not applicable): cpu_module.c:1185-1413 */

static int32_t swalign_rule(const uint8_t* q, int32_t rows, const uint8_t* t, int32_t cols,
                            const int32_t* mat, int32_t go, int32_t ge, int32_t* coords, uint8_t* path,
                            int32_t path_cap, int32_t* out_score) {
    const int32_t SMIN = -1000000000;
    size_t cells = (size_t)rows * (size_t)cols;
    int8_t* mv = (int8_t*)malloc(cells);
    int32_t* hg = (int32_t*)calloc(cells, 4);
    int32_t* vg = (int32_t*)calloc(cells, 4);
    int32_t* hs = (int32_t*)calloc((size_t)cols, 4);
    int32_t* ha = (int32_t*)malloc((size_t)cols * 4);
    for (int32_t c = 0; c < cols; ++c) ha[c] = SMIN;
    int32_t best = 0, er = 0, ec = 0;
    for (int32_t r = 0; r < rows; ++r) {
        int32_t iscr = 0, iaff = SMIN, diag = 0;
        for (int32_t c = 0; c < cols; ++c) {
            size_t k = (size_t)r * (size_t)cols + (size_t)c;
            int32_t mch = mat[q[r] * NLET + t[c]] + diag;
            int32_t ins = imax(iscr - go, iaff - ge);
            hg[k] = (ins == iaff - ge && c > 0) ? hg[k - 1] + 1 : 0;
            int32_t del = imax(hs[c] - go, ha[c] - ge);
            vg[k] = (del == ha[c] - ge && r > 0) ? vg[k - (size_t)cols] + 1 : 0;
            int32_t scr = imax(imax(0, mch), imax(ins, del));
            if (del == scr) mv[k] = 3; else if (ins == scr) mv[k] = 2; else if (mch == scr) mv[k] = 1; else mv[k] = 0;
            if (scr > best) { best = scr; er = r; ec = c; }
            iscr = scr; iaff = ins; diag = hs[c]; hs[c] = scr; ha[c] = del;
        }
    }
    *out_score = best;
    int32_t rc;
    if (best == 0) { coords[0] = coords[1] = coords[2] = coords[3] = 0; rc = 0; }
    else {
        int32_t r = er, c = ec, n = 0;
        uint8_t* rev = (uint8_t*)malloc((size_t)rows + (size_t)cols + 2);
        while (r >= 0 && c >= 0) {
            size_t k = (size_t)r * (size_t)cols + (size_t)c;
            int8_t m = mv[k];
            if (m == 1) { rev[n++] = 1; --r; --c; }
            else if (m == 2) { int32_t g = hg[k]; for (int32_t a = 0; a <= g; ++a) rev[n++] = 2; c -= g + 1; }
            else if (m == 3) { int32_t g = vg[k]; for (int32_t a = 0; a <= g; ++a) rev[n++] = 3; r -= g + 1; }
            else { ++r; ++c; break; }
        }
        if (r == -1 || c == -1) { ++r; ++c; }
        coords[0] = r; coords[1] = er; coords[2] = c; coords[3] = ec;
        if (n > path_cap) rc = -1;
        else { for (int32_t a = 0; a < n; ++a) path[a] = rev[n - 1 - a]; rc = n; }
        free(rev);
    }
    free(mv); free(hg); free(vg); free(hs); free(ha);
    return rc;
}

int32_t s4g_oracle_align(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                         const int32_t* mat, int32_t go, int32_t ge, int32_t score, int32_t* coords,
                         uint8_t* path, int32_t path_cap) {
    int ssw_ok = score <= 32767 && abs(go) <= 127 && abs(ge) <= 127;   /* sse_module.c:181-236 */
    for (int i = 0; i < NLET * NLET && ssw_ok; ++i) if (abs(mat[i]) > 127) ssw_ok = 0;
    if (!ssw_ok) {
        int32_t s2;
        return swalign_rule(q, qlen, t, tlen, mat, go, ge, coords, path, path_cap, &s2);
    }
    int32_t s1;
    s4g_oracle_ssw_endpoints(q, qlen, t, tlen, mat, go, ge, &s1, coords);
    if (s1 <= 0 || coords[0] < 0) return -3;
    return s4g_oracle_ssw_banded(q + coords[0], coords[1] - coords[0] + 1, t + coords[2],
                                 coords[3] - coords[2] + 1, mat, go, ge, s1, path, path_cap);
}

/* ------------------------------------------------------------------------------------------ */
/* SIFT4G's selection of the alignments a prediction is built from (the step right behind the hot path). */

/* alignmentsExtract + aligmentStr, sift4g/src/select_alignments.cpp:127-180,244-299: the hit as a string over the query
 * positions -- 'X' in front of qstart, then one character per path move that consumes a query residue (the aligned target
 * letter for a DIAG move, 'X' for a gap in the target), 'X' behind the end of the path. */
void s4g_oracle_alignment_string(const uint8_t* t, int32_t qlen, int32_t qstart, int32_t tstart, const uint8_t* path,
                                 int32_t path_len, char* out) {
    int32_t j = 0, ti = tstart;
    for (; j < qstart; ++j) out[j] = 'X';
    for (int32_t k = 0; k < path_len; ++k) {
        if (path[k] == 1) { out[j++] = (char)('A' + t[ti]); ++ti; }         /* MOVE_DIAG */
        else if (path[k] == 3) out[j++] = 'X';                              /* MOVE_UP: query residue against a gap */
        else ++ti;                                                          /* MOVE_LEFT: target residue against a gap: no column */
    }
    for (; j < qlen; ++j) out[j] = 'X';
}

static int flt_cmp(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }

/* getMedian, sift4g/src/constants.hpp:77-86: std::sort(&a[0], &a[len - 1]) leaves the last element where it is */
static float sift_median(float* a, int32_t len) {
    if (len > 1) qsort(a, (size_t)(len - 1), sizeof(float), flt_cmp);
    if (len % 2 == 0) return (float)((a[len / 2 - 1] + a[len / 2]) / 2.0);
    return a[len / 2];
}

/* alignmentsSelect, select_alignments.cpp:182-242: strings[i] (qlen characters each) are added one at a time while the
 * median conservation stays above the threshold; returns how many are kept. */
int32_t s4g_oracle_alignments_select(const char* const* strings, int32_t n, int32_t qlen, float threshold) {
    const double kLog_2_20 = 4.321928095;                                   /* constants.hpp:10 */
    int nums[26];
    float* pos_freq = (float*)calloc((size_t)(qlen > 0 ? qlen : 1), sizeof(float));
    float median = (float)kLog_2_20;
    int32_t i;
    for (int k = 0; k < 26; ++k) nums[k] = 0;
    for (i = 1; median > threshold && i <= n; ++i) {
        for (int32_t j = 0; j < qlen; ++j) {
            int valid = 0;
            for (int32_t k = 0; k < i; ++k) {
                const char c = strings[k][j];
                if (c != 'X') { ++valid; nums[c - 'A']++; }
            }
            for (int k = 0; k < 26; ++k)
                if (nums[k] != 0) pos_freq[j] += nums[k] / (float)valid * log2f(nums[k] / (float)valid);
            pos_freq[j] += kLog_2_20;
            for (int k = 0; k < 26; ++k) nums[k] = 0;
        }
        median = sift_median(pos_freq, qlen);
        for (int32_t j = 0; j < qlen; ++j) pos_freq[j] = 0.0f;
    }
    free(pos_freq);
    return i - 1;
}

/* outputDatabaseBlastM8's counts, sw/post_proc.c:962-1003: identities, mismatches and gap openings along the path, with its
 * state machine as written (only an identical pair closes an open gap).  stats = {identity, mismatches, gap openings, length}. */
void s4g_oracle_alignment_stats(const uint8_t* q, const uint8_t* t, int32_t qstart, int32_t tstart, const uint8_t* path,
                                int32_t path_len, int32_t* stats) {
    int identity = 0, mismatches = 0, openings = 0, openq = 0, opent = 0;
    int32_t qi = qstart, ti = tstart;
    for (int32_t k = 0; k < path_len; ++k) {
        /* aligmentStr (post_proc.c, private copy of select_alignments.cpp:244-299): '-' in the query string for MOVE_LEFT,
         * in the target string for MOVE_UP */
        const int qgap = path[k] == 2, tgap = path[k] == 3;
        const int qc = qgap ? -1 : q[qi], tc = tgap ? -1 : t[ti];
        if (!qgap) ++qi;
        if (!tgap) ++ti;
        if (qc == tc) { ++identity; openq = 0; opent = 0; }
        else if (qgap) { if (!openq) ++openings; openq = 1; opent = 0; }
        else if (tgap) { if (!opent) ++openings; openq = 0; opent = 1; }
        else ++mismatches;
    }
    stats[0] = identity; stats[1] = mismatches; stats[2] = openings; stats[3] = path_len;
}
