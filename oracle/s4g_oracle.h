/* TEST INFRASTRUCTURE -- not product code.
 *
 * CPU restatement ("oracle") of the SIFT4G database-search hot path, in plain C.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this; the product library (sift4g_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED BY EXECUTION.  The reference ships no golden vectors; every
 * function here is checked against outputs of the reference itself (oracle/_ref, built by
 * oracle/Makefile from /root/reference) by tests/test_oracle_vs_reference.py in this container, and
 * against fixtures generated from it (tests/golden/, script tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference,
 * "sw/" = vendor/swsharp/swsharp/src/).
 */
#ifndef S4G_ORACLE_H
#define S4G_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sw/constants.c:87-114 (BLOSUM_62_TABLE, 26x26, row-major table[a*26+b]) */
void s4g_oracle_blosum62(int32_t* out676);

/* sw/scorer.c:45-72,313-315; sw/chain.c:87-99: letters -> 0..25 (case folded), everything else dropped.
 * Returns the number of codes written. */
int64_t s4g_oracle_encode(const char* str, int64_t n, uint8_t* out);

/* sift4g/src/hash.cpp:21-90 + database_search.cpp:185-280 (+ merge :132-154, output :170-180).
 * db/q: concatenated codes + offsets[n+1].  Candidate rule: top max_candidates by (score desc,
 * id asc) -- the deterministic member of the reference's tie family; ids are then
 * sorted ascending per query like database_search.cpp:179.
 * out_ids/out_scores: nq * max_candidates (row q holds out_counts[q] entries; scores follow ids).
 * all_scores: optional dense nq * n_db float matrix of LIS/len (0 when no hit), or NULL.
 * Returns total residues (database_search.cpp:125-126,182). */
uint64_t s4g_oracle_prefilter(const uint8_t* db_codes, const int64_t* db_off, int64_t n_db,
                              const uint8_t* q_codes, const int64_t* q_off, int32_t nq, int32_t k,
                              int32_t max_candidates, uint32_t* out_ids, float* out_scores,
                              uint32_t* out_counts, float* all_scores);

/* sift4g/src/database_search.cpp:255-280 */
int32_t s4g_oracle_lis(const int32_t* src, int32_t n);

/* sw/swimd/Swimd.cpp:241-275 (semantics; exact integer after 8->16->32 escalation :412-449) */
int32_t s4g_oracle_sw_score(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                            const int32_t* mat676, int32_t gap_open, int32_t gap_extend);

/* sw/evalue.cu:73-88,148-220,436-489 for BLOSUM_62 rows; returns the protein E-value. */
double s4g_oracle_evalue(int32_t score, int32_t qlen, int32_t tlen, int64_t db_len, int32_t gap_open,
                         int32_t gap_extend);

/* sw/database.c:847-869,1043-1059: keep k = min(#{value<=thr}, max_alignments) entries in the order
 * (value asc, score desc, strcmp(name) asc).  names: array of n C strings.  out_idx receives k
 * indices into the inputs.  Returns k. */
int32_t s4g_oracle_select(const double* values, const int32_t* scores, const char* const* names,
                          int32_t n, double threshold, int32_t max_alignments, int32_t* out_idx);

/* Path of one scored pair.
 * score <= 32767 (and |gaps|,|matrix| <= 127): SSW rules -- sw/sse_module.c:62-122,178-265,
 *   sw/ssw/ssw.c:123-345,371-547 (end/begin cells), :549-727 (banded_sw), :771-856 (ssw_align).
 * otherwise: swAlign rules -- sw/cpu_module.c:1185-1413 (without its result-neutral pruning).
 * coords = {qstart,qend,tstart,tend} 0-based inclusive; path bytes: 1=DIAG 2=LEFT 3=UP
 * (sw/alignment.h:43-65).  Returns path length, or <0: -1 capacity too small, -2 traceback left the
 * band (reference behaviour undefined there), -3 internal. */
int32_t s4g_oracle_align(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                         const int32_t* mat676, int32_t gap_open, int32_t gap_extend, int32_t score,
                         int32_t* coords, uint8_t* path, int32_t path_cap);

/* The two halves of the SSW rule, exposed for unit tests. */
void s4g_oracle_ssw_endpoints(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                              const int32_t* mat676, int32_t gap_open, int32_t gap_extend,
                              int32_t* score, int32_t* coords);
int32_t s4g_oracle_ssw_banded(const uint8_t* q, int32_t qlen, const uint8_t* t, int32_t tlen,
                              const int32_t* mat676, int32_t gap_open, int32_t gap_extend,
                              int32_t score, uint8_t* path, int32_t path_cap);

/* alignmentsExtract / alignmentsSelect of sift4g/src/select_alignments.cpp:127-242 (what SIFT4G does with the hot path's
 * alignments before the prediction stage): the hit as a string over the query positions, and the number of leading
 * hits kept under the median-conservation rule (getMedian's all-but-last sort included). */
void s4g_oracle_alignment_string(const uint8_t* t, int32_t qlen, int32_t qstart, int32_t tstart, const uint8_t* path,
                                 int32_t path_len, char* out);
int32_t s4g_oracle_alignments_select(const char* const* strings, int32_t n, int32_t qlen, float threshold);
/* identities, mismatches, gap openings, length of an alignment as the --sub-results table counts them (sw/post_proc.c:962-1003) */
void s4g_oracle_alignment_stats(const uint8_t* q, const uint8_t* t, int32_t qstart, int32_t tstart, const uint8_t* path,
                                int32_t path_len, int32_t* stats);

#ifdef __cplusplus
}
#endif
#endif
