/*
 * sift4g_b200 -- C ABI of the B200-native SIFT4G database-search hot path.
 *
 * This is the drop-in boundary: plain C types, caller-allocated outputs, 0 = OK / negative = error,
 * no exceptions and no exit() across it.  The C++ host shims in sift4g_b200/host/ keep the
 * reference's own C++ seams (searchDatabase / alignDatabase) and call only these entry points.
 *
 * Reference interfaces replaced (paths relative to the rvaser/sift4g tree, "sw/" =
 * vendor/swsharp/swsharp/src/):
 *   s4g_prefilter      <- searchDatabase()            sift4g/src/database_search.hpp:17-19
 *                         (threadSearchDatabase        sift4g/src/database_search.cpp:185-253,
 *                          Hash / createKmerVector     sift4g/src/hash.cpp:21-90)
 *   s4g_sw_score       <- scoreDatabasesGpu()         sw/gpu_module.h:280-291  (the reference's own
 *                         GPU plug-in seam, a stub in CPU builds: sw/gpu_module.cu:22-95) and
 *                         scoreDatabaseCpu()           sw/cpu_module.h:143-144
 *   s4g_sw_align       <- alignScoredPair()           sw/align.h (sw/align.c:235-255) ->
 *                         alignScoredPairCpu()         sw/cpu_module.h:67-68
 *   s4g_alignment_strings / s4g_alignments_select  <- alignmentsExtract() / alignmentsSelect()
 *                         sift4g/src/select_alignments.cpp:127-242
 *   s4g_search         <- searchDatabase() + alignDatabase() as called by sift4g/src/main.cpp:203-220
 *   s4g_db_* / s4g_queries_*  <- chainDatabaseGpuCreate/Delete()  sw/gpu_module.h:210-240
 *   s4g_db_open_fasta  <- readFastaChainsPart()       sw/pre_proc.h:76-81 (reader quirks kept)
 *   s4g_db_pack_fasta / s4g_db_open_packed  <- dumpFastaChains()/readFastaChains() ".swsharp" cache
 *                         sw/pre_proc.c:309-374,540-595 (record layout sw/chain.c:225-250)
 *
 * Pointer arguments marked [io] live on the host when `where == S4G_HOST` (the call copies in and
 * out and returns when results are on the host) or on the device of the context when
 * `where == S4G_DEVICE` (the call only enqueues work on the context's stream; use s4g_sync()).
 *
 * Residue codes are the reference's: 'A'..'Z' -> 0..25, case folded, everything else dropped
 * (sw/scorer.c:45-72).  Sequence ids are FASTA-order indices (id_base + local index).
 * Threading: one in-flight call per s4g_ctx.
 */
#ifndef SIFT4G_B200_H
#define SIFT4G_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S4G_OK 0
#define S4G_ERR_CUDA (-1)
#define S4G_ERR_ARG (-2)
#define S4G_ERR_NOMEM (-3)
#define S4G_ERR_IO (-4)
#define S4G_ERR_CAPACITY (-5)
#define S4G_ERR_INTERNAL (-6)

#define S4G_HOST 0
#define S4G_DEVICE 1

typedef struct s4g_ctx s4g_ctx;
typedef struct s4g_db s4g_db;
typedef struct s4g_queries s4g_queries;
typedef struct s4g_stripe s4g_stripe;
typedef struct s4g_view s4g_view;

/* ---- context ---------------------------------------------------------------------------- */
int s4g_version(void);
/* number of CUDA devices visible to the process (0 when there is none or the driver cannot be reached): what the CLI's
 * --cards default ("all available CUDA cards", sift4g/src/main.cpp:329-332) enumerates -- replaces cudaGetCards /
 * cudaCheckCards (sw/cuda_utils.h:73-83) */
int s4g_device_count(void);
int s4g_init(int device, s4g_ctx** out);
void s4g_shutdown(s4g_ctx* ctx);
/* last error text of this context (or of the calling thread when ctx == NULL) */
const char* s4g_last_error(const s4g_ctx* ctx);
/* run all later work of this context on an existing cudaStream_t (e.g. torch's current stream) */
int s4g_set_stream(s4g_ctx* ctx, void* cuda_stream);
int s4g_sync(s4g_ctx* ctx);
/* kernels launched by this context since creation / last reset (bench.py's gpu_launches) */
int64_t s4g_launch_count(const s4g_ctx* ctx);
void s4g_launch_count_reset(s4g_ctx* ctx);

/* ---- database shard (resident in HBM) ---------------------------------------------------- */
/* codes: concatenated residue codes; offsets[n_seqs+1]; id_base = FASTA index of sequence 0 of
 * this shard (multi-GPU: one resident shard per GPU). */
int s4g_db_create(s4g_ctx* ctx, const uint8_t* codes, const int64_t* offsets, int64_t n_seqs,
                  uint32_t id_base, int where, s4g_db** out);
/* Parse a FASTA file like the reference reader (sw/pre_proc.c:437-538) and keep shard
 * `shard` of `n_shards` (contiguous ranges of FASTA order).  Names are kept on the host. */
int s4g_db_open_fasta(s4g_ctx* ctx, const char* path, int shard, int n_shards, s4g_db** out);
/* Packed on-disk database (".s4gdb"): what the FASTA reader produces (names, lengths, codes in FASTA order),
 * written once so that later runs skip the parse -- the role of the reference's ".swsharp" cache
 * (sw/pre_proc.c:309-374,540-595), in the layout the shard has in HBM.  Little endian:
 *   64-byte header {char magic[8] = "S4GDB\0\0\1"; u64 n_seqs, n_residues, names_bytes, codes_pos; u64 reserved[3]}
 *   i64 offsets[n_seqs+1] | i64 name_offsets[n_seqs+1] | names (NUL terminated) | zero pad | at codes_pos:
 *   u8 codes[n_residues] (0..25).
 * s4g_db_pack_fasta needs no GPU and no context (errors: s4g_last_error(NULL)); s4g_db_open_packed reads only the
 * byte ranges of shard `shard` of `n_shards`; s4g_db_open picks the reader by the file's magic. */
int s4g_db_pack_fasta(const char* fasta_path, const char* out_path);
int s4g_db_file_info(const char* packed_path, int64_t* n_seqs, uint64_t* n_residues);
int s4g_db_open_packed(s4g_ctx* ctx, const char* path, int shard, int n_shards, s4g_db** out);
int s4g_db_open(s4g_ctx* ctx, const char* path, int shard, int n_shards, s4g_db** out);
/* One process driving several GPUs (the CLI with --cards): reads / parses the file ONCE and leaves shard d of n_ctx
 * (contiguous FASTA ranges) resident on the device of ctxs[d]; out[d] receives the shard.  With fewer sequences than
 * contexts the trailing shards are empty (n_seqs = 0), which every stage accepts.  n_threads: host threads of the
 * FASTA parse (<= 0: all cores; the CLI passes -t). */
int s4g_db_open_sharded(s4g_ctx* const* ctxs, int n_ctx, const char* path, int n_threads, s4g_db** out);
void s4g_db_close(s4g_db* db);
int64_t s4g_db_num_seqs(const s4g_db* db);
uint64_t s4g_db_num_residues(const s4g_db* db);   /* = `cells` returned by searchDatabase() */
uint32_t s4g_db_id_base(const s4g_db* db);
/* the whole file when the shard was opened from a file (the E-value's database length with n_shards > 1);
 * equal to num_seqs / num_residues otherwise */
int64_t s4g_db_total_seqs(const s4g_db* db);
uint64_t s4g_db_total_residues(const s4g_db* db);
/* host-side metadata (valid until s4g_db_close) */
const int64_t* s4g_db_host_offsets(const s4g_db* db);
const uint8_t* s4g_db_host_codes(const s4g_db* db);      /* NULL for device-created shards */
const char* s4g_db_name(const s4g_db* db, int64_t local_index); /* NULL unless opened from a file */

/* ---- NVLink-striped database (multi-GPU without a merge) ---------------------------------- */
/* Replaces the reference's per-card split of the DATABASE with a host merge (sw/database.c:497-532: one host thread per
 * card scores a slice of the sequences, results joined afterwards; the `cards` argument of alignDatabase,
 * sift4g/src/database_alignment.hpp:18-23).  Here every GPU keeps one stripe of the residue array resident (1/N of the
 * bytes, multiples of s4g_stripe_granularity), maps all N stripes into one contiguous virtual range and runs the whole hot
 * path for ITS OWN queries against the full database, reading the other GPUs' pages over NVLink: no candidate list or hit
 * ever has to be merged.  Ids of a view database are global FASTA indices (id_base 0).
 *   s4g_stripe_create     physical memory on ctx's device, shareable with other processes
 *   s4g_stripe_export_fd  POSIX file descriptor of the stripe (send it to the peers with SCM_RIGHTS; close it afterwards)
 *   s4g_view_open         stripe i is local[i] when non-NULL (same process: the CLI's one-thread-per-GPU host), else it is
 *                         imported from fds[i]; bytes[i] as created.  The view is readable and writable from ctx's device.
 *   s4g_view_write/_fill  device-side copy of device memory / constant fill at byte `at` of the view (stores to a peer's
 *                         pages travel over NVLink); enqueued on ctx's stream.
 *   s4g_db_create_view    database over the view's bytes [0, offsets[n_seqs]) (+ 256 readable pad bytes, written here):
 *                         the codes are NOT copied; close the database before the view. */
uint64_t s4g_stripe_granularity(s4g_ctx* ctx);
int s4g_stripe_create(s4g_ctx* ctx, uint64_t bytes, s4g_stripe** out);
uint64_t s4g_stripe_bytes(const s4g_stripe* stripe);
int s4g_stripe_export_fd(s4g_stripe* stripe, int* out_fd);
void s4g_stripe_free(s4g_stripe* stripe);
int s4g_view_open(s4g_ctx* ctx, int n_stripes, s4g_stripe* const* local /*may be NULL*/, const int* fds /*may be NULL*/,
                  const uint64_t* bytes, s4g_view** out);
void* s4g_view_ptr(const s4g_view* view);
uint64_t s4g_view_bytes(const s4g_view* view);
int s4g_view_write(s4g_view* view, uint64_t at, const uint8_t* src /*[dev]*/, uint64_t bytes);
int s4g_view_fill(s4g_view* view, uint64_t at, int value, uint64_t bytes);
void s4g_view_close(s4g_view* view);
int s4g_db_create_view(s4g_ctx* ctx, s4g_view* view, const int64_t* offsets, int64_t n_seqs, int where, s4g_db** out);

/* ---- query batch -------------------------------------------------------------------------- */
int s4g_queries_create(s4g_ctx* ctx, const uint8_t* codes, const int64_t* offsets, int32_t n_queries,
                       int where, s4g_queries** out);
void s4g_queries_free(s4g_queries* q);

/* ---- stage 1: k-mer candidate prefilter ---------------------------------------------------- */
/* For every query: the max_candidates sequences of this shard with the largest
 * LIS(k-mer hit positions)/len score (float32), ties broken by ascending id (the deterministic
 * member of the reference's tie family).  out_ids / out_scores: n_queries x max_candidates, row q
 * holds out_counts[q] entries.  `sorted_by_id` != 0: rows ordered by ascending id (what
 * searchDatabase() returns); 0: rows ordered best-first (what the multi-GPU merge consumes). */
int s4g_prefilter(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int kmer_length, int max_candidates,
                  int sorted_by_id, uint32_t* out_ids /*[io]*/, float* out_scores /*[io], may be NULL*/,
                  uint32_t* out_counts /*[io]*/, int where);

/* Multi-GPU merge step: given per-rank best-first candidate rows gathered from n_ranks shards
 * (gathered_ids/scores: n_ranks x n_queries x max_candidates, gathered_counts: n_ranks x n_queries)
 * select the global top max_candidates per query by (score desc, id asc) and write them sorted by
 * ascending id.  Device pointers only (it runs right after the NCCL all-gather). */
int s4g_merge_candidates(s4g_ctx* ctx, int n_ranks, int n_queries, int max_candidates,
                         const uint32_t* gathered_ids, const float* gathered_scores,
                         const uint32_t* gathered_counts, uint32_t* out_ids, float* out_scores,
                         uint32_t* out_counts);

/* Host form of the same merge, for callers that drive several GPUs from one process without a collective (the CLI:
 * sift4g_b200/host/database_search.cpp) -- it is the reference's merge of its per-thread lists
 * (sift4g/src/database_search.cpp:132-154,173-180) under the deterministic tie rule.  ids/scores/counts[r] point to shard
 * r's s4g_prefilter output with sorted_by_id = 0 (n_queries x max_candidates rows, counts per query); host pointers, no
 * GPU and no context involved.  out_ids (n_queries x max_candidates): the global best max_candidates of every query by
 * (score desc, id asc), written in ascending id; out_counts[q] of them.  n_threads <= 0: all host cores. */
int s4g_merge_candidates_host(int n_ranks, int n_queries, int max_candidates, const uint32_t* const* ids,
                              const float* const* scores, const uint32_t* const* counts, int n_threads,
                              uint32_t* out_ids, uint32_t* out_counts);

/* Multi-GPU, query-owner protocol (what the pipeline uses; no row is copied or re-sorted).  The merge of
 * per-thread lists in searchDatabase() (sift4g/src/database_search.cpp:132-154) keeps, per query, the rows up
 * to the max_candidates-th best (score desc, id asc) key of the union -- so the shards only have to agree on
 * that key.  After an all-to-all every rank holds the best-first rows of all n_ranks shards for the
 * n_queries queries it owns (same layout as for s4g_merge_candidates); s4g_topn_cutoff writes their cut-off
 * keys ((~score bits) << 32 | id; all ones when the union has at most max_candidates rows).  After an
 * all-gather of the cut-offs, s4g_cutoff_counts tells a shard how long the prefix of each of its own
 * best-first rows is that survives (out_counts[q] <= counts[q]).  Device pointers only, n_ranks <= 32. */
int s4g_topn_cutoff(s4g_ctx* ctx, int n_ranks, int n_queries, int max_candidates,
                    const uint32_t* gathered_ids, const float* gathered_scores,
                    const uint32_t* gathered_counts, uint64_t* out_cutoff);
int s4g_cutoff_counts(s4g_ctx* ctx, int n_queries, int max_candidates, const uint32_t* ids,
                      const float* scores, const uint32_t* counts, const uint64_t* cutoff,
                      uint32_t* out_counts);

/* ---- stage 2: Smith-Waterman affine-gap scores ------------------------------------------------ */
/* cand_ids: concatenated per-query candidate ids (global ids inside this shard's range),
 * cand_offsets[n_queries+1].  matrix: 26x26 int32 row-major (sw/scorer.c:206-208).
 * A gap of length L costs gap_open + (L-1)*gap_extend.  out_scores[i] = exact integer score of
 * cand_ids[i] against its query (sw/swimd/Swimd.cpp:241-275 semantics). */
int s4g_sw_score(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* cand_ids /*[io]*/,
                 const int64_t* cand_offsets /*[io]*/, int64_t n_pairs, const int32_t* matrix /*host*/,
                 int gap_open, int gap_extend, int32_t* out_scores /*[io]*/, int where);

/* ---- stage 3: traceback of kept hits ---------------------------------------------------------- */
/* For pair i: query pair_q[i], target id pair_t[i], known score pair_score[i].
 * out_coords[4*i..] = {qstart,qend,tstart,tend} (0-based inclusive);
 * path bytes 1=DIAG 2=LEFT 3=UP (sw/alignment.h:43-65) are written to
 * out_paths[out_path_offsets[i] .. out_path_offsets[i+1]).  Rules: SSW's for score <= 32767
 * (sw/ssw/ssw.c:549-856), swAlign's otherwise (sw/cpu_module.c:1185-1413).
 * path_capacity: bytes available in out_paths; S4G_ERR_CAPACITY if too small (a safe bound is
 * sum(qlen+tlen)). */
int s4g_sw_align(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_pairs, const uint32_t* pair_q,
                 const uint32_t* pair_t, const int32_t* pair_score, const int32_t* matrix /*host*/,
                 int gap_open, int gap_extend, int32_t* out_coords, uint8_t* out_paths,
                 int64_t path_capacity, int64_t* out_path_offsets, int where);

/* ---- between stage 2 and 3: E-value + hit selection (host side of the boundary) ------------------- */
/* Restates extractThread / eValues / dbAlignmentDataCmp (sw/database.c:821-875,1043-1059,
 * sw/evalue.cu:148-220,436-489) for callers that do not link the reference's host objects: for every
 * query, E-values of its scored candidates in IEEE double with libm erf/exp/sqrt in the reference's
 * operation order, then the k = min(#{E <= max_evalue}, max_alignments) best under
 * (E asc, score desc, name asc).  `names` may be NULL: ties then fall back to ascending target id
 * (equal to name order for zero-padded numeric names).  Runs on `n_threads` host threads (0 = all).
 * Inputs are host pointers.  out_* arrays need capacity n_queries * max_alignments; hits of query q are
 * written contiguously, out_offsets[n_queries+1] delimits them.
 * matrix_name (NULL = "BLOSUM_62") picks the constants like createEValueParams (sw/evalue.cu:148-220): the reference's
 * table (:73-88) has protein rows for BLOSUM_62 only, so BLOSUM_62 with listed gap penalties takes its row and every
 * other case -- unlisted penalties, any other protein matrix -- the reference's fallback, row 0 (:178-191).
 * "EDNA_FULL" would select the DNA formula, which is not on this path: S4G_ERR_ARG. */
int s4g_select_hits(s4g_ctx* ctx, int32_t n_queries, const int32_t* query_lens, const uint32_t* cand_ids,
                    const int64_t* cand_offsets, const int32_t* cand_scores, const int32_t* cand_lens,
                    const char* const* cand_names, const char* matrix_name, uint64_t db_residues, int gap_open,
                    int gap_extend, double max_evalue, int max_alignments, int n_threads, uint32_t* out_q,
                    uint32_t* out_t, int32_t* out_score, double* out_evalue, int64_t* out_offsets);

/* Multi-GPU (one database shard per GPU): the global `max_alignments` best hits of every query out of the per-shard
 * selections.  gathered: n_ranks blocks of rank_stride rows of 3 doubles {E-value, score, target id}; block r holds rank
 * r's s4g_select_hits output (hits grouped by query, in order) and gathered_counts[r * n_queries + q] says how many rows
 * query q has there -- what two all-gathers of every rank's selection give (host pointers).  Order and truncation are
 * those of dbAlignmentsMerge (sw/post_proc.c:299-339,432-456): E asc, score desc, then name/id asc.  Only hits whose
 * target id lies in [own_lo, own_hi) are written (the caller's shard; 0, 0xffffffff for all), in global order;
 * out_* need capacity n_queries * max_alignments. */
int s4g_merge_hits(s4g_ctx* ctx, int n_ranks, int32_t n_queries, int max_alignments, const double* gathered, int64_t rank_stride,
                   const int64_t* gathered_counts, uint32_t own_lo, uint32_t own_hi, int n_threads, uint32_t* out_q, uint32_t* out_t,
                   int32_t* out_score, double* out_evalue, int64_t* out_offsets);

/* GPU pre-screen in front of s4g_select_hits (optional, changes no result): evaluates the E-value of every
 * scored candidate on the device in double (CUDA libm, within a few ulp of the host's) and keeps those with
 * E <= max_evalue * (1 + 1e-6), compacted in candidate order.  Survivors are re-evaluated exactly on the host
 * by s4g_select_hits, so the final hit lists equal a host-only evaluation of every candidate; what the screen
 * saves is the D2H copy and the host libm work for the ~95 % of candidates that cannot pass.
 * cand_ids / cand_offsets / scores: device pointers (as given to / produced by s4g_sw_score with S4G_DEVICE).
 * out_*: device arrays of capacity n_pairs; out_count: device uint32. */
int s4g_evalue_screen(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* cand_ids, const int64_t* cand_offsets,
                      int64_t n_pairs, const int32_t* scores, const char* matrix_name, uint64_t db_residues, int gap_open,
                      int gap_extend, double max_evalue, uint32_t* out_query, uint32_t* out_id, int32_t* out_score, int32_t* out_tlen,
                      uint32_t* out_count);

/* ---- the stages as one call ------------------------------------------------------------------------ */
/* Stage 2 + E-value pre-screen for callers that select over several shards (the CLI with --cards): scores every
 * (query, candidate) pair like s4g_sw_score and returns only the pairs s4g_evalue_screen lets through -- the full score
 * array never crosses the bus.  cand_ids / cand_offsets: host (where = S4G_HOST) or device pointers.  db_residues: the
 * E-value's database length (0 = the whole file the shard was opened from).  The arrays of `out` are pinned host memory
 * owned by the context, in candidate order (grouped by ascending query), valid until the next s4g_score_screen / s4g_search
 * on this context.  Replaces the scoring + eValues sweep of scoreDatabase / extractThread (sw/database.c:402-646,821-875). */
typedef struct s4g_survivors {
    int64_t n;                  /* pairs that passed the screen */
    const uint32_t* query;      /* query index */
    const uint32_t* id;         /* target id */
    const int32_t* score;       /* exact SW score */
    const int32_t* tlen;        /* target length */
    int64_t n_pairs;            /* pairs scored */
    uint64_t sw_cells;          /* sum over the scored pairs of len(query) * len(target) */
    float sw_kernel_ms;         /* device time of the dominant score kernel */
} s4g_survivors;
int s4g_score_screen(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* cand_ids /*[io]*/, const int64_t* cand_offsets /*[io]*/,
                     int64_t n_pairs, int where, const int32_t* matrix /*host*/, const char* matrix_name, uint64_t db_residues,
                     int gap_open, int gap_extend, double max_evalue, s4g_survivors* out);

/* The whole hot path for one resident shard and one query batch: k-mer prefilter -> SW scores -> E-value selection ->
 * traceback, i.e. searchDatabase() followed by alignDatabase() (sift4g/src/main.cpp:203-220, database_search.hpp:17-19,
 * database_alignment.hpp:18-23) behind one entry point.  Candidate lists, scores and survivors stay in HBM between the
 * stages; the exact selection runs on host threads with the reference's libm arithmetic and order (ties by target name when
 * the shard was opened from a file, else by id).  Result arrays are pinned host memory owned by the context, valid until the
 * next s4g_search / s4g_score_screen on it. */
typedef struct s4g_search_params {
    int kmer_length, max_candidates;            /* prefilter: sift4g -k / --max-candidates (main.cpp:88-95) */
    const int32_t* matrix;                      /* 26x26 scorer table */
    const char* matrix_name;                    /* E-value constants; NULL = "BLOSUM_62" */
    int gap_open, gap_extend;
    double max_evalue;
    int max_alignments;
    int n_threads;                              /* host threads of the selection; <= 0: all cores (capped) */
    int want_candidates;                        /* != 0: copy the candidate lists to the host as well */
    int want_alignments;                        /* != 0: trace the kept hits back (stage 3) */
    int device_results;                         /* != 0: coords / paths / path_offsets of the result are DEVICE pointers (valid until the
                                                   next call on the context) and are not copied to the host; the hit lists still are */
} s4g_search_params;
typedef struct s4g_search_result {
    int32_t n_queries;
    int64_t n_pairs;                            /* (query, candidate) pairs scored */
    int64_t n_survivors;                        /* pairs that passed the device screen */
    uint64_t sw_cells, db_residues;
    const uint32_t* cand_ids;                   /* want_candidates: ascending ids per query ... */
    const int64_t* cand_offsets;                /* ... delimited by cand_offsets[n_queries + 1] */
    int64_t n_hits;                             /* kept hits, grouped by query in the reference's order (E asc, score desc, name asc) */
    const uint32_t* hit_query;
    const uint32_t* hit_target;
    const int32_t* hit_score;
    const double* hit_evalue;
    const int64_t* hit_offsets;                 /* [n_queries + 1] */
    const int32_t* coords;                      /* want_alignments: {qstart, qend, tstart, tend} per hit */
    const uint8_t* paths;                       /* moves 1 = DIAG, 2 = LEFT, 3 = UP */
    const int64_t* path_offsets;                /* [n_hits + 1] */
    float ms_prefilter, ms_score, ms_select, ms_align, sw_kernel_ms;     /* wall time per stage, device time of the score kernel */
    int64_t h2d_bytes, d2h_bytes;               /* bytes this call copied over the bus */
} s4g_search_result;
int s4g_search(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const s4g_search_params* params, s4g_search_result* out);

/* ---- behind the hot path: the alignments SIFT4G keeps for the prediction -------------------- */
/* alignmentsExtract (sift4g/src/select_alignments.cpp:127-180 with aligmentStr :244-299): every hit as a string over its
 * query's positions -- the target letter ('A'..'Z') aligned to each position, 'X' in front of / behind the alignment and
 * where the query residue faces a gap.  Host arrays in the layout s4g_search / s4g_sw_align return them; the strings of hit h
 * are written to out_strings[out_string_offsets[h] .. out_string_offsets[h + 1]) (length of its query); out_strings needs
 * sum over hits of len(query) bytes, out_string_offsets n_hits + 1 entries.  The targets must lie in this shard. */
int s4g_alignment_strings(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_hits, const uint32_t* hit_query, const uint32_t* hit_target,
                          const int32_t* coords, const uint8_t* paths, const int64_t* path_offsets, uint8_t* out_strings,
                          int64_t* out_string_offsets);
/* alignmentsSelect (select_alignments.cpp:182-242, getMedian sift4g/src/constants.hpp:77-86): out_selected[q] = how many of
 * query q's strings (hit_offsets[q] .. hit_offsets[q + 1], in the hit order) are kept: strings are added one at a time while
 * the median over the query positions of  log2(20) + sum_a p_a log2 p_a  (float arithmetic of the reference, its
 * all-but-the-last-element sort included) stays above `threshold` (the CLI's --median-threshold, default 2.75).  `strings`:
 * the strings of all hits back to back (what s4g_alignment_strings wrote).  At most 2047 hits per query. */
int s4g_alignments_select(s4g_ctx* ctx, int32_t n_queries, const int32_t* query_lens, const int64_t* hit_offsets, const uint8_t* strings,
                          float threshold, int32_t* out_selected);

/* ---- the --sub-results alignment table from the result buffers -------------------------------- */
/* Per hit {identities, mismatches, gap openings, alignment length} as outputDatabaseBlastM8 counts them
 * (sw/post_proc.c:962-1003), computed on the GPU from the path and the residues.  Host arrays in s4g_search's layout;
 * out_stats: 4 int32 per hit.  The targets must lie in this shard. */
int s4g_alignment_stats(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_hits, const uint32_t* hit_query, const uint32_t* hit_target,
                        const int32_t* coords, const uint8_t* paths, const int64_t* path_offsets, int32_t* out_stats);
/* outputShotgunDatabase() for SW_OUT_DB_BLASTM8 (with_header = 0) / SW_OUT_DB_BLASTM9 (with_header != 0: the "# Fields:" block
 * in front of every query) -- sw/post_proc.c:253-266,962-1049 -- written straight from the buffers: one line per hit, names cut
 * like the reference cuts them.  path NULL = stdout.  Needs no GPU and no context (errors: s4g_last_error(NULL)). */
int s4g_write_blast_tab(const char* path, int with_header, int32_t n_queries, const int64_t* hit_offsets, const char* const* query_names,
                        const char* const* target_names, const int32_t* stats, const int32_t* coords, const double* evalues,
                        const int32_t* scores);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* Sustained issue rate of the DPX / integer ALU pipe (lane-operations per second of
 * VIADDMNMX.S16x2), measured for ~`millis` ms on the context's device: the denominator of the SW
 * roofline (BASELINE.md "Algorithmic work definitions"). */
int s4g_measure_dpx_peak(s4g_ctx* ctx, int millis, double* lane_ops_per_s);
/* device time (ms) spent in the dominant kernel of the most recent s4g_sw_score call, measured with
 * CUDA events on the context's stream (valid after s4g_sync), and the number of DP cells it covered
 * including padding (algorithmic cells are computed by the caller from lengths). */
/* Sustained rate of random 8-byte lookups into an L2-resident 8 MiB table on this device (lookups/s): the request-rate
 * ceiling of the prefilter scan, which probes its presence/rank table once per k-mer position (profiles/r02_prefilter.md). */
int s4g_measure_gather_peak(s4g_ctx* ctx, int millis, double* lookups_per_s);
int s4g_last_sw_kernel_ms(s4g_ctx* ctx, float* ms);
/* Traceback phases of the last s4g_sw_align / s4g_search on the context (CUDA events on its stream): ms3 = {end cells, begin
 * cells, paths}; cells3 = the DP cells each phase has to cover: end cells sum qlen x (t_end + 1) (columns up to the first one
 * that holds the score), begin cells sum (q_end + 1) x (t_end - t_begin + 1), paths sum (q_end - q_begin + 1) x
 * (2 (|dt - dq| + 1) + 1) (the first band SSW tries, ssw.c:549-727).  For the roofline of stage 3 (traceback
 * cells <= 2 qlen tlen + band x qspan per kept hit). */
int s4g_last_align_profile(s4g_ctx* ctx, float* ms3, uint64_t* cells3);

#ifdef __cplusplus
}
#endif
#endif
