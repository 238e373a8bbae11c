// Shared declarations of the sift4g_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sift4g_b200.h"

#define S4G_NLET 26          // reference alphabet: 'A'..'Z' (sw/scorer.c:45-72)
#define S4G_PAD_CODE 26      // extra profile row used beyond sequence ends
#define S4G_DB_TAIL_PAD 256  // readable slack after the last residue (vector loads never fault)

struct s4g_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    int64_t launches = 0;
    cudaEvent_t ev_sw0 = nullptr, ev_sw1 = nullptr;
    bool sw_timed = false;
    // traceback phases of the last s4g_sw_align: events around end cells / begin cells / paths, algorithmic cells of each
    cudaEvent_t ev_al[4] = {nullptr, nullptr, nullptr, nullptr};
    bool al_timed = false;
    unsigned long long al_cells[3] = {0, 0, 0};
    // grow-only scratch arena, one buffer per slot (device memory)
    static const int kSlots = 64;
    void* slot_ptr[kSlots] = {nullptr};
    size_t slot_bytes[kSlots] = {0};
    // pinned host staging, grow-only
    static const int kPins = 24;
    void* pin_ptr[kPins] = {nullptr};
    size_t pin_bytes[kPins] = {0};
    // candidate-buffer budget of the prefilter, remembered per batch shape (see s4g_prefilter_device)
    size_t pf_budget = 0;
    int pf_budget_nq = -1;
    uint32_t pf_budget_n = 0;
    // S4G_TRACE=1: per-phase wall times (stream synchronised at every mark) printed to stderr
    bool trace = false;
    double trace_last = 0.0;
    std::vector<std::pair<std::string, double>> trace_acc;
};

struct s4g_db {
    s4g_ctx* ctx = nullptr;
    uint8_t* d_codes = nullptr;     // concatenated codes, FASTA order, + S4G_DB_TAIL_PAD
    bool borrowed_codes = false;    // d_codes points into an s4g_view (NVLink-striped database): not freed here
    int64_t local_lo = 0, local_hi = INT64_MAX;   // bytes of d_codes resident in THIS GPU's HBM (a view: its local stripes)
    int64_t* d_off = nullptr;       // n+1
    uint32_t* d_order = nullptr;    // local sequence indices by ascending length (built by the first prefilter call)
    int64_t n = 0;
    uint64_t residues = 0;
    uint32_t id_base = 0;
    int32_t max_len = 0;
    std::vector<int64_t> h_off;     // always kept (metadata for the host shims)
    std::vector<uint8_t> h_codes;   // kept when created from host memory
    std::vector<std::string> names; // kept when opened from FASTA / a packed file
    int64_t total_seqs = 0;         // whole file (all shards) when opened from a file, else 0
    uint64_t total_residues = 0;
};

struct s4g_queries {
    s4g_ctx* ctx = nullptr;
    uint8_t* d_codes = nullptr;
    int64_t* d_off = nullptr;
    int32_t* d_len_order = nullptr; // query indices by descending length (tile order of the score kernels: neighbouring tiles share a row class)
    int32_t* d_len_rank = nullptr;  // inverse: position of every query in d_len_order
    int32_t n = 0;
    int32_t max_len = 0;
    std::vector<int64_t> h_off;
    std::vector<uint8_t> h_codes;
};

// ---- error plumbing ---------------------------------------------------------------------------
void s4g_set_error(s4g_ctx* ctx, const char* fmt, ...);

#define S4G_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            s4g_set_error((ctx), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                          cudaGetErrorString(e_));                                            \
            return S4G_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define S4G_CHECK_LAUNCH(ctx)                                                                 \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        S4G_CUDA((ctx), cudaGetLastError());                                                  \
    } while (0)

// phase tracing (no-ops unless the context was created with S4G_TRACE=1 in the environment)
void s4g_trace_start(s4g_ctx* ctx);
void s4g_trace_mark(s4g_ctx* ctx, const char* label);      // time since the previous mark/start is charged to `label`
void s4g_trace_report(s4g_ctx* ctx, const char* header);

// grow-only device scratch; returns nullptr (and sets the error) on failure
void* s4g_scratch(s4g_ctx* ctx, int slot, size_t bytes);
void* s4g_pinned(s4g_ctx* ctx, int slot, size_t bytes);

// scratch slot ids
enum {
    SLOT_IO_A = 0, SLOT_IO_B, SLOT_IO_C, SLOT_IO_D, SLOT_IO_E, SLOT_IO_F,
    SLOT_SW_KEYS, SLOT_SW_KEYS2, SLOT_SW_VALS, SLOT_SW_VALS2, SLOT_SW_CUB, SLOT_SW_TILES,
    SLOT_SW_MISC, SLOT_SW_OVF, SLOT_SW_BOUND, SLOT_SW_MAT,
    SLOT_PF_INDEX, SLOT_PF_BITMAP, SLOT_PF_RANK, SLOT_PF_BUCKET, SLOT_PF_HITS, SLOT_PF_CAND,
    SLOT_PF_COUNT, SLOT_PF_THR, SLOT_PF_CUB, SLOT_PF_TMP, SLOT_PF_TMP2, SLOT_PF_SPILL, SLOT_PF_GBUF, SLOT_SW_STRIP,
    SLOT_AL_WORK, SLOT_AL_DIR, SLOT_AL_MISC, SLOT_AL_OUT, SLOT_PF_ENTRY,
    SLOT_SR_ROWS, SLOT_SR_CNT, SLOT_SR_OFF, SLOT_SR_CAND, SLOT_SR_SCORES, SLOT_SR_SURV,
    SLOT_SR_AL_COORDS, SLOT_SR_AL_PATHS, SLOT_SR_AL_POFF, SLOT_SR_HITS, SLOT_SR_OUT_COORDS, SLOT_SR_OUT_PATHS, SLOT_SR_OUT_POFF, SLOT_AL_GROUPS, SLOT_AL_GATHER, SLOT_AL_GATHER_IDX,
    SLOT_AL_LPT_KEYS, SLOT_AL_LPT_KEYS2, SLOT_AL_LPT_LIST,
    SLOT_COUNT
};
static_assert(SLOT_COUNT <= s4g_ctx::kSlots, "scratch slot table too small");

// resident byte range of a view (csrc/view.cu), for s4g_db_create_view
extern "C" void s4g_view_local_range(const s4g_view* v, uint64_t* lo, uint64_t* hi);

// s4g_select_hits that also reports where every kept hit sits in the candidate arrays (select.cu)
int s4g_select_hits_indexed(s4g_ctx* ctx, int32_t nq, const int32_t* query_lens, const uint32_t* cand_ids, const int64_t* cand_offsets,
                            const int32_t* cand_scores, const int32_t* cand_lens, const char* const* cand_names, const char* matrix_name,
                            uint64_t db_residues, int gap_open, int gap_extend, double max_evalue, int max_alignments, int n_threads,
                            uint32_t* out_q, uint32_t* out_t, int32_t* out_score, double* out_evalue, int64_t* out_offsets, uint32_t* out_index);

// ---- stage launchers (device pointers, enqueue on ctx->stream) ----------------------------------
int s4g_sw_score_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, const uint32_t* d_cand_ids,
                        const int64_t* d_cand_off, int64_t n_pairs, const int32_t* h_matrix, int gap_open,
                        int gap_extend, int32_t* d_out);

int s4g_sw_forward_ends_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n, const uint32_t* d_pair_q, const uint32_t* d_pair_t,
                               const int32_t* d_pair_score, const int8_t* d_mat8, int gap_open, int gap_extend, int32_t* d_coords,
                               unsigned long long* d_flags);
int s4g_sw_reverse_begins_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n, const uint32_t* d_pair_q, const uint32_t* d_pair_t,
                                 const int32_t* d_pair_score, const int8_t* d_mat8, int gap_open, int gap_extend, int swalign_all,
                                 int32_t* d_coords, unsigned long long* d_flags);
int s4g_sw_long_query_rows();      // queries longer than this many residues are not handled by the packed kernels

int s4g_prefilter_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int k, int max_candidates,
                         int sorted_by_id, uint32_t* d_ids, float* d_scores, uint32_t* d_counts);

int s4g_sw_align_device(s4g_ctx* ctx, s4g_db* db, s4g_queries* q, int64_t n_pairs, const uint32_t* d_pair_q,
                        const uint32_t* d_pair_t, const int32_t* d_pair_score, const int32_t* h_matrix,
                        int gap_open, int gap_extend, int32_t* d_coords, uint8_t* d_paths, int64_t path_capacity,
                        int64_t* d_path_off, const int64_t* h_path_off);
