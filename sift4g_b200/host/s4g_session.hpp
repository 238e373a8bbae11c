// Process-wide session shared by the two reference seams (searchDatabase / alignDatabase):
// one C-ABI context per GPU, the database opened once -- one resident shard per GPU (contiguous FASTA ranges) --
// and kept in HBM between the prefilter and the alignment stage (the reference parses the FASTA twice:
// sift4g/src/database_search.cpp:81-97, database_alignment.cpp:36-48).
// GPUs: the CLI's own `--cards <digits>` (sift4g/src/main.cpp:123-125,254-262); without the flag every visible GPU, as the
// reference's help text says ("default: all available CUDA cards", main.cpp:329-332).  searchDatabase runs before the seam
// that receives `cards` (alignDatabase), so the flag is read from the process's command line; S4G_DEVICES="0,1,.." (a device
// may be listed twice: two shards on it) or S4G_DEVICE=<n> override it.  One host thread per GPU drives the C ABI inside
// the seams (the reference's analogue: one host thread per card, sw/database.c:497-532); `-t` sets the host threads of the
// FASTA parse and of the exact hit selection.
#pragma once

#include <string>
#include <thread>
#include <vector>

#include "sift4g_b200.h"

struct S4gShard {
    s4g_ctx* ctx = nullptr;
    s4g_db* db = nullptr;
    s4g_queries* queries = nullptr;
    int device = 0;
    uint32_t lo = 0, hi = 0;             // FASTA indices [lo, hi) resident on this GPU
};

struct S4gSession {
    std::vector<S4gShard> shards;
    std::string db_path;
    const void* queries_key = nullptr;   // Chain** the batch was built from
    int queries_n = 0;
    int64_t total_seqs = 0;
    uint64_t total_residues = 0;
    int host_threads = 0;                // -t of the CLI (0: all cores)
    // the kept hits of the last alignDatabase() call in the C ABI's layout, for the selection step behind it
    // (selectAlignments): valid while `hits_key` is the DbAlignment*** that call handed out
    const void* hits_key = nullptr;
    std::vector<uint32_t> hit_q, hit_t;
    std::vector<int32_t> hit_coords, hit_score;
    std::vector<double> hit_evalue;
    std::vector<int64_t> hit_off, hit_path_off;
    std::vector<uint8_t> hit_paths;
    int shardOf(uint32_t id) const {     // shards are contiguous and ascending
        int d = 0;
        while (d + 1 < (int)shards.size() && id >= shards[d].hi) ++d;
        return d;
    }
};

S4gSession& s4gSession();
// exits like the reference's ASSERT (sift4g/src/utils.hpp:13-19) when rc != S4G_OK
void s4gCheck(int rc, const char* what, s4g_ctx* ctx = nullptr);
void s4gOpenDatabase(const std::string& path);
struct Chain;
void s4gUploadQueries(Chain** queries, int queries_length);
// devices the session runs on, in shard order
std::vector<int> s4gDevices();

// f(shard index) on one host thread per GPU (inline when there is a single GPU)
template <class F>
void s4gForEachShard(F f) {
    S4gSession& s = s4gSession();
    const int n = (int)s.shards.size();
    if (n == 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int d = 0; d < n; ++d) th.emplace_back([&f, d] { f(d); });
    for (auto& t : th) t.join();
}
