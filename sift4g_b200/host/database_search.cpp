// Drop-in body of searchDatabase() (signature: sift4g/src/database_search.hpp:17-19): the k-mer prefilter
// runs on the GPU through the C ABI (s4g_prefilter); everything the reference computed in
// database_search.cpp:66-253 / hash.cpp is behind that call.  num_threads (-t) sets the host threads of the session (FASTA
// parse, candidate merge, hit selection); the candidate set no longer depends on it -- see DESIGN.md, tie rule.
// With several GPUs (S4G_DEVICES) every GPU filters its resident shard and the per-query lists are merged on the
// host exactly like the reference merges its per-thread lists (database_search.cpp:132-154): best max_candidates
// by (score desc, id asc) -- the same result as one GPU, whatever the number of shards.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "database_search.hpp"
#include "s4g_session.hpp"

uint64_t searchDatabase(std::vector<std::vector<uint32_t>>& dst, const std::string& database_path, Chain** queries,
                        int32_t queries_length, uint32_t kmer_length, uint32_t max_candidates, uint32_t num_threads) {
    fprintf(stderr, "** Searching database for candidate sequences **\n");
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    S4gSession& s = s4gSession();
    s.host_threads = (int)num_threads;        // -t: host threads of the FASTA parse, the candidate merge and the hit selection
    const double t_ctx = since();
    s4gOpenDatabase(database_path);
    const double t_open = since();
    s4gUploadQueries(queries, queries_length);
    const int n_shards = (int)s.shards.size();
    const size_t row = max_candidates;

    dst.clear();
    dst.resize(queries_length);
    if (n_shards == 1) {
        std::vector<uint32_t> ids((size_t)queries_length * row);
        std::vector<uint32_t> counts(queries_length);
        S4gShard& sh = s.shards[0];
        s4gCheck(s4g_prefilter(sh.ctx, sh.db, sh.queries, (int)kmer_length, (int)max_candidates, /*sorted_by_id=*/1, ids.data(), nullptr,
                               counts.data(), S4G_HOST), "s4g_prefilter", sh.ctx);
        for (int32_t i = 0; i < queries_length; ++i) dst[i].assign(ids.begin() + (size_t)i * row, ids.begin() + (size_t)i * row + counts[i]);
    } else {
        // best-first rows (ids + float32 scores) of every shard
        std::vector<std::vector<uint32_t>> ids(n_shards), counts(n_shards);
        std::vector<std::vector<float>> scores(n_shards);
        s4gForEachShard([&](int d) {
            S4gShard& sh = s.shards[d];
            ids[d].resize((size_t)queries_length * row);
            scores[d].resize((size_t)queries_length * row);
            counts[d].resize(queries_length);
            s4gCheck(s4g_prefilter(sh.ctx, sh.db, sh.queries, (int)kmer_length, (int)max_candidates, /*sorted_by_id=*/0, ids[d].data(),
                                   scores[d].data(), counts[d].data(), S4G_HOST), "s4g_prefilter", sh.ctx);
        });
        // host merge (the reference's merge of its per-thread lists, database_search.cpp:132-154)
        std::vector<const uint32_t*> p_ids(n_shards), p_counts(n_shards);
        std::vector<const float*> p_scores(n_shards);
        for (int d = 0; d < n_shards; ++d) { p_ids[d] = ids[d].data(); p_scores[d] = scores[d].data(); p_counts[d] = counts[d].data(); }
        std::vector<uint32_t> merged((size_t)queries_length * row), merged_counts(queries_length);
        s4gCheck(s4g_merge_candidates_host(n_shards, queries_length, (int)max_candidates, p_ids.data(), p_scores.data(), p_counts.data(), (int)num_threads,
                                           merged.data(), merged_counts.data()), "s4g_merge_candidates_host");
        for (int32_t i = 0; i < queries_length; ++i) dst[i].assign(merged.begin() + (size_t)i * row, merged.begin() + (size_t)i * row + merged_counts[i]);
    }
    fprintf(stderr, "* sift4g_b200: GPU contexts %.3f s, database resident after %.3f s (%lld sequences), prefilter + candidate lists %.3f s *\n",
            t_ctx, t_open - t_ctx, (long long)s.total_seqs, since() - t_open);
    fprintf(stderr, "* processing database part 1 (size ~%.2f GB): 100.00/100.00%% *\n\n", s.total_residues / 1e9);
    return s.total_residues;
}
