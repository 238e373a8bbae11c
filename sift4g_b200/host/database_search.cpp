// Drop-in body of searchDatabase() (signature: sift4g/src/database_search.hpp:17-19): the k-mer prefilter
// runs on the GPU through the C ABI (s4g_prefilter); everything the reference computed in
// database_search.cpp:66-253 / hash.cpp is behind that call.  num_threads is accepted and ignored (the
// candidate set no longer depends on it -- see DESIGN.md, tie rule).
// With several GPUs (S4G_DEVICES) every GPU filters its resident shard and the per-query lists are merged on the
// host exactly like the reference merges its per-thread lists (database_search.cpp:132-154): best max_candidates
// by (score desc, id asc) -- the same result as one GPU, whatever the number of shards.
#include <algorithm>
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
    (void)num_threads;
    fprintf(stderr, "** Searching database for candidate sequences **\n");
    S4gSession& s = s4gSession();
    s4gOpenDatabase(database_path);
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
                               counts.data(), S4G_HOST), "s4g_prefilter");
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
                                   scores[d].data(), counts[d].data(), S4G_HOST), "s4g_prefilter");
        });
        // host merge, queries dealt to the host threads: key = (~score bits, id) ascending = (score desc, id asc)
        const int n_threads = std::max(1, std::min<int>(32, (int)std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t] {
                std::vector<unsigned long long> keys;
                for (int32_t i = t; i < queries_length; i += n_threads) {
                    keys.clear();
                    for (int d = 0; d < n_shards; ++d) {
                        const uint32_t* id = ids[d].data() + (size_t)i * row;
                        const float* sc = scores[d].data() + (size_t)i * row;
                        for (uint32_t j = 0; j < counts[d][i]; ++j) {
                            uint32_t bits;
                            memcpy(&bits, sc + j, 4);
                            keys.push_back(((unsigned long long)(~bits) << 32) | id[j]);
                        }
                    }
                    const size_t keep = std::min<size_t>(keys.size(), row);
                    if (keep < keys.size()) std::nth_element(keys.begin(), keys.begin() + keep, keys.end());
                    dst[i].resize(keep);
                    for (size_t j = 0; j < keep; ++j) dst[i][j] = (uint32_t)keys[j];
                    std::sort(dst[i].begin(), dst[i].end());          // output ids ascending (database_search.cpp:173-180)
                }
            });
        for (auto& x : th) x.join();
    }
    fprintf(stderr, "* processing database part 1 (size ~%.2f GB): 100.00/100.00%% *\n\n", s.total_residues / 1e9);
    return s.total_residues;
}
