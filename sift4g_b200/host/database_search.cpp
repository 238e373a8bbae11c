// Drop-in body of searchDatabase() (signature: sift4g/src/database_search.hpp:17-19): the k-mer prefilter
// runs on the GPU through the C ABI (s4g_prefilter); everything the reference computed in
// database_search.cpp:66-253 / hash.cpp is behind that call.  num_threads is accepted and ignored (the
// candidate set no longer depends on it -- see DESIGN.md, tie rule).
#include <cstdio>
#include <cstdlib>
#include <string>
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

    std::vector<uint32_t> ids((size_t)queries_length * max_candidates);
    std::vector<uint32_t> counts(queries_length);
    s4gCheck(s4g_prefilter(s.ctx, s.db, s.queries, (int)kmer_length, (int)max_candidates, /*sorted_by_id=*/1, ids.data(), nullptr,
                           counts.data(), S4G_HOST), "s4g_prefilter");
    dst.clear();
    dst.resize(queries_length);
    for (int32_t i = 0; i < queries_length; ++i)
        dst[i].assign(ids.begin() + (size_t)i * max_candidates, ids.begin() + (size_t)i * max_candidates + counts[i]);
    fprintf(stderr, "* processing database part 1 (size ~%.2f GB): 100.00/100.00%% *\n\n", s4g_db_num_residues(s.db) / 1e9);
    return s4g_db_num_residues(s.db);
}
