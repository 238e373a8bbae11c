// Drop-in body of sift4g's alignDatabase() (signature: sift4g/src/database_alignment.hpp:18-23).
// What the reference did per query through swsharp (database.c:402-646: score every candidate, E-values,
// keep the best <= max_alignments, trace them back) is done for the whole query batch at once:
//   scores  : s4g_sw_score (GPU)
//   E-values: the reference's own eValues() host routine, so the doubles are bit-identical to the CPU build
//             (vendor/swsharp/swsharp/src/evalue.cu:227-273,436-489); ordering rule of database.c:1043-1059
//   paths   : s4g_sw_align (GPU)
// The results are handed over as the reference's own DbAlignment objects (malloc'ed paths, borrowed Chain
// pointers into `database`), which is what selectAlignments / outputShotgunDatabase consume.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "database_alignment.hpp"
#include "s4g_session.hpp"

namespace {

struct Row {
    int idx;          // position in the query's candidate list
    int score;
    double value;
    const char* name;
};

bool rowLess(const Row& a, const Row& b) {    // dbAlignmentDataCmp, database.c:1043-1059
    if (a.value == b.value) {
        if (a.score == b.score) return strcmp(a.name, b.name) < 0;
        return a.score > b.score;
    }
    return a.value < b.value;
}

}  // namespace

void alignDatabase(DbAlignment**** alignments, int** alignments_lengths, Chain*** _database, int32_t* _database_length,
                   const std::string& database_path, Chain** queries, int32_t queries_length,
                   std::vector<std::vector<uint32_t>>& indices, int32_t algorithm, EValueParams* evalue_params, double max_evalue,
                   uint32_t max_alignments, Scorer* scorer, int32_t* cards, int32_t cards_length) {
    (void)cards; (void)cards_length;
    fprintf(stderr, "** Aligning queries with candidate sequences **\n");
    if (algorithm != SW_ALIGN) {
        fprintf(stderr, "[ERROR:sift4g_b200] only the SW algorithm is provided by the B200 path\n");
        exit(-1);
    }
    S4gSession& s = s4gSession();
    s4gOpenDatabase(database_path);
    s4gUploadQueries(queries, queries_length);
    const int64_t n_db = s4g_db_num_seqs(s.db);
    const int64_t* db_off = s4g_db_host_offsets(s.db);
    const uint8_t* db_codes = s4g_db_host_codes(s.db);

    // ---- scores of every (query, candidate) ----
    std::vector<int64_t> cand_off(queries_length + 1, 0);
    for (int32_t i = 0; i < queries_length; ++i) cand_off[i + 1] = cand_off[i] + (int64_t)indices[i].size();
    const int64_t n_pairs = cand_off[queries_length];
    std::vector<uint32_t> cand_ids(n_pairs);
    for (int32_t i = 0; i < queries_length; ++i) std::copy(indices[i].begin(), indices[i].end(), cand_ids.begin() + cand_off[i]);
    std::vector<int32_t> scores(n_pairs);
    const int* table = scorerGetTable(scorer);
    if (scorerGetMaxCode(scorer) != 26) { fprintf(stderr, "[ERROR:sift4g_b200] protein scorer expected\n"); exit(-1); }
    s4gCheck(s4g_sw_score(s.ctx, s.db, s.queries, cand_ids.data(), cand_off.data(), n_pairs, table, scorerGetGapOpen(scorer),
                          scorerGetGapExtend(scorer), scores.data(), S4G_HOST), "s4g_sw_score");

    // ---- E-values + selection (host, reference arithmetic) ----
    // eValues() only reads chain lengths: serve it length-only views of one dummy chain
    int max_len = 1;
    for (int64_t i = 0; i < n_db; ++i) max_len = std::max<int>(max_len, (int)(db_off[i + 1] - db_off[i]));
    std::string dummy((size_t)max_len, 'A');
    Chain* dummy_chain = chainCreate((char*)"len", 3, (char*)dummy.c_str(), max_len);
    std::map<int, Chain*> len_view;
    auto view_of = [&](int len) {
        auto it = len_view.find(len);
        if (it != len_view.end()) return it->second;
        Chain* v = chainCreateView(dummy_chain, 0, len - 1, 0);
        len_view[len] = v;
        return v;
    };

    std::vector<std::vector<Row>> kept(queries_length);
    std::vector<uint32_t> pair_q, pair_t;
    std::vector<int32_t> pair_s;
    for (int32_t i = 0; i < queries_length; ++i) {
        const int n = (int)indices[i].size();
        if (n == 0) continue;
        std::vector<Chain*> views(n);
        for (int j = 0; j < n; ++j) views[j] = view_of((int)(db_off[indices[i][j] + 1] - db_off[indices[i][j]]));
        std::vector<double> values(n);
        eValues(values.data(), scores.data() + cand_off[i], queries[i], views.data(), n, nullptr, 0, evalue_params);
        std::vector<Row> rows(n);
        int thresholded = 0;
        for (int j = 0; j < n; ++j) {
            rows[j] = {j, scores[cand_off[i] + j], values[j], s4g_db_name(s.db, indices[i][j])};
            if (values[j] <= max_evalue) ++thresholded;
        }
        const int k = std::min<int>(thresholded, std::min<int>((int)max_alignments, n));   // database.c:347-349,866
        std::partial_sort(rows.begin(), rows.begin() + k, rows.end(), rowLess);
        rows.resize(k);
        for (int j = 0; j < k; ++j) {
            pair_q.push_back((uint32_t)i);
            pair_t.push_back(indices[i][rows[j].idx]);
            pair_s.push_back(rows[j].score);
        }
        kept[i].swap(rows);
    }

    // ---- paths of the kept hits ----
    const int64_t n_hits = (int64_t)pair_q.size();
    std::vector<int32_t> coords(4 * n_hits);
    std::vector<int64_t> path_off(n_hits + 1, 0);
    int64_t cap = 16;
    for (int64_t h = 0; h < n_hits; ++h)
        cap += chainGetLength(queries[pair_q[h]]) + (db_off[pair_t[h] + 1] - db_off[pair_t[h]]);
    std::vector<uint8_t> paths(cap);
    s4gCheck(s4g_sw_align(s.ctx, s.db, s.queries, n_hits, pair_q.data(), pair_t.data(), pair_s.data(), table, scorerGetGapOpen(scorer),
                          scorerGetGapExtend(scorer), coords.data(), paths.data(), cap, path_off.data(), S4G_HOST), "s4g_sw_align");

    // ---- hand over as reference objects ----
    Chain** database = (Chain**)calloc((size_t)std::max<int64_t>(n_db, 1), sizeof(Chain*));
    DbAlignment*** out = (DbAlignment***)malloc(queries_length * sizeof(DbAlignment**));
    int* out_len = (int*)malloc(queries_length * sizeof(int));
    int64_t h = 0;
    std::string text;
    for (int32_t i = 0; i < queries_length; ++i) {
        const int k = (int)kept[i].size();
        out_len[i] = k;
        out[i] = k ? (DbAlignment**)malloc(k * sizeof(DbAlignment*)) : nullptr;
        for (int j = 0; j < k; ++j, ++h) {
            const uint32_t t = pair_t[h];
            if (database[t] == nullptr) {
                const int len = (int)(db_off[t + 1] - db_off[t]);
                text.resize(len);
                for (int x = 0; x < len; ++x) text[x] = (char)('A' + db_codes[db_off[t] + x]);
                const char* name = s4g_db_name(s.db, t);
                database[t] = chainCreate((char*)name, (int)strlen(name), (char*)text.c_str(), len);
            }
            const int plen = (int)(path_off[h + 1] - path_off[h]);
            char* path = (char*)malloc(plen > 0 ? plen : 1);
            memcpy(path, paths.data() + path_off[h], plen);
            out[i][j] = dbAlignmentCreate(queries[i], coords[4 * h + 0], coords[4 * h + 1], 0, database[t], coords[4 * h + 2],
                                          coords[4 * h + 3], kept[i][j].idx, kept[i][j].value, kept[i][j].score, scorer, path, plen);
        }
        indices[i].clear();   // the reference consumes the candidate lists (database_alignment.cpp:159-161)
    }
    for (auto& kv : len_view) chainDelete(kv.second);
    chainDelete(dummy_chain);

    fprintf(stderr, "* processing database part 1 (size ~%.2f GB): 100.00/100.00%% *\n\n", s4g_db_num_residues(s.db) / 1e9);
    *alignments = out;
    *alignments_lengths = out_len;
    *_database = database;
    *_database_length = (int32_t)n_db;
}
