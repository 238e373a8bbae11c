// Drop-in body of sift4g's alignDatabase() (signature: sift4g/src/database_alignment.hpp:18-23).
// What the reference did per query through swsharp (database.c:402-646: score every candidate, E-values,
// keep the best <= max_alignments, trace them back) is done for the whole query batch at once:
//   scores  : s4g_sw_score (GPU)
//   E-values: the reference's own eValues() host routine, so the doubles are bit-identical to the CPU build
//             (vendor/swsharp/swsharp/src/evalue.cu:227-273,436-489); ordering rule of database.c:1043-1059
//   paths   : s4g_sw_align (GPU)
// The results are handed over as the reference's own DbAlignment objects (malloc'ed paths, borrowed Chain
// pointers into `database`), which is what selectAlignments / outputShotgunDatabase consume.
// With several GPUs (S4G_DEVICES) every GPU scores and traces the candidates that lie in its resident shard; the
// E-values and the selection see all scores of a query at once, as on one GPU.
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
    (void)cards; (void)cards_length;      // GPUs are named by S4G_DEVICES / S4G_DEVICE (s4g_session.hpp)
    fprintf(stderr, "** Aligning queries with candidate sequences **\n");
    if (algorithm != SW_ALIGN) {
        fprintf(stderr, "[ERROR:sift4g_b200] only the SW algorithm is provided by the B200 path\n");
        exit(-1);
    }
    S4gSession& s = s4gSession();
    s4gOpenDatabase(database_path);
    s4gUploadQueries(queries, queries_length);
    const int n_shards = (int)s.shards.size();
    const int64_t n_db = s.total_seqs;
    // host metadata of a sequence by its FASTA index (each shard keeps its own range)
    struct Loc { const S4gShard* sh; int64_t local; };
    auto locate = [&](uint32_t id) { const S4gShard& sh = s.shards[s.shardOf(id)]; return Loc{&sh, (int64_t)id - sh.lo}; };
    auto length_of = [&](uint32_t id) { const Loc l = locate(id); const int64_t* off = s4g_db_host_offsets(l.sh->db); return (int)(off[l.local + 1] - off[l.local]); };
    auto name_of = [&](uint32_t id) { const Loc l = locate(id); return s4g_db_name(l.sh->db, l.local); };

    // ---- scores of every (query, candidate) ----
    std::vector<int64_t> cand_off(queries_length + 1, 0);
    for (int32_t i = 0; i < queries_length; ++i) cand_off[i + 1] = cand_off[i] + (int64_t)indices[i].size();
    const int64_t n_pairs = cand_off[queries_length];
    std::vector<uint32_t> cand_ids(n_pairs);
    for (int32_t i = 0; i < queries_length; ++i) std::copy(indices[i].begin(), indices[i].end(), cand_ids.begin() + cand_off[i]);
    std::vector<int32_t> scores(n_pairs);
    const int* table = scorerGetTable(scorer);
    if (scorerGetMaxCode(scorer) != 26) { fprintf(stderr, "[ERROR:sift4g_b200] protein scorer expected\n"); exit(-1); }
    if (n_shards == 1) {
        S4gShard& sh = s.shards[0];
        s4gCheck(s4g_sw_score(sh.ctx, sh.db, sh.queries, cand_ids.data(), cand_off.data(), n_pairs, table, scorerGetGapOpen(scorer),
                              scorerGetGapExtend(scorer), scores.data(), S4G_HOST), "s4g_sw_score");
    } else {
        // candidate ids ascend within a query (database_search.cpp:173-180), shards are contiguous id ranges: the part of a
        // query's list that a GPU owns is one sub-range, and the scores go back to the same positions
        s4gForEachShard([&](int d) {
            S4gShard& sh = s.shards[d];
            std::vector<int64_t> off(queries_length + 1, 0), first(queries_length, 0);
            for (int32_t i = 0; i < queries_length; ++i) {
                const uint32_t* b = cand_ids.data() + cand_off[i];
                const uint32_t* e = cand_ids.data() + cand_off[i + 1];
                const uint32_t* lo = std::lower_bound(b, e, sh.lo);
                const uint32_t* hi = std::lower_bound(lo, e, sh.hi);
                first[i] = lo - cand_ids.data();
                off[i + 1] = off[i] + (hi - lo);
            }
            std::vector<uint32_t> ids(off[queries_length]);
            for (int32_t i = 0; i < queries_length; ++i) std::copy(cand_ids.begin() + first[i], cand_ids.begin() + first[i] + (off[i + 1] - off[i]), ids.begin() + off[i]);
            std::vector<int32_t> sc(ids.size());
            s4gCheck(s4g_sw_score(sh.ctx, sh.db, sh.queries, ids.data(), off.data(), (int64_t)ids.size(), table, scorerGetGapOpen(scorer),
                                  scorerGetGapExtend(scorer), sc.data(), S4G_HOST), "s4g_sw_score");
            for (int32_t i = 0; i < queries_length; ++i) std::copy(sc.begin() + off[i], sc.begin() + off[i + 1], scores.begin() + first[i]);
        });
    }

    // ---- E-values + selection (host, reference arithmetic) ----
    // eValues() only reads chain lengths: serve it length-only views of one dummy chain
    int max_len = 1;
    for (const S4gShard& sh : s.shards) {
        const int64_t* off = s4g_db_host_offsets(sh.db);
        for (int64_t i = 0; i < (int64_t)(sh.hi - sh.lo); ++i) max_len = std::max<int>(max_len, (int)(off[i + 1] - off[i]));
    }
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
        for (int j = 0; j < n; ++j) views[j] = view_of(length_of(indices[i][j]));
        std::vector<double> values(n);
        eValues(values.data(), scores.data() + cand_off[i], queries[i], views.data(), n, nullptr, 0, evalue_params);
        std::vector<Row> rows(n);
        int thresholded = 0;
        for (int j = 0; j < n; ++j) {
            rows[j] = {j, scores[cand_off[i] + j], values[j], name_of(indices[i][j])};
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
    std::vector<int64_t> path_off(n_hits + 1, 0);       // single GPU: offsets into `paths`
    std::vector<uint8_t> paths;
    // several GPUs: hit h is the where[h].second-th hit of shard where[h].first
    std::vector<std::pair<int, int64_t>> where(n_shards > 1 ? n_hits : 0);
    std::vector<std::vector<uint8_t>> sh_paths(n_shards);
    std::vector<std::vector<int64_t>> sh_path_off(n_shards);
    if (n_shards == 1) {
        S4gShard& sh = s.shards[0];
        int64_t cap = 16;
        for (int64_t h = 0; h < n_hits; ++h) cap += chainGetLength(queries[pair_q[h]]) + length_of(pair_t[h]);
        paths.resize(cap);
        s4gCheck(s4g_sw_align(sh.ctx, sh.db, sh.queries, n_hits, pair_q.data(), pair_t.data(), pair_s.data(), table, scorerGetGapOpen(scorer),
                              scorerGetGapExtend(scorer), coords.data(), paths.data(), cap, path_off.data(), S4G_HOST), "s4g_sw_align");
    } else {
        std::vector<std::vector<int64_t>> mine(n_shards);            // global hit numbers per shard, in order
        for (int64_t h = 0; h < n_hits; ++h) {
            const int d = s.shardOf(pair_t[h]);
            where[h] = {d, (int64_t)mine[d].size()};
            mine[d].push_back(h);
        }
        s4gForEachShard([&](int d) {
            S4gShard& sh = s.shards[d];
            const int64_t n = (int64_t)mine[d].size();
            std::vector<uint32_t> pq(n), pt(n);
            std::vector<int32_t> ps(n), co(4 * n);
            int64_t cap = 16;
            for (int64_t x = 0; x < n; ++x) {
                const int64_t h = mine[d][x];
                pq[x] = pair_q[h]; pt[x] = pair_t[h]; ps[x] = pair_s[h];
                cap += chainGetLength(queries[pq[x]]) + length_of(pt[x]);
            }
            sh_paths[d].resize(cap);
            sh_path_off[d].assign(n + 1, 0);
            s4gCheck(s4g_sw_align(sh.ctx, sh.db, sh.queries, n, pq.data(), pt.data(), ps.data(), table, scorerGetGapOpen(scorer),
                                  scorerGetGapExtend(scorer), co.data(), sh_paths[d].data(), cap, sh_path_off[d].data(), S4G_HOST), "s4g_sw_align");
            for (int64_t x = 0; x < n; ++x) std::copy(co.begin() + 4 * x, co.begin() + 4 * x + 4, coords.begin() + 4 * mine[d][x]);
        });
    }
    auto path_of = [&](int64_t h, int& plen) -> const uint8_t* {
        if (n_shards == 1) { plen = (int)(path_off[h + 1] - path_off[h]); return paths.data() + path_off[h]; }
        const int d = where[h].first;
        const int64_t x = where[h].second;
        plen = (int)(sh_path_off[d][x + 1] - sh_path_off[d][x]);
        return sh_paths[d].data() + sh_path_off[d][x];
    };

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
                const Loc l = locate(t);
                const int64_t* off = s4g_db_host_offsets(l.sh->db);
                const uint8_t* codes = s4g_db_host_codes(l.sh->db);
                const int len = (int)(off[l.local + 1] - off[l.local]);
                text.resize(len);
                for (int x = 0; x < len; ++x) text[x] = (char)('A' + codes[off[l.local] + x]);
                const char* name = s4g_db_name(l.sh->db, l.local);
                database[t] = chainCreate((char*)name, (int)strlen(name), (char*)text.c_str(), len);
            }
            int plen = 0;
            const uint8_t* src = path_of(h, plen);
            char* path = (char*)malloc(plen > 0 ? plen : 1);
            memcpy(path, src, plen);
            out[i][j] = dbAlignmentCreate(queries[i], coords[4 * h + 0], coords[4 * h + 1], 0, database[t], coords[4 * h + 2],
                                          coords[4 * h + 3], kept[i][j].idx, kept[i][j].value, kept[i][j].score, scorer, path, plen);
        }
        indices[i].clear();   // the reference consumes the candidate lists (database_alignment.cpp:159-161)
    }
    for (auto& kv : len_view) chainDelete(kv.second);
    chainDelete(dummy_chain);

    fprintf(stderr, "* processing database part 1 (size ~%.2f GB): 100.00/100.00%% *\n\n", s.total_residues / 1e9);
    *alignments = out;
    *alignments_lengths = out_len;
    *_database = database;
    *_database_length = (int32_t)n_db;
}
