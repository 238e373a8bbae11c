// Drop-in body of sift4g's alignDatabase() (signature: sift4g/src/database_alignment.hpp:18-23).
// What the reference did per query through swsharp (database.c:402-646: score every candidate, E-values,
// keep the best <= max_alignments, trace them back) is done for the whole query batch at once:
//   scores  : s4g_score_screen (GPU): scores + an E-value pre-screen on the device; only the few per cent of pairs that can
//             pass max_evalue come back
//   E-values: s4g_select_hits on `-t` host threads: the reference's formula in IEEE double with libm, evaluated in its
//             operation order (vendor/swsharp/swsharp/src/evalue.cu:227-273,436-489) -- bit-identical doubles (pinned by the
//             golden E-values of tests/golden) -- and the ordering rule of database.c:1043-1059 incl. the name tie key
//   paths   : s4g_sw_align (GPU)
// The results are handed over as the reference's own DbAlignment objects (malloc'ed paths, borrowed Chain
// pointers into `database`), which is what selectAlignments / outputShotgunDatabase consume.
// With several GPUs (S4G_DEVICES) every GPU scores and traces the candidates that lie in its resident shard; the
// E-values and the selection see all scores of a query at once, as on one GPU.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "database_alignment.hpp"
#include "s4g_session.hpp"

namespace {

struct Row {
    int idx;          // position in the query's candidate list
    int score;
    double value;
};

}  // namespace

void alignDatabase(DbAlignment**** alignments, int** alignments_lengths, Chain*** _database, int32_t* _database_length,
                   const std::string& database_path, Chain** queries, int32_t queries_length,
                   std::vector<std::vector<uint32_t>>& indices, int32_t algorithm, EValueParams* evalue_params, double max_evalue,
                   uint32_t max_alignments, Scorer* scorer, int32_t* cards, int32_t cards_length) {
    (void)evalue_params;                  // the same constants are derived from (matrix name, gap penalties, database length) below
    fprintf(stderr, "** Aligning queries with candidate sequences **\n");
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    if (algorithm != SW_ALIGN) {
        fprintf(stderr, "[ERROR:sift4g_b200] only the SW algorithm is provided by the B200 path\n");
        exit(-1);
    }
    S4gSession& s = s4gSession();
    // `cards` reaches this seam only; the session (opened by searchDatabase) read the same flag from the command line
    if (cards_length > 0 && !getenv("S4G_DEVICES") && !getenv("S4G_DEVICE")) {
        bool same = (int)s.shards.size() == cards_length;
        for (int d = 0; same && d < cards_length; ++d) same = s.shards[d].device == cards[d];
        if (!same) { fprintf(stderr, "[ERROR:sift4g_b200] --cards does not match the GPUs the database shards are resident on\n"); exit(-1); }
    }
    s4gOpenDatabase(database_path);
    s4gUploadQueries(queries, queries_length);
    const int n_shards = (int)s.shards.size();
    const int64_t n_db = s.total_seqs;
    // host metadata of a sequence by its FASTA index (each shard keeps its own range)
    struct Loc { const S4gShard* sh; int64_t local; };
    auto locate = [&](uint32_t id) { const S4gShard& sh = s.shards[s.shardOf(id)]; return Loc{&sh, (int64_t)id - sh.lo}; };
    auto length_of = [&](uint32_t id) { const Loc l = locate(id); const int64_t* off = s4g_db_host_offsets(l.sh->db); return (int)(off[l.local + 1] - off[l.local]); };

    // ---- scores of every (query, candidate) + E-value pre-screen, on the GPU that holds the candidate ----
    std::vector<int64_t> cand_off(queries_length + 1, 0);
    for (int32_t i = 0; i < queries_length; ++i) cand_off[i + 1] = cand_off[i] + (int64_t)indices[i].size();
    const int* table = scorerGetTable(scorer);
    const int gap_open = scorerGetGapOpen(scorer), gap_extend = scorerGetGapExtend(scorer);
    const char* matrix_name = scorerGetName(scorer);
    if (scorerGetMaxCode(scorer) != 26) { fprintf(stderr, "[ERROR:sift4g_b200] protein scorer expected\n"); exit(-1); }
    // candidate ids ascend within a query (database_search.cpp:173-180) and shards are contiguous id ranges: the part of a
    // query's list that a GPU owns is one sub-range
    std::vector<s4g_survivors> surv(n_shards);
    s4gForEachShard([&](int d) {
        S4gShard& sh = s.shards[d];
        std::vector<int64_t> off(queries_length + 1, 0);
        std::vector<uint32_t> ids;
        ids.reserve(n_shards == 1 ? (size_t)cand_off[queries_length] : (size_t)cand_off[queries_length] / n_shards * 2);
        for (int32_t i = 0; i < queries_length; ++i) {
            const auto lo = std::lower_bound(indices[i].begin(), indices[i].end(), sh.lo);
            const auto hi = std::lower_bound(lo, indices[i].end(), sh.hi);
            ids.insert(ids.end(), lo, hi);
            off[i + 1] = (int64_t)ids.size();
        }
        s4gCheck(s4g_score_screen(sh.ctx, sh.db, sh.queries, ids.data(), off.data(), (int64_t)ids.size(), S4G_HOST, table, matrix_name, s.total_residues,
                                  gap_open, gap_extend, max_evalue, &surv[d]), "s4g_score_screen", sh.ctx);
    });

    const double t_score = since();
    // ---- exact E-values + selection on host threads (reference arithmetic and order: sw/evalue.cu:436-489, database.c:1043-1059)
    // survivors of a query from all shards, shard after shard (each shard lists its survivors by ascending query)
    int64_t n_surv = 0;
    for (int d = 0; d < n_shards; ++d) n_surv += surv[d].n;
    std::vector<uint32_t> s_id(n_surv);
    std::vector<int32_t> s_sc(n_surv), s_tl(n_surv), q_lens(queries_length);
    std::vector<const char*> s_name(n_surv);
    std::vector<int64_t> s_off(queries_length + 1, 0);
    {
        std::vector<int64_t> cur(n_shards, 0);
        int64_t w = 0;
        for (int32_t i = 0; i < queries_length; ++i) {
            q_lens[i] = chainGetLength(queries[i]);
            for (int d = 0; d < n_shards; ++d) {
                const s4g_survivors& sv = surv[d];
                int64_t& c = cur[d];
                for (; c < sv.n && sv.query[c] == (uint32_t)i; ++c, ++w) {
                    s_id[w] = sv.id[c]; s_sc[w] = sv.score[c]; s_tl[w] = sv.tlen[c];
                    s_name[w] = s4g_db_name(s.shards[d].db, (int64_t)sv.id[c] - s.shards[d].lo);
                }
            }
            s_off[i + 1] = w;
        }
    }
    const size_t hit_cap = (size_t)queries_length * max_alignments;
    std::vector<uint32_t> pair_q(hit_cap), pair_t(hit_cap);
    std::vector<int32_t> pair_s(hit_cap);
    std::vector<double> pair_e(hit_cap);
    std::vector<int64_t> hit_off(queries_length + 1, 0);
    s4gCheck(s4g_select_hits(s.shards[0].ctx, queries_length, q_lens.data(), s_id.data(), s_off.data(), s_sc.data(), s_tl.data(), s_name.data(), matrix_name,
                             s.total_residues, gap_open, gap_extend, max_evalue, (int)max_alignments, s.host_threads, pair_q.data(), pair_t.data(),
                             pair_s.data(), pair_e.data(), hit_off.data()), "s4g_select_hits", s.shards[0].ctx);
    const int64_t n_hits = hit_off[queries_length];
    std::vector<std::vector<Row>> kept(queries_length);
    for (int32_t i = 0; i < queries_length; ++i) {
        kept[i].resize(hit_off[i + 1] - hit_off[i]);
        for (int64_t h = hit_off[i]; h < hit_off[i + 1]; ++h) {
            const int idx = (int)(std::lower_bound(indices[i].begin(), indices[i].end(), pair_t[h]) - indices[i].begin());
            kept[i][h - hit_off[i]] = {idx, pair_s[h], pair_e[h]};
        }
    }

    // ---- paths of the kept hits ----
    const double t_select = since();
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
        s4gCheck(s4g_sw_align(sh.ctx, sh.db, sh.queries, n_hits, pair_q.data(), pair_t.data(), pair_s.data(), table, gap_open, gap_extend, coords.data(),
                              paths.data(), cap, path_off.data(), S4G_HOST), "s4g_sw_align", sh.ctx);
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
            s4gCheck(s4g_sw_align(sh.ctx, sh.db, sh.queries, n, pq.data(), pt.data(), ps.data(), table, gap_open, gap_extend, co.data(), sh_paths[d].data(),
                                  cap, sh_path_off[d].data(), S4G_HOST), "s4g_sw_align", sh.ctx);
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

    const double t_align = since();
    // ---- hand over as reference objects ----
    // main.cpp deletes its Scorer right after this call and BEFORE it writes the --sub-results table (main.cpp:222-228), whose
    // bm0 / light writers still reach the scorer through the DbAlignment objects (sw/post_proc.c:812-960): a use after free in
    // the reference that its own allocation pattern happens to survive.  The result objects get a scorer of their own
    // (same name, table and penalties) that lives as long as the session.
    static Scorer* result_scorer = nullptr;
    if (result_scorer) scorerDelete(result_scorer);
    result_scorer = scorerCreate(scorerGetName(scorer), (int*)scorerGetTable(scorer), scorerGetMaxCode(scorer), gap_open, gap_extend);
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
                                          coords[4 * h + 3], kept[i][j].idx, kept[i][j].value, kept[i][j].score, result_scorer, path, plen);
        }
        indices[i].clear();   // the reference consumes the candidate lists (database_alignment.cpp:159-161)
    }
    // keep the hits in the C ABI's layout for the selection step (host/select_alignments.cpp)
    s.hits_key = (const void*)out;
    s.hit_q.assign(pair_q.begin(), pair_q.begin() + n_hits);
    s.hit_t.assign(pair_t.begin(), pair_t.begin() + n_hits);
    s.hit_coords = coords;
    s.hit_score.assign(pair_s.begin(), pair_s.begin() + n_hits);
    s.hit_evalue.assign(pair_e.begin(), pair_e.begin() + n_hits);
    s.hit_off = hit_off;
    s.hit_path_off.assign(n_hits + 1, 0);
    for (int64_t x = 0; x < n_hits; ++x) { int plen = 0; path_of(x, plen); s.hit_path_off[x + 1] = s.hit_path_off[x] + plen; }
    s.hit_paths.resize((size_t)s.hit_path_off[n_hits]);
    for (int64_t x = 0; x < n_hits; ++x) { int plen = 0; const uint8_t* src = path_of(x, plen); memcpy(s.hit_paths.data() + s.hit_path_off[x], src, plen); }
    fprintf(stderr, "* sift4g_b200: scores + screen %.3f s, selection %.3f s (%lld hits), paths %.3f s, result objects %.3f s *\n", t_score,
            t_select - t_score, (long long)n_hits, t_align - t_select, since() - t_align);
    fprintf(stderr, "* processing database part 1 (size ~%.2f GB): 100.00/100.00%% *\n\n", s.total_residues / 1e9);
    *alignments = out;
    *alignments_lengths = out_len;
    *_database = database;
    *_database_length = (int32_t)n_db;
}
