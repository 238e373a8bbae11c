// selectAlignments() of the CLI (signature: sift4g/src/select_alignments.hpp:17-19) on the GPU -- SURVEY section 8f, row F3.
// The reference's own file is still compiled, unchanged, into this binary: its selectAlignments is given another name at
// compile time (-DselectAlignments=selectAlignments_reference, host/Makefile) so that main.cpp's call lands here, and
// S4G_SELECT=reference sends it on to the reference code (the e2e tests run both and compare the files byte for byte).
// outputSelectedAlignments / deleteSelectedAlignments stay the reference's.
//
// What the reference does per query on its thread pool (threadSelectAlignments, select_alignments.cpp:301-321:
// alignmentsExtract, alignmentsSelect, delete the strings that were not selected) is done for the whole batch by
// s4g_alignment_strings (on the GPU that holds the hit's target) and s4g_alignments_select.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "swsharp/swsharp.h"
#include "s4g_session.hpp"

void selectAlignments_reference(std::vector<std::vector<Chain*>>& dst, DbAlignment*** alignments, int32_t* alignments_lengths, Chain** queries,
                                int32_t queries_length, float threshold);

void selectAlignments(std::vector<std::vector<Chain*>>& dst, DbAlignment*** alignments, int32_t* alignments_lengths, Chain** queries,
                      int32_t queries_length, float threshold) {
    S4gSession& s = s4gSession();
    const char* mode = getenv("S4G_SELECT");
    if ((mode && strcmp(mode, "reference") == 0) || s.hits_key != (const void*)alignments) {
        selectAlignments_reference(dst, alignments, alignments_lengths, queries, queries_length, threshold);
        return;
    }
    fprintf(stderr, "** Selecting alignments with median threshold: %.2f **\n", threshold);
    dst.clear();
    dst.resize(queries_length);
    const int64_t n_hits = (int64_t)s.hit_q.size();
    const int n_shards = (int)s.shards.size();
    std::vector<int32_t> q_lens(queries_length);
    std::vector<int64_t> str_off(n_hits + 1, 0);
    for (int32_t i = 0; i < queries_length; ++i) q_lens[i] = chainGetLength(queries[i]);
    for (int64_t h = 0; h < n_hits; ++h) str_off[h + 1] = str_off[h] + q_lens[s.hit_q[h]];
    std::vector<uint8_t> strings((size_t)str_off[n_hits] + 1);
    // strings on the GPU that holds the target
    std::vector<std::vector<int64_t>> mine(n_shards);
    for (int64_t h = 0; h < n_hits; ++h) mine[s.shardOf(s.hit_t[h])].push_back(h);
    s4gForEachShard([&](int d) {
        S4gShard& sh = s.shards[d];
        const int64_t n = (int64_t)mine[d].size();
        if (n == 0) return;
        std::vector<uint32_t> pq(n), pt(n);
        std::vector<int32_t> co(4 * n);
        std::vector<int64_t> po(n + 1, 0), so(n + 1, 0);
        for (int64_t x = 0; x < n; ++x) {
            const int64_t h = mine[d][x];
            pq[x] = s.hit_q[h]; pt[x] = s.hit_t[h];
            memcpy(&co[4 * x], &s.hit_coords[4 * h], 16);
            po[x + 1] = po[x] + (s.hit_path_off[h + 1] - s.hit_path_off[h]);
        }
        std::vector<uint8_t> pa((size_t)po[n] + 1);
        for (int64_t x = 0; x < n; ++x) memcpy(pa.data() + po[x], s.hit_paths.data() + s.hit_path_off[mine[d][x]], (size_t)(po[x + 1] - po[x]));
        int64_t total = 0;
        for (int64_t x = 0; x < n; ++x) total += q_lens[pq[x]];
        std::vector<uint8_t> out((size_t)total + 1);
        s4gCheck(s4g_alignment_strings(sh.ctx, sh.db, sh.queries, n, pq.data(), pt.data(), co.data(), pa.data(), po.data(), out.data(), so.data()),
                 "s4g_alignment_strings", sh.ctx);
        for (int64_t x = 0; x < n; ++x) memcpy(strings.data() + str_off[mine[d][x]], out.data() + so[x], (size_t)(so[x + 1] - so[x]));
    });
    std::vector<int32_t> selected(queries_length, 0);
    s4gCheck(s4g_alignments_select(s.shards[0].ctx, queries_length, q_lens.data(), s.hit_off.data(), strings.data(), threshold, selected.data()),
             "s4g_alignments_select", s.shards[0].ctx);
    for (int32_t i = 0; i < queries_length; ++i) {
        if (alignments_lengths[i] == 0) continue;
        dst[i].reserve(selected[i]);
        for (int32_t j = 0; j < selected[i]; ++j) {
            const int64_t h = s.hit_off[i] + j;
            Chain* target = dbAlignmentGetTarget(alignments[i][j]);
            const char* name = chainGetName(target);
            dst[i].push_back(chainCreate((char*)name, (int)strlen(name), (char*)(strings.data() + str_off[h]), q_lens[i]));
        }
    }
    fprintf(stderr, "\n");
}
