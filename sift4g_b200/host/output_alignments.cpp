// outputShotgunDatabase() of the CLI's --sub-results (sw/post_proc.h, sw/post_proc.c:253-266) fed from the result buffers --
// SURVEY section 8f, row F4.  For the tabular formats (bm8, bm9 = the CLI's default) the per-hit identity / mismatch / gap
// counts come from the GPU (s4g_alignment_stats, on the shard that holds the target) and the lines are written by
// s4g_write_blast_tab; nothing walks the DbAlignment objects.  The reference's writer is still in the binary under another
// name (-DoutputShotgunDatabase=outputShotgunDatabase_reference on post_proc.c, host/Makefile): it serves the other formats
// (bm0, light), calls with alignments this session did not produce, and S4G_WRITER=reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "swsharp/swsharp.h"
#include "s4g_session.hpp"

extern "C" void outputShotgunDatabase_reference(DbAlignment*** dbAlignments, int* dbAlignmentsLens, int dbAlignmentsLen, char* path, int type);

extern "C" void outputShotgunDatabase(DbAlignment*** dbAlignments, int* dbAlignmentsLens, int dbAlignmentsLen, char* path, int type) {
    S4gSession& s = s4gSession();
    const char* mode = getenv("S4G_WRITER");
    const bool tabular = type == SW_OUT_DB_BLASTM8 || type == SW_OUT_DB_BLASTM9;
    if (!tabular || (mode && strcmp(mode, "reference") == 0) || s.hits_key != (const void*)dbAlignments || (int)s.hit_off.size() != dbAlignmentsLen + 1) {
        outputShotgunDatabase_reference(dbAlignments, dbAlignmentsLens, dbAlignmentsLen, path, type);
        return;
    }
    const int64_t n_hits = (int64_t)s.hit_q.size();
    const int n_shards = (int)s.shards.size();
    std::vector<int32_t> stats(4 * (size_t)n_hits + 4);
    std::vector<std::vector<int64_t>> mine(n_shards);
    for (int64_t h = 0; h < n_hits; ++h) mine[s.shardOf(s.hit_t[h])].push_back(h);
    s4gForEachShard([&](int d) {
        S4gShard& sh = s.shards[d];
        const int64_t n = (int64_t)mine[d].size();
        if (n == 0) return;
        std::vector<uint32_t> pq(n), pt(n);
        std::vector<int32_t> co(4 * n), st(4 * n);
        std::vector<int64_t> po(n + 1, 0);
        for (int64_t x = 0; x < n; ++x) {
            const int64_t h = mine[d][x];
            pq[x] = s.hit_q[h]; pt[x] = s.hit_t[h];
            memcpy(&co[4 * x], &s.hit_coords[4 * h], 16);
            po[x + 1] = po[x] + (s.hit_path_off[h + 1] - s.hit_path_off[h]);
        }
        std::vector<uint8_t> pa((size_t)po[n] + 1);
        for (int64_t x = 0; x < n; ++x) memcpy(pa.data() + po[x], s.hit_paths.data() + s.hit_path_off[mine[d][x]], (size_t)(po[x + 1] - po[x]));
        s4gCheck(s4g_alignment_stats(sh.ctx, sh.db, sh.queries, n, pq.data(), pt.data(), co.data(), pa.data(), po.data(), st.data()), "s4g_alignment_stats", sh.ctx);
        for (int64_t x = 0; x < n; ++x) memcpy(&stats[4 * mine[d][x]], &st[4 * x], 16);
    });
    std::vector<const char*> qnames(dbAlignmentsLen, ""), tnames((size_t)n_hits + 1, "");
    for (int i = 0; i < dbAlignmentsLen; ++i)
        if (dbAlignmentsLens[i] > 0) qnames[i] = chainGetName(dbAlignmentGetQuery(dbAlignments[i][0]));
    for (int64_t h = 0; h < n_hits; ++h) {
        const S4gShard& sh = s.shards[s.shardOf(s.hit_t[h])];
        tnames[h] = s4g_db_name(sh.db, (int64_t)s.hit_t[h] - sh.lo);
    }
    s4gCheck(s4g_write_blast_tab(path, type == SW_OUT_DB_BLASTM9, dbAlignmentsLen, s.hit_off.data(), qnames.data(), tnames.data(), stats.data(),
                                 s.hit_coords.data(), s.hit_evalue.data(), s.hit_score.data()), "s4g_write_blast_tab");
}
