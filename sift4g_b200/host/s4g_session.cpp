#include "s4g_session.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "swsharp/swsharp.h"

S4gSession& s4gSession() {
    static S4gSession s;
    if (s.shards.empty()) {
        std::vector<int> devices;
        if (const char* list = getenv("S4G_DEVICES")) {
            for (const char* p = list; *p;) {
                char* end = nullptr;
                const long v = strtol(p, &end, 10);
                if (end == p) break;
                devices.push_back((int)v);
                p = *end == ',' ? end + 1 : end;
            }
        }
        if (devices.empty()) {
            const char* dev = getenv("S4G_DEVICE");
            devices.push_back(dev ? atoi(dev) : 0);
        }
        s.shards.resize(devices.size());
        for (size_t d = 0; d < devices.size(); ++d) s4gCheck(s4g_init(devices[d], &s.shards[d].ctx), "s4g_init");
    }
    return s;
}

void s4gCheck(int rc, const char* what) {
    if (rc == S4G_OK) return;
    fprintf(stderr, "[ERROR:sift4g_b200] %s failed (%d): %s\n", what, rc, s4g_last_error(nullptr));
    exit(-1);
}

void s4gOpenDatabase(const std::string& path) {
    S4gSession& s = s4gSession();
    if (s.shards[0].db && s.db_path == path) return;
    const int n = (int)s.shards.size();
    // FASTA, or a packed .s4gdb written by bin/s4g_pack (told apart by the magic); shard d of n on GPU d
    s4gForEachShard([&](int d) {
        S4gShard& sh = s.shards[d];
        if (sh.db) { s4g_db_close(sh.db); sh.db = nullptr; }
        s4gCheck(s4g_db_open(sh.ctx, path.c_str(), d, n, &sh.db), "s4g_db_open");
        sh.lo = s4g_db_id_base(sh.db);
        sh.hi = sh.lo + (uint32_t)s4g_db_num_seqs(sh.db);
    });
    s.total_seqs = s4g_db_total_seqs(s.shards[0].db);
    s.total_residues = s4g_db_total_residues(s.shards[0].db);
    s.db_path = path;
}

void s4gUploadQueries(Chain** queries, int queries_length) {
    S4gSession& s = s4gSession();
    if (s.shards[0].queries && s.queries_key == (const void*)queries && s.queries_n == queries_length) return;
    std::vector<int64_t> off(queries_length + 1, 0);
    for (int i = 0; i < queries_length; ++i) off[i + 1] = off[i] + chainGetLength(queries[i]);
    std::vector<uint8_t> codes(off[queries_length]);
    for (int i = 0; i < queries_length; ++i) {
        const char* c = chainGetCodes(queries[i]);
        for (int j = 0; j < chainGetLength(queries[i]); ++j) codes[off[i] + j] = (uint8_t)c[j];
    }
    s4gForEachShard([&](int d) {
        S4gShard& sh = s.shards[d];
        if (sh.queries) { s4g_queries_free(sh.queries); sh.queries = nullptr; }
        s4gCheck(s4g_queries_create(sh.ctx, codes.data(), off.data(), queries_length, S4G_HOST, &sh.queries), "s4g_queries_create");
    });
    s.queries_key = (const void*)queries;
    s.queries_n = queries_length;
}
