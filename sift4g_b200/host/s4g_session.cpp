#include "s4g_session.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "swsharp/swsharp.h"

S4gSession& s4gSession() {
    static S4gSession s;
    if (!s.ctx) {
        const char* dev = getenv("S4G_DEVICE");
        s4gCheck(s4g_init(dev ? atoi(dev) : 0, &s.ctx), "s4g_init");
    }
    return s;
}

void s4gCheck(int rc, const char* what) {
    if (rc == S4G_OK) return;
    S4gSession* s = nullptr;
    (void)s;
    fprintf(stderr, "[ERROR:sift4g_b200] %s failed (%d): %s\n", what, rc, s4g_last_error(nullptr));
    exit(-1);
}

void s4gOpenDatabase(const std::string& path) {
    S4gSession& s = s4gSession();
    if (s.db && s.db_path == path) return;
    if (s.db) { s4g_db_close(s.db); s.db = nullptr; }
    // FASTA, or a packed .s4gdb written by bin/s4g_pack (told apart by the magic)
    s4gCheck(s4g_db_open(s.ctx, path.c_str(), 0, 1, &s.db), "s4g_db_open");
    s.db_path = path;
}

void s4gUploadQueries(Chain** queries, int queries_length) {
    S4gSession& s = s4gSession();
    if (s.queries && s.queries_key == (const void*)queries && s.queries_n == queries_length) return;
    if (s.queries) { s4g_queries_free(s.queries); s.queries = nullptr; }
    std::vector<int64_t> off(queries_length + 1, 0);
    for (int i = 0; i < queries_length; ++i) off[i + 1] = off[i] + chainGetLength(queries[i]);
    std::vector<uint8_t> codes(off[queries_length]);
    for (int i = 0; i < queries_length; ++i) {
        const char* c = chainGetCodes(queries[i]);
        for (int j = 0; j < chainGetLength(queries[i]); ++j) codes[off[i] + j] = (uint8_t)c[j];
    }
    s4gCheck(s4g_queries_create(s.ctx, codes.data(), off.data(), queries_length, S4G_HOST, &s.queries), "s4g_queries_create");
    s.queries_key = (const void*)queries;
    s.queries_n = queries_length;
}
