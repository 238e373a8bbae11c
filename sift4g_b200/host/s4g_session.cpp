#include "s4g_session.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "swsharp/swsharp.h"

namespace {

// `--cards <digits>` of this process (sift4g/src/main.cpp:123-125: one card index per character)
bool cardsFromCommandLine(std::vector<int>& devices) {
    FILE* f = fopen("/proc/self/cmdline", "rb");
    if (!f) return false;
    std::vector<std::string> args;
    std::string cur;
    for (int c; (c = fgetc(f)) != EOF;) {
        if (c == 0) { args.push_back(cur); cur.clear(); } else cur.push_back((char)c);
    }
    fclose(f);
    if (!cur.empty()) args.push_back(cur);
    for (size_t i = 1; i < args.size(); ++i) {
        std::string v;
        if (args[i] == "--cards" && i + 1 < args.size()) v = args[i + 1];
        else if (args[i].compare(0, 8, "--cards=") == 0) v = args[i].substr(8);
        else continue;
        for (char ch : v) if (ch >= '0' && ch <= '9') devices.push_back(ch - '0');
        return !devices.empty();
    }
    return false;
}

}  // namespace

std::vector<int> s4gDevices() {
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
    if (devices.empty()) if (const char* dev = getenv("S4G_DEVICE")) devices.push_back(atoi(dev));
    if (devices.empty()) cardsFromCommandLine(devices);
    if (devices.empty()) {
        const int n = s4g_device_count();
        for (int i = 0; i < n; ++i) devices.push_back(i);
    }
    if (devices.empty()) devices.push_back(0);           // s4g_init then reports that there is no device (no CPU fallback)
    return devices;
}

S4gSession& s4gSession() {
    static S4gSession s;
    if (s.shards.empty()) {
        const std::vector<int> devices = s4gDevices();
        s.shards.resize(devices.size());
        for (size_t d = 0; d < devices.size(); ++d) {
            s.shards[d].device = devices[d];
            s4gCheck(s4g_init(devices[d], &s.shards[d].ctx), "s4g_init");
        }
        fprintf(stderr, "* sift4g_b200: %zu GPU%s (device", devices.size(), devices.size() == 1 ? "" : "s");
        for (int d : devices) fprintf(stderr, " %d", d);
        fprintf(stderr, "), one resident database shard each *\n");
    }
    return s;
}

void s4gCheck(int rc, const char* what, s4g_ctx* ctx) {
    if (rc == S4G_OK) return;
    fprintf(stderr, "[ERROR:sift4g_b200] %s failed (%d): %s\n", what, rc, s4g_last_error(ctx));
    exit(-1);
}

void s4gOpenDatabase(const std::string& path) {
    S4gSession& s = s4gSession();
    if (s.shards[0].db && s.db_path == path) return;
    const int n = (int)s.shards.size();
    // FASTA, or a packed .s4gdb written by bin/s4g_pack (told apart by the magic): read / parsed once, shard d of n on GPU d
    std::vector<s4g_ctx*> ctxs(n);
    std::vector<s4g_db*> dbs(n, nullptr);
    for (int d = 0; d < n; ++d) {
        if (s.shards[d].db) { s4g_db_close(s.shards[d].db); s.shards[d].db = nullptr; }
        ctxs[d] = s.shards[d].ctx;
    }
    s4gCheck(s4g_db_open_sharded(ctxs.data(), n, path.c_str(), s.host_threads, dbs.data()), "s4g_db_open_sharded", ctxs[0]);
    for (int d = 0; d < n; ++d) {
        S4gShard& sh = s.shards[d];
        sh.db = dbs[d];
        sh.lo = s4g_db_id_base(sh.db);
        sh.hi = sh.lo + (uint32_t)s4g_db_num_seqs(sh.db);
    }
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
        s4gCheck(s4g_queries_create(sh.ctx, codes.data(), off.data(), queries_length, S4G_HOST, &sh.queries), "s4g_queries_create", sh.ctx);
    });
    s.queries_key = (const void*)queries;
    s.queries_n = queries_length;
}
