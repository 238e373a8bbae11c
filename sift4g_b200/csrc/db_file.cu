// Database files: the FASTA reader (reference quirks kept, parsed on all host cores) and the packed on-disk
// database ".s4gdb"  that replaces the reference's two FASTA parses per run
// (sift4g/src/database_search.cpp:81-97, database_alignment.cpp:36-48) and its own cache format
// (".swsharp", sw/pre_proc.c:309-374,540-595) with a file that is the HBM layout: open = pread + one H2D copy.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace {

struct MappedFile {
    const char* p = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        size = (size_t)st.st_size;
        if (size == 0) return true;
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, size, MADV_SEQUENTIAL);
        p = (const char*)m;
        return true;
    }
    ~MappedFile() {
        if (p) munmap((void*)p, size);
        if (fd >= 0) ::close(fd);
    }
};

// One piece of the file: [begin, end) starts at a record header ('>' right after a '\n', or byte 0) and ends in
// front of the next such header, so it can be parsed with the reader's start state.
struct Piece {
    size_t begin = 0, end = 0;
    int64_t n_records = 0, n_residues = 0, name_bytes = 0;
    long long error_at = -1;
};

// The reference reader's state machine (sw/pre_proc.c:437-538 + sw/chain.c:59-105) over one piece:
//  * a record ends at the next '>' met outside a header line, or at the LAST BYTE of the file (consumed as a
//    terminator, so a file without trailing newline loses it) -- unless the file size is a multiple of the
//    reader's 1 MiB buffer, in which case the last record is never closed (sw/pre_proc.c:465-488);
//  * name = header line without leading '>'/whitespace, without '\r', without trailing whitespace;
//  * residues: letters only, case folded to 0..25; everything else is dropped;
//  * a record with an empty name or zero residues is an error (the reference aborts).
// EMIT = false counts, EMIT = true writes codes / lengths / names at the piece's slots.
template <bool EMIT>
void parse_piece(const char* data, size_t total, Piece& pc, bool last_piece, uint8_t* codes, int64_t* lens, std::string* names) {
    bool in_name = true;
    int64_t cur_len = 0, n_rec = 0, n_res = 0, name_bytes = 0;
    std::string name;
    const bool eof_closes = (total % (1u << 20)) != 0;
    auto close_record = [&](size_t pos) -> bool {
        while (!name.empty() && isspace((unsigned char)name.back())) name.pop_back();
        if (name.empty() || cur_len == 0) { pc.error_at = (long long)pos; return false; }
        if (EMIT) { lens[n_rec] = cur_len; names[n_rec].swap(name); }
        name_bytes += EMIT ? 0 : (int64_t)name.size() + 1;
        name.clear();
        ++n_rec;
        cur_len = 0;
        return true;
    };
    for (size_t pos = pc.begin; pos < pc.end; ++pos) {
        const char c = data[pos];
        if (!in_name) {
            const unsigned char u = (unsigned char)c;
            // the file's last byte is a terminator, not a residue: it is neither stored nor counted (the count pass
            // sized `codes` without it)
            const bool term = last_piece && pos == total - 1 && eof_closes;
            if (!term && (unsigned)((u | 32) - 'a') < 26u) {   // the common case first
                if (EMIT) codes[n_res] = (uint8_t)((u | 32) - 'a');
                ++n_res; ++cur_len;
                continue;
            }
            if (c == '>' || term) {
                if (!close_record(pos)) return;
                in_name = true;
            } else {
                continue;
            }
        }
        if (c == '\n') in_name = false;
        else if (!(name.empty() && (c == '>' || isspace((unsigned char)c))) && c != '\r') name.push_back(c);
    }
    // the next piece starts with a header: that '>' closes our last record
    if (!last_piece && !in_name && !close_record(pc.end)) return;
    if (!last_piece && in_name) { pc.error_at = (long long)pc.end; return; }
    pc.n_records = n_rec; pc.n_residues = n_res; pc.name_bytes = name_bytes;
}

struct Parsed {
    std::vector<uint8_t> codes;
    std::vector<int64_t> off;
    std::vector<std::string> names;
};

int parse_fasta(s4g_ctx* ctx, const char* path, Parsed& out, int want_threads = 0) {
    MappedFile f;
    if (!f.open(path)) { s4g_set_error(ctx, "cannot open '%s'", path); return S4G_ERR_IO; }
    const char* data = f.p;
    const size_t total = f.size;
    unsigned hw = want_threads > 0 ? (unsigned)want_threads : std::thread::hardware_concurrency();
    if (const char* t = getenv("S4G_HOST_THREADS")) hw = (unsigned)atoi(t);
    const size_t n_threads = std::max<size_t>(1, std::min<size_t>(hw ? hw : 1, 64));
    // cut points: a '>' that directly follows a '\n' is always met outside a header line, i.e. always starts a record
    std::vector<Piece> pieces;
    const size_t want = total < (1u << 20) ? 1 : n_threads * 4;
    size_t begin = 0;
    for (size_t i = 1; i <= want && begin < total; ++i) {
        size_t cut = i == want ? total : std::max(begin + 1, total / want * i);
        while (cut < total && !(data[cut] == '>' && data[cut - 1] == '\n')) {
            const void* nl = memchr(data + cut, '\n', total - cut);
            cut = nl ? (size_t)((const char*)nl - data) + 1 : total;
        }
        Piece pc; pc.begin = begin; pc.end = cut;
        pieces.push_back(pc);
        begin = cut;
    }
    if (pieces.empty()) { Piece pc; pieces.push_back(pc); }
    auto run = [&](bool emit, std::vector<int64_t>* rec0, std::vector<int64_t>* res0, std::vector<int64_t>* lens) {
        std::atomic<size_t> next(0);
        std::vector<std::thread> pool;
        auto work = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < pieces.size();) {
                const bool last = i + 1 == pieces.size();
                if (!emit) parse_piece<false>(data, total, pieces[i], last, nullptr, nullptr, nullptr);
                else parse_piece<true>(data, total, pieces[i], last, out.codes.data() + (*res0)[i], lens->data() + (*rec0)[i], out.names.data() + (*rec0)[i]);
            }
        };
        for (size_t t = 1; t < std::min(n_threads, pieces.size()); ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    };
    const bool trace = getenv("S4G_TRACE") && getenv("S4G_TRACE")[0] && getenv("S4G_TRACE")[0] != '0';
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    run(false, nullptr, nullptr, nullptr);
    const double t1 = now();
    std::vector<int64_t> rec0(pieces.size() + 1, 0), res0(pieces.size() + 1, 0);
    for (size_t i = 0; i < pieces.size(); ++i) {
        if (pieces[i].error_at >= 0) { s4g_set_error(ctx, "'%s': empty record near byte %lld", path, pieces[i].error_at); return S4G_ERR_IO; }
        rec0[i + 1] = rec0[i] + pieces[i].n_records;
        res0[i + 1] = res0[i] + pieces[i].n_residues;
    }
    const int64_t n = rec0.back();
    out.codes.resize((size_t)res0.back());
    out.names.resize((size_t)n);
    std::vector<int64_t> lens((size_t)n);
    const double t2 = now();
    run(true, &rec0, &res0, &lens);
    out.off.assign((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) out.off[i + 1] = out.off[i] + lens[i];
    if (trace) fprintf(stderr, "[s4g trace] fasta: %zu bytes, %zu pieces on %zu threads: count=%.1fms alloc=%.1fms fill=%.1fms\n", total, pieces.size(), n_threads,
                       t1 - t0, t2 - t1, now() - t2);
    return S4G_OK;
}

// ---- packed file ------------------------------------------------------------------------------------
// little endian:  Header (64 B) | int64 offsets[n+1] | int64 name_off[n+1] | names (NUL terminated) | pad to 64 B |
//                 uint8 codes[n_residues]   (the HBM layout: 1 byte per residue, FASTA order)
struct DbFileHeader {
    char magic[8];
    uint64_t n_seqs, n_residues, names_bytes, codes_pos;
    uint64_t reserved[3];
};
static_assert(sizeof(DbFileHeader) == 64, "header is 64 bytes");
const char kMagic[8] = {'S', '4', 'G', 'D', 'B', 0, 0, 1};

bool read_at(int fd, void* dst, size_t bytes, uint64_t pos) {
    char* p = (char*)dst;
    while (bytes) {
        ssize_t got = pread(fd, p, std::min<size_t>(bytes, 1u << 30), (off_t)pos);
        if (got <= 0) return false;
        p += got; pos += (uint64_t)got; bytes -= (size_t)got;
    }
    return true;
}

bool read_header(int fd, DbFileHeader& h) {
    return read_at(fd, &h, sizeof(h), 0) && memcmp(h.magic, kMagic, 8) == 0;
}

// true when pred(i) holds for every i in [0, n): checked on several host threads
template <class F>
static bool parallel_all(int64_t n, F pred) {
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((int64_t)std::thread::hardware_concurrency(), 16), n / (1 << 20) + 1));
    std::atomic<bool> ok(true);
    auto work = [&](int t) {
        const int64_t a = n * t / n_threads, b = n * (t + 1) / n_threads;
        bool mine = true;
        for (int64_t i = a; i < b && mine; ++i) mine = pred(i);
        if (!mine) ok = false;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return ok;
}

}  // namespace

extern "C" {

int s4g_db_open_fasta(s4g_ctx* ctx, const char* path, int shard, int n_shards, s4g_db** out) {
    if (!ctx || !path || !out || n_shards < 1 || shard < 0 || shard >= n_shards) return S4G_ERR_ARG;
    *out = nullptr;
    try {
    Parsed p;
    int rc = parse_fasta(ctx, path, p);
    if (rc != S4G_OK) return rc;
    const int64_t n_all = (int64_t)p.names.size();
    const int64_t lo = n_all * shard / n_shards, hi = n_all * (shard + 1) / n_shards;
    std::vector<int64_t> soff(hi - lo + 1);
    for (int64_t i = lo; i <= hi; ++i) soff[i - lo] = p.off[i] - p.off[lo];
    rc = s4g_db_create(ctx, p.codes.data() + p.off[lo], soff.data(), hi - lo, (uint32_t)lo, S4G_HOST, out);
    if (rc == S4G_OK) {
        (*out)->names.resize(hi - lo);
        for (int64_t i = lo; i < hi; ++i) (*out)->names[i - lo].swap(p.names[i]);
        (*out)->total_seqs = n_all;
        (*out)->total_residues = (uint64_t)p.off[n_all];
    }
    return rc;
    } catch (const std::exception& e) {          // nothing is thrown across the C ABI
        if (*out) { s4g_db_close(*out); *out = nullptr; }
        s4g_set_error(ctx, "'%s': %s while reading the FASTA file", path, e.what());
        return S4G_ERR_NOMEM;
    }
}

int s4g_db_pack_fasta(const char* fasta_path, const char* out_path) {
    if (!fasta_path || !out_path) return S4G_ERR_ARG;
    try {
    Parsed p;
    int rc = parse_fasta(nullptr, fasta_path, p);
    if (rc != S4G_OK) return rc;
    const uint64_t n = p.names.size();
    std::vector<int64_t> name_off(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) name_off[i + 1] = name_off[i] + (int64_t)p.names[i].size() + 1;
    DbFileHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 8);
    h.n_seqs = n; h.n_residues = (uint64_t)p.off[n]; h.names_bytes = (uint64_t)name_off[n];
    uint64_t pos = sizeof(h) + 2 * sizeof(int64_t) * (n + 1) + h.names_bytes;
    h.codes_pos = (pos + 63) / 64 * 64;
    FILE* f = fopen(out_path, "wb");
    if (!f) { s4g_set_error(nullptr, "cannot create '%s'", out_path); return S4G_ERR_IO; }
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && fwrite(p.off.data(), sizeof(int64_t), n + 1, f) == n + 1;
    ok = ok && fwrite(name_off.data(), sizeof(int64_t), n + 1, f) == n + 1;
    for (uint64_t i = 0; ok && i < n; ++i) ok = fwrite(p.names[i].c_str(), 1, p.names[i].size() + 1, f) == p.names[i].size() + 1;
    const char zeros[64] = {0};
    ok = ok && (h.codes_pos == pos || fwrite(zeros, 1, h.codes_pos - pos, f) == h.codes_pos - pos);
    ok = ok && (h.n_residues == 0 || fwrite(p.codes.data(), 1, h.n_residues, f) == h.n_residues);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { s4g_set_error(nullptr, "write to '%s' failed", out_path); return S4G_ERR_IO; }
    return S4G_OK;
    } catch (const std::exception& e) {
        s4g_set_error(nullptr, "'%s': %s while packing", fasta_path, e.what());
        return S4G_ERR_NOMEM;
    }
}

int s4g_db_file_info(const char* path, int64_t* n_seqs, uint64_t* n_residues) {
    if (!path) return S4G_ERR_ARG;
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) { s4g_set_error(nullptr, "cannot open '%s'", path); return S4G_ERR_IO; }
    DbFileHeader h;
    const bool ok = read_header(fd, h);
    ::close(fd);
    if (!ok) { s4g_set_error(nullptr, "'%s' is not a packed sift4g_b200 database", path); return S4G_ERR_IO; }
    if (n_seqs) *n_seqs = (int64_t)h.n_seqs;
    if (n_residues) *n_residues = h.n_residues;
    return S4G_OK;
}

int s4g_db_open_packed(s4g_ctx* ctx, const char* path, int shard, int n_shards, s4g_db** out) {
    if (!ctx || !path || !out || n_shards < 1 || shard < 0 || shard >= n_shards) return S4G_ERR_ARG;
    *out = nullptr;
    try {
        // the file is the HBM layout: it is mapped, checked in place on the host cores and copied once (H2D + the host mirror)
        MappedFile f;
        if (!f.open(path)) { s4g_set_error(ctx, "cannot open '%s'", path); return S4G_ERR_IO; }
        DbFileHeader h;
        if (f.size < sizeof(h) || (memcpy(&h, f.p, sizeof(h)), memcmp(h.magic, kMagic, 8) != 0)) {
            s4g_set_error(ctx, "'%s' is not a packed sift4g_b200 database", path);
            return S4G_ERR_IO;
        }
        // Nothing of the header is trusted before it is checked against the size of the file: the tables, the names and the
        // codes must all lie inside it (a corrupt n_seqs would otherwise size the loops below).
        const uint64_t fsize = f.size;
        const uint64_t tables = sizeof(h) + 2 * sizeof(int64_t) * (h.n_seqs + 1);
        const bool header_ok = h.n_seqs < ((uint64_t)1 << 32) && tables <= fsize && h.names_bytes <= fsize - tables &&
                               h.codes_pos >= tables + h.names_bytes && h.codes_pos <= fsize && h.n_residues <= fsize - h.codes_pos;
        if (!header_ok) { s4g_set_error(ctx, "'%s': corrupt packed database header", path); return S4G_ERR_IO; }
        const int64_t n_all = (int64_t)h.n_seqs;
        const int64_t lo = n_all * shard / n_shards, hi = n_all * (shard + 1) / n_shards, n = hi - lo;
        const int64_t* off = reinterpret_cast<const int64_t*>(f.p + sizeof(h)) + lo;
        const int64_t* name_off = reinterpret_cast<const int64_t*>(f.p + sizeof(h) + sizeof(int64_t) * (n_all + 1)) + lo;
        const char* names = f.p + tables;
        const uint8_t* codes = reinterpret_cast<const uint8_t*>(f.p + h.codes_pos);
        bool ok = off[0] >= 0 && (uint64_t)off[n] <= h.n_residues && name_off[0] >= 0 && (uint64_t)name_off[n] <= h.names_bytes;
        // every entry, not only the ends: sequences are non-empty, a name holds at least its NUL and ends with it
        ok = ok && parallel_all(n, [&](int64_t i) { return off[i + 1] > off[i] && name_off[i + 1] > name_off[i] && (uint64_t)name_off[i + 1] <= h.names_bytes &&
                                                           (uint64_t)off[i + 1] <= h.n_residues; });
        ok = ok && parallel_all(n, [&](int64_t i) { return names[name_off[i + 1] - 1] == 0; });
        if (!ok) { s4g_set_error(ctx, "'%s': truncated or corrupt packed database", path); return S4G_ERR_IO; }
        const uint8_t* my_codes = codes + off[0];
        const int64_t my_res = off[n] - off[0];
        // 8 codes per step: any byte above 25 has bit 5, 6 or 7 set, or is 26..31 (bits 3+4 and one of bits 1, 2)
        const int64_t words = my_res / 8;
        bool codes_ok = parallel_all(words, [&](int64_t w) {
            uint64_t x;
            memcpy(&x, my_codes + 8 * w, 8);
            const uint64_t hi3 = x & 0xE0E0E0E0E0E0E0E0ull;
            const uint64_t big = (x & 0x1818181818181818ull) == 0 ? 0 : ((x >> 3) & (x >> 4) & ((x >> 1) | (x >> 2)) & 0x0101010101010101ull);
            return (hi3 | big) == 0;
        });
        for (int64_t i = 8 * words; codes_ok && i < my_res; ++i) codes_ok = my_codes[i] < S4G_NLET;
        if (!codes_ok) { s4g_set_error(ctx, "'%s': residue code out of range", path); return S4G_ERR_IO; }
        std::vector<int64_t> loc(n + 1);
        for (int64_t i = 0; i <= n; ++i) loc[i] = off[i] - off[0];
        int rc = s4g_db_create(ctx, my_codes, loc.data(), n, (uint32_t)lo, S4G_HOST, out);
        if (rc != S4G_OK) return rc;
        (*out)->names.resize(n);
        parallel_all(n, [&](int64_t i) { (*out)->names[i].assign(names + name_off[i]); return true; });
        (*out)->total_seqs = n_all;
        (*out)->total_residues = h.n_residues;
        return S4G_OK;
    } catch (const std::exception& e) {          // nothing is thrown across the C ABI
        if (*out) { s4g_db_close(*out); *out = nullptr; }
        s4g_set_error(ctx, "'%s': %s while reading the packed database", path, e.what());
        return S4G_ERR_NOMEM;
    }
}

static bool is_packed_file(const char* path, bool* opened) {
    char magic[8] = {0};
    FILE* f = fopen(path, "rb");
    *opened = f != nullptr;
    if (!f) return false;
    const size_t got = fread(magic, 1, 8, f);
    fclose(f);
    return got == 8 && memcmp(magic, kMagic, 8) == 0;
}

int s4g_db_open_sharded(s4g_ctx* const* ctxs, int n_ctx, const char* path, int n_threads, s4g_db** out) {
    if (!ctxs || n_ctx < 1 || !path || !out) return S4G_ERR_ARG;
    for (int d = 0; d < n_ctx; ++d) { if (!ctxs[d]) return S4G_ERR_ARG; out[d] = nullptr; }
    bool opened = false;
    const bool packed = is_packed_file(path, &opened);
    if (!opened) { s4g_set_error(ctxs[0], "cannot open '%s'", path); return S4G_ERR_IO; }
    std::vector<int> rcs(n_ctx, S4G_OK);
    try {
        Parsed p;
        int64_t n_all = 0;
        if (!packed) {
            const int rc = parse_fasta(ctxs[0], path, p, n_threads);
            if (rc != S4G_OK) return rc;
            n_all = (int64_t)p.names.size();
        }
        // one host thread per GPU: the H2D copies of the shards (and, for a packed file, the reads of their byte ranges) overlap
        auto open_one = [&](int d) {
            if (packed) { rcs[d] = s4g_db_open_packed(ctxs[d], path, d, n_ctx, &out[d]); return; }
            const int64_t lo = n_all * d / n_ctx, hi = n_all * (d + 1) / n_ctx;
            std::vector<int64_t> soff(hi - lo + 1);
            for (int64_t i = lo; i <= hi; ++i) soff[i - lo] = p.off[i] - p.off[lo];
            rcs[d] = s4g_db_create(ctxs[d], p.codes.data() + p.off[lo], soff.data(), hi - lo, (uint32_t)lo, S4G_HOST, &out[d]);
            if (rcs[d] != S4G_OK) return;
            out[d]->names.resize(hi - lo);
            for (int64_t i = lo; i < hi; ++i) out[d]->names[i - lo].swap(p.names[i]);
            out[d]->total_seqs = n_all;
            out[d]->total_residues = (uint64_t)p.off[n_all];
        };
        if (n_ctx == 1) open_one(0);
        else {
            std::vector<std::thread> th;
            for (int d = 0; d < n_ctx; ++d) th.emplace_back(open_one, d);
            for (auto& t : th) t.join();
        }
    } catch (const std::exception& e) {
        s4g_set_error(ctxs[0], "'%s': %s while opening the database", path, e.what());
        rcs[0] = S4G_ERR_NOMEM;
    }
    for (int d = 0; d < n_ctx; ++d)
        if (rcs[d] != S4G_OK) {
            const std::string msg = ctxs[d]->err;           // closing the other shards must not lose the text
            for (int x = 0; x < n_ctx; ++x) if (out[x]) { s4g_db_close(out[x]); out[x] = nullptr; }
            s4g_set_error(ctxs[0], "%s", msg.c_str());
            return rcs[d];
        }
    return S4G_OK;
}

int s4g_db_open(s4g_ctx* ctx, const char* path, int shard, int n_shards, s4g_db** out) {
    if (!ctx || !path || !out) return S4G_ERR_ARG;
    char magic[8] = {0};
    FILE* f = fopen(path, "rb");
    if (!f) { *out = nullptr; s4g_set_error(ctx, "cannot open '%s'", path); return S4G_ERR_IO; }
    const size_t got = fread(magic, 1, 8, f);
    fclose(f);
    if (got == 8 && memcmp(magic, kMagic, 8) == 0) return s4g_db_open_packed(ctx, path, shard, n_shards, out);
    return s4g_db_open_fasta(ctx, path, shard, n_shards, out);
}

}  // extern "C"
