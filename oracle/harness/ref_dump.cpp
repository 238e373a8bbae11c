// TEST INFRASTRUCTURE -- not product code.
//
// Seam-dump harness linked against the UNMODIFIED reference objects built by oracle/Makefile
// (oracle/_ref/libswsharp_ref.a + the sift4g objects).  It calls the reference's own entry points
// and prints what they return, so that the C restatement (oracle/s4g_oracle.c) and the CUDA path
// can be compared against the real thing:
//
//   ref_dump candidates Q.fa DB.fa k max_candidates threads
//        -> searchDatabase()                      sift4g/src/database_search.cpp:66
//   ref_dump scores Q.fa DB.fa CANDS.txt
//        -> scoreDatabaseCpu() (swimd)            vendor/swsharp/swsharp/src/cpu_module.c:179
//   ref_dump align Q.fa DB.fa PAIRS.txt [gap_open gap_extend]
//        -> alignScoredPairCpu() (SSW / swAlign)  vendor/swsharp/swsharp/src/cpu_module.c:111
//   ref_dump pipeline Q.fa DB.fa k max_candidates threads max_evalue max_alignments
//        -> searchDatabase() + alignDatabase()    sift4g/src/main.cpp:203-220
//   ref_dump matrix
//        -> scorerCreateMatrix("BLOSUM_62") table   vendor/swsharp/swsharp/src/pre_proc.c:399, constants.c:87-114
//   ref_dump fasta FILE.fa
//        -> readFastaChains(): per record "name<TAB>residues" (what the reader kept)   vendor/swsharp/swsharp/src/pre_proc.c:437-538
//   ref_dump scorebench Q.fa DB.fa CANDS.txt threads
//        -> the reference's threaded scoring loop (tasks of 1000 targets, database.c:896-996) timed
//
// Text formats (all whitespace separated):
//   CANDS.txt : per query one line  "<n> id id id ..."          (ids = FASTA-order indices)
//   PAIRS.txt : per line "<query idx> <target idx> <score>"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "database_alignment.hpp"
#include "database_search.hpp"
#include "swsharp/evalue.h"
#include "swsharp/swsharp.h"

extern "C" void scoreDatabaseCpu(int* scores, int type, Chain* query, Chain** database,
                                 int databaseLen, Scorer* scorer);
extern "C" void alignScoredPairCpu(Alignment** alignment, int type, Chain* query, Chain* target,
                                   Scorer* scorer, int score);

static double now_s() {
    using namespace std::chrono;
    return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}

static std::vector<std::vector<uint32_t>> read_cands(const char* path, int nq) {
    std::vector<std::vector<uint32_t>> c(nq);
    FILE* f = fopen(path, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    for (int q = 0; q < nq; ++q) {
        long n = 0;
        if (fscanf(f, "%ld", &n) != 1) { fprintf(stderr, "short candidate file\n"); exit(2); }
        c[q].resize(n);
        for (long i = 0; i < n; ++i) {
            unsigned v; if (fscanf(f, "%u", &v) != 1) exit(2);
            c[q][i] = v;
        }
    }
    fclose(f);
    return c;
}

static int cmd_candidates(int argc, char** argv) {
    if (argc < 7) return 2;
    int k = atoi(argv[4]), maxc = atoi(argv[5]), threads = atoi(argv[6]);
    threadPoolInitialize(threads);
    Chain** queries = nullptr; int nq = 0;
    readFastaChains(&queries, &nq, argv[2]);
    std::vector<std::vector<uint32_t>> idx;
    double t0 = now_s();
    uint64_t cells = searchDatabase(idx, argv[3], queries, nq, k, maxc, threads);
    double t1 = now_s();
    printf("cells %llu\n", (unsigned long long)cells);
    fprintf(stderr, "search_seconds %.6f\n", t1 - t0);
    for (int q = 0; q < nq; ++q) {
        printf("%zu", idx[q].size());
        for (uint32_t id : idx[q]) printf(" %u", id);
        printf("\n");
    }
    deleteFastaChains(queries, nq);
    threadPoolTerminate();
    return 0;
}

static int cmd_scores(int argc, char** argv) {
    if (argc < 5) return 2;
    int go = argc > 6 ? atoi(argv[5]) : 10, ge = argc > 6 ? atoi(argv[6]) : 1;
    threadPoolInitialize(1);
    Chain** queries = nullptr; int nq = 0;
    Chain** db = nullptr; int nd = 0;
    readFastaChains(&queries, &nq, argv[2]);
    readFastaChains(&db, &nd, argv[3]);
    auto cands = read_cands(argv[4], nq);
    Scorer* scorer = nullptr;
    scorerCreateMatrix(&scorer, (char*)"BLOSUM_62", go, ge);
    for (int q = 0; q < nq; ++q) {
        size_t n = cands[q].size();
        std::vector<Chain*> tg(n);
        for (size_t i = 0; i < n; ++i) tg[i] = db[cands[q][i]];
        std::vector<int> sc(n);
        if (n) scoreDatabaseCpu(sc.data(), SW_ALIGN, queries[q], tg.data(), (int)n, scorer);
        printf("%zu", n);
        for (size_t i = 0; i < n; ++i) printf(" %d", sc[i]);
        printf("\n");
    }
    threadPoolTerminate();
    return 0;
}

static void print_alignment(Alignment* a) {
    int n = alignmentGetPathLen(a);
    printf("%d %d %d %d %d %d ", alignmentGetQueryStart(a), alignmentGetQueryEnd(a),
           alignmentGetTargetStart(a), alignmentGetTargetEnd(a), alignmentGetScore(a), n);
    for (int i = 0; i < n; ++i) putchar('0' + alignmentGetMove(a, i));
    if (n == 0) putchar('-');
    putchar('\n');
}

static int cmd_align(int argc, char** argv) {
    if (argc < 5) return 2;
    int go = argc > 6 ? atoi(argv[5]) : 10, ge = argc > 6 ? atoi(argv[6]) : 1;
    threadPoolInitialize(1);
    Chain** queries = nullptr; int nq = 0;
    Chain** db = nullptr; int nd = 0;
    readFastaChains(&queries, &nq, argv[2]);
    readFastaChains(&db, &nd, argv[3]);
    Scorer* scorer = nullptr;
    scorerCreateMatrix(&scorer, (char*)"BLOSUM_62", go, ge);
    FILE* f = fopen(argv[4], "r");
    if (!f) return 2;
    int q, t, s;
    while (fscanf(f, "%d %d %d", &q, &t, &s) == 3) {
        if (s < 0) {   // score unknown: ask the reference scorer first (same call order as database.c)
            Chain* tg = db[t];
            scoreDatabaseCpu(&s, SW_ALIGN, queries[q], &tg, 1, scorer);
        }
        Alignment* a = nullptr;
        alignScoredPairCpu(&a, SW_ALIGN, queries[q], db[t], scorer, s);
        print_alignment(a);
        alignmentDelete(a);
    }
    fclose(f);
    threadPoolTerminate();
    return 0;
}

static int cmd_pipeline(int argc, char** argv) {
    if (argc < 9) return 2;
    int k = atoi(argv[4]), maxc = atoi(argv[5]), threads = atoi(argv[6]);
    double max_evalue = atof(argv[7]);
    int max_alignments = atoi(argv[8]);
    bool quiet = argc > 9 && strcmp(argv[9], "quiet") == 0;
    threadPoolInitialize(threads);
    Chain** queries = nullptr; int nq = 0;
    readFastaChains(&queries, &nq, argv[2]);
    std::vector<std::vector<uint32_t>> idx;
    double t0 = now_s();
    uint64_t cells = searchDatabase(idx, argv[3], queries, nq, k, maxc, threads);
    double t1 = now_s();
    std::vector<std::vector<uint32_t>> idx_copy(idx);
    Scorer* scorer = nullptr;
    scorerCreateMatrix(&scorer, (char*)"BLOSUM_62", 10, 1);
    EValueParams* ep = createEValueParams(cells, scorer);
    DbAlignment*** al = nullptr; int* al_len = nullptr;
    Chain** db = nullptr; int32_t nd = 0;
    double t2 = now_s();
    alignDatabase(&al, &al_len, &db, &nd, argv[3], queries, nq, idx, SW_ALIGN, ep, max_evalue,
                  max_alignments, scorer, nullptr, 0);
    double t3 = now_s();
    // SW cells actually scored = sum_q len(q) * sum_{t in cand(q)} len(t); needs target lengths:
    // re-read the database (the reference frees unused chains).
    Chain** db2 = nullptr; int nd2 = 0;
    readFastaChains(&db2, &nd2, argv[3]);
    unsigned long long sw_cells = 0, pairs = 0;
    for (int q = 0; q < nq; ++q) {
        unsigned long long tl = 0;
        for (uint32_t id : idx_copy[q]) tl += chainGetLength(db2[id]);
        sw_cells += tl * (unsigned long long)chainGetLength(queries[q]);
        pairs += idx_copy[q].size();
    }
    printf("cells %llu\n", (unsigned long long)cells);
    printf("timing search_s %.6f align_s %.6f sw_cells %llu pairs %llu threads %d\n", t1 - t0,
           t3 - t2, sw_cells, pairs, threads);
    if (!quiet) {
        for (int q = 0; q < nq; ++q) {
            printf("query %d cands %zu hits %d\n", q, idx_copy[q].size(), al_len[q]);
            for (int j = 0; j < al_len[q]; ++j) {
                DbAlignment* a = al[q][j];
                // target index: alignDatabase leaves the chunk-local filtered index; recover the
                // FASTA index through the target pointer name instead (names are unique in tests).
                const char* name = chainGetName(dbAlignmentGetTarget(a));
                int n = dbAlignmentGetPathLen(a);
                printf("hit %s %d %a %d %d %d %d %d ", name, dbAlignmentGetScore(a),
                       dbAlignmentGetValue(a), dbAlignmentGetQueryStart(a), dbAlignmentGetQueryEnd(a),
                       dbAlignmentGetTargetStart(a), dbAlignmentGetTargetEnd(a), n);
                for (int i = 0; i < n; ++i) putchar('0' + dbAlignmentGetMove(a, i));
                putchar('\n');
            }
        }
    }
    threadPoolTerminate();
    return 0;
}

struct ScoreTask { int* out; Chain* query; Chain** targets; int n; Scorer* scorer; };
static void* score_task(void* p) {
    ScoreTask* t = (ScoreTask*)p;
    scoreDatabaseCpu(t->out, SW_ALIGN, t->query, t->targets, t->n, t->scorer);
    return nullptr;
}

static int cmd_scorebench(int argc, char** argv) {
    if (argc < 6) return 2;
    int threads = atoi(argv[5]);
    threadPoolInitialize(threads);
    Chain** queries = nullptr; int nq = 0;
    Chain** db = nullptr; int nd = 0;
    readFastaChains(&queries, &nq, argv[2]);
    readFastaChains(&db, &nd, argv[3]);
    auto cands = read_cands(argv[4], nq);
    Scorer* scorer = nullptr;
    scorerCreateMatrix(&scorer, (char*)"BLOSUM_62", 10, 1);
    std::vector<std::vector<Chain*>> tg(nq);
    std::vector<std::vector<int>> sc(nq);
    unsigned long long cells = 0;
    for (int q = 0; q < nq; ++q) {
        unsigned long long tl = 0;
        for (uint32_t id : cands[q]) { tg[q].push_back(db[id]); tl += chainGetLength(db[id]); }
        sc[q].resize(cands[q].size());
        cells += tl * (unsigned long long)chainGetLength(queries[q]);
    }
    const int chunk = 1000;   // CPU_THREAD_CHUNK, database.c:44
    std::vector<ScoreTask> tasks;
    for (int q = 0; q < nq; ++q)
        for (size_t b = 0; b < tg[q].size(); b += chunk)
            tasks.push_back({sc[q].data() + b, queries[q], tg[q].data() + b,
                             (int)std::min<size_t>(chunk, tg[q].size() - b), scorer});
    double t0 = now_s();
    std::vector<ThreadPoolTask*> h(tasks.size());
    for (size_t i = 0; i < tasks.size(); ++i) h[i] = threadPoolSubmit(score_task, &tasks[i]);
    for (size_t i = 0; i < tasks.size(); ++i) { threadPoolTaskWait(h[i]); threadPoolTaskDelete(h[i]); }
    double t1 = now_s();
    long long checksum = 0;
    for (int q = 0; q < nq; ++q) for (int s : sc[q]) checksum += s;
    printf("scorebench cells %llu seconds %.6f gcups %.4f threads %d checksum %lld\n", cells, t1 - t0,
           cells / (t1 - t0) * 1e-9, threads, checksum);
    threadPoolTerminate();
    return 0;
}

static int cmd_matrix(int argc, char** argv) {
    int go = argc > 3 ? atoi(argv[2]) : 10, ge = argc > 3 ? atoi(argv[3]) : 1;
    Scorer* scorer = nullptr;
    scorerCreateMatrix(&scorer, (char*)"BLOSUM_62", go, ge);
    const int* t = scorerGetTable(scorer);
    int n = scorerGetMaxCode(scorer);
    printf("%d", n);
    for (int i = 0; i < n * n; ++i) printf(" %d", t[i]);
    printf("\n");
    return 0;
}

static int cmd_fasta(int argc, char** argv) {
    if (argc < 3) return 2;
    Chain** chains = NULL;
    int n = 0;
    readFastaChains(&chains, &n, argv[2]);
    printf("%d\n", n);
    for (int i = 0; i < n; ++i) {
        printf("%s\t", chainGetName(chains[i]));
        for (int j = 0; j < chainGetLength(chains[i]); ++j) putchar(chainGetChar(chains[i], j));
        putchar('\n');
    }
    deleteFastaChains(chains, n);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_dump <candidates|scores|align|pipeline|scorebench|matrix|fasta> ...\n"); return 2; }
    std::string c = argv[1];
    if (c == "candidates") return cmd_candidates(argc, argv);
    if (c == "scores") return cmd_scores(argc, argv);
    if (c == "align") return cmd_align(argc, argv);
    if (c == "pipeline") return cmd_pipeline(argc, argv);
    if (c == "scorebench") return cmd_scorebench(argc, argv);
    if (c == "matrix") return cmd_matrix(argc, argv);
    if (c == "fasta") return cmd_fasta(argc, argv);
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 2;
}
