// Process-wide session shared by the two reference seams (searchDatabase / alignDatabase):
// one C-ABI context per process, the database shard opened once and kept resident in HBM between the
// prefilter and the alignment stage (the reference parses the FASTA twice:
// sift4g/src/database_search.cpp:81-97, database_alignment.cpp:36-48).
#pragma once

#include <string>

#include "sift4g_b200.h"

struct S4gSession {
    s4g_ctx* ctx = nullptr;
    s4g_db* db = nullptr;
    std::string db_path;
    s4g_queries* queries = nullptr;
    const void* queries_key = nullptr;   // Chain** the batch was built from
    int queries_n = 0;
};

S4gSession& s4gSession();
// exits like the reference's ASSERT (sift4g/src/utils.hpp:13-19) when rc != S4G_OK
void s4gCheck(int rc, const char* what);
void s4gOpenDatabase(const std::string& path);
struct Chain;
void s4gUploadQueries(Chain** queries, int queries_length);
