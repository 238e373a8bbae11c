// s4g_pack FASTA OUT.s4gdb -- writes the packed database file (include/sift4g_b200.h, s4g_db_pack_fasta) that
// sift4g_b200 -d accepts in place of the FASTA.  The role of the reference's ".swsharp" cache, which its reader writes
// next to the FASTA on first use (vendor/swsharp/swsharp/src/pre_proc.c:309-374); here it is an explicit step.
#include <cstdio>

#include "sift4g_b200.h"

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s <database.fasta> <out.s4gdb>\n", argv[0]);
        return 2;
    }
    const int rc = s4g_db_pack_fasta(argv[1], argv[2]);
    if (rc != S4G_OK) {
        fprintf(stderr, "[ERROR:s4g_pack] %s (%d)\n", s4g_last_error(nullptr), rc);
        return 1;
    }
    int64_t n = 0;
    uint64_t residues = 0;
    if (s4g_db_file_info(argv[2], &n, &residues) == S4G_OK)
        fprintf(stderr, "%s: %lld sequences, %llu residues\n", argv[2], (long long)n, (unsigned long long)residues);
    return 0;
}
