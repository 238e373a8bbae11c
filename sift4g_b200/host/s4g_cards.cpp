// cudaGetCards / cudaCheckCards of the CLI (declared in sw/cuda_utils.h:73-83, called from sift4g/src/main.cpp:183-186): the
// reference's CPU build answers "no cards" (sw/cuda_utils.cu:30-62); here they enumerate / validate the GPUs the C-ABI
// library can drive, so that `--cards 01` selects devices exactly like the reference's `make gpu` build did.
#include <cstdlib>

#include "sift4g_b200.h"

extern "C" {

void cudaGetCards(int** cards, int* cardsLen) {
    const int n = s4g_device_count();
    *cardsLen = n;
    *cards = (int*)malloc((n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) (*cards)[i] = i;
}

int cudaCheckCards(int* cards, int cardsLen) {
    const int n = s4g_device_count();
    for (int i = 0; i < cardsLen; ++i) if (cards[i] < 0 || cards[i] >= n) return 0;
    return 1;
}

size_t cudaMinimalGlobalMemory(int*, int) { return 0; }      // not used by the parts of the reference this CLI links

}
