// The reference data-model objects reference a few engine entry points that the B200 build never calls
// (checkAlignment -> scorePair etc.).  They are defined here as hard failures so that no reference
// Smith-Waterman engine is linked into the product binary.
#include <cstdio>
#include <cstdlib>

