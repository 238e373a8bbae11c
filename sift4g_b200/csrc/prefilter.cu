#include "common.cuh"
int s4g_prefilter_device(s4g_ctx* ctx, s4g_db*, s4g_queries*, int, int, int, uint32_t*, float*, uint32_t*) {
    s4g_set_error(ctx, "prefilter not built yet");
    return S4G_ERR_INTERNAL;
}
extern "C" int s4g_merge_candidates(s4g_ctx* ctx, int, int, int, const uint32_t*, const float*, const uint32_t*, uint32_t*, float*, uint32_t*) {
    s4g_set_error(ctx, "merge not built yet");
    return S4G_ERR_INTERNAL;
}
