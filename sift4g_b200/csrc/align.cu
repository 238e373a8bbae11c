#include "common.cuh"
extern "C" int s4g_sw_align(s4g_ctx* ctx, s4g_db*, s4g_queries*, int64_t, const uint32_t*, const uint32_t*, const int32_t*,
                            const int32_t*, int, int, int32_t*, uint8_t*, int64_t, int64_t*, int) {
    s4g_set_error(ctx, "align not built yet");
    return S4G_ERR_INTERNAL;
}
