// Entry points declared in include/daisy_b200.h whose kernels are not built yet.  They fail loudly
// (DAISY_EUNSUPPORTED with a message) -- there is no fallback path behind them.
#include "ctx.cuh"

extern "C" int daisy_topk_full(daisy_handle_t, const float *, const float *, const int32_t *, int64_t, int,
                               const int64_t *, const int32_t *, int32_t *, float *, daisy_stream_t) {
    daisy_set_error("daisy_topk_full: kernel not built in this revision");
    return DAISY_EUNSUPPORTED;
}
