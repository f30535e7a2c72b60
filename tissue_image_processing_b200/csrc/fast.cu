// placeholder until the multirate score stage lands
#include "common.cuh"
namespace tsp {
size_t fast_workspace_bytes(int, int, int) { return 256; }
int launch_fast_score_argmax(tsp_handle*, const uint16_t*, int32_t*, int, int, int, int, int, int32_t*, void*,
                             cudaStream_t) {
    set_error("fast mode not built yet");
    return TSP_ERR_INVALID;
}
int tsp_debug_coarse_taps_impl(double*, int) { return TSP_ERR_INVALID; }
}  // namespace tsp
extern "C" int tsp_debug_coarse_taps(double* out, int capacity) { return tsp::tsp_debug_coarse_taps_impl(out, capacity); }
