// placeholder until the tcgen05 scoring kernel lands (next commit)
#include "common.cuh"
using namespace ttam;
extern "C" int64_t ttam_topk_bf16_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 256; }
extern "C" int ttam_topk_bf16(const uint16_t*, const uint16_t*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t*,
                              float*, void*, int64_t, void*) {
  set_error("topk_bf16: not built yet");
  return TTAM_EUNSUPPORTED;
}
