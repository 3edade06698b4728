// Batched small dense linear algebra for simple update (placeholder).
#include "itn_internal.h"

extern "C" int itn_apply2(itn_net*, const int32_t*, int, const void*, int, double, int, int, int32_t*, double*,
                          double*, int) {
  itn_set_error("itn_apply2: not implemented yet");
  return ITN_EUNSUPPORTED;
}
extern "C" int itn_map_eigvals(itn_ctx*, int, int, int, int, const void*, void*, double) {
  itn_set_error("itn_map_eigvals: not implemented yet");
  return ITN_EUNSUPPORTED;
}
extern "C" int itn_nccl_unique_id(void*) {
  itn_set_error("itn_nccl_unique_id: not implemented yet");
  return ITN_EUNSUPPORTED;
}
extern "C" int itn_ctx_init_dist(itn_ctx*, int, int, const void*) {
  itn_set_error("itn_ctx_init_dist: not implemented yet");
  return ITN_EUNSUPPORTED;
}
