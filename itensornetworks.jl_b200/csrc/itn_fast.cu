// DMMA fast path for synchronous BP sweeps (placeholder until the tensor-core kernels land).
#include "itn_internal.h"

bool itn_fast_bp_supported(itn_net*, const std::vector<int>&) { return false; }
void itn_fast_bp_sweep(itn_net*, const std::vector<int>&, double**) {
  throw ItnError(ITN_EUNSUPPORTED, "fast path not built");
}
void itn_fast_release(itn_net*) {}
