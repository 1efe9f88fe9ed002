// placeholder: tcgen05 path lands next
#include "infonce.cuh"
namespace rmcl {
bool infonce_tc_built() { return false; }
int infonce_tc_launch(const __nv_bfloat16*, const void*, int, int, long long, long long, float, const InfoNcePlan&,
                      InfoNcePartials, cudaStream_t) {
  set_error("tcgen05 InfoNCE not built");
  return RMCL_E_UNSUPPORTED_DIM;
}
}  // namespace rmcl
