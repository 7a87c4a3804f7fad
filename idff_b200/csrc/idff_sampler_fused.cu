// placeholder until the persistent kernel lands
#include "idff_common.cuh"
namespace idff {
bool fused_supports(const idff_weights_t*, const idff_field_params_t*) { return false; }
int fused_sample_rk4(const idff_weights_t*, const idff_field_params_t*, const float*, const Stage*, const float*, int,
                     float*, float*, int64_t, void*) { set_error("fused kernel not built"); return IDFF_EUNSUPPORTED; }
int fused_field_eval(const idff_weights_t*, const idff_field_params_t*, float, const float*, float*, int64_t, void*) {
  set_error("fused kernel not built"); return IDFF_EUNSUPPORTED; }
}
