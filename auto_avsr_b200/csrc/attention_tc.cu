// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace avsr {
int attention_tc(const float*, const float*, const float*, const float*, const float*, const int32_t*, float*, int,
                 int, int, int, int, int, cudaStream_t) {
  set_error("attention_tc: tcgen05 attention not built yet");
  return AVSR_E_INVALID;
}
}  // namespace avsr
