// bottomup_decode.cu -- placeholder until the bottom-up kernels land.
#include "common.cuh"
using namespace pc;
extern "C" int pc_bottomup_decode(const float*, const float*, const uint8_t*, float*, float*,
                                  float*, float*, float*, const pc_bottomup_decode_params*,
                                  int64_t, void*) {
  set_error("pc_bottomup_decode: not implemented yet");
  return PC_ERR_UNSUPPORTED;
}
