// grouping.cu -- placeholder until the grouping kernel lands.
#include "common.cuh"
using namespace pc;
extern "C" int pc_group_by_tag(const float*, const float*, const float*, float*, int32_t*,
                               float*, const pc_group_params*, int64_t, void*) {
  set_error("pc_group_by_tag: not implemented yet");
  return PC_ERR_UNSUPPORTED;
}
extern "C" int pc_transform_keypoints(float*, const int32_t*, const double*, const double*,
                                      const double*, float, int32_t, int64_t, void*) {
  set_error("pc_transform_keypoints: not implemented yet");
  return PC_ERR_UNSUPPORTED;
}
