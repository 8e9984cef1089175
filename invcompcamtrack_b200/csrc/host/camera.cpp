#include "camera.h"

namespace CTR {

// The per-level table comes from the library (ict_camera_levels == camera.cpp:32-43 of the reference), so host and
// device use one definition of the intrinsics.
CamClass::CamClass(const int noscales_in, const float* fc_in, const float* cc_in, const int* wh_in, const int padding_in)
    : noscales(noscales_in), padding(padding_in) {
  for (int k = 0; k < 2; ++k) { fc_org[k] = fc_in[k]; cc_org[k] = cc_in[k]; wh_org[k] = wh_in[k]; }
  std::memset(lv, 0, sizeof(lv));
  ict_camera_levels(noscales, fc_org, cc_org, wh_org, padding, lv);
}

}  // namespace CTR
