// camera.h — CamClass with the reference's interface (camera.h:16-56): per-level intrinsics and image extents.
#ifndef ICT_HOST_CAM_HEADER
#define ICT_HOST_CAM_HEADER

#include "utilities.h"

namespace CTR {

class CamClass {
 public:
  CamClass(const int noscales_in, const float* fc_in, const float* cc_in, const int* wh_in, const int padding_in);
  ~CamClass() {}

  inline float getfx(int sc) const { return lv[8 * sc + 0]; }
  inline float getfy(int sc) const { return lv[8 * sc + 1]; }
  inline float getcx(int sc) const { return lv[8 * sc + 2]; }
  inline float getcy(int sc) const { return lv[8 * sc + 3]; }
  inline float getswo(int sc) const { return lv[8 * sc + 4]; }
  inline float getsho(int sc) const { return lv[8 * sc + 5]; }
  inline float getsw(int sc) const { return lv[8 * sc + 6]; }
  inline float getsh(int sc) const { return lv[8 * sc + 7]; }

  // what the device-side tracker is created from (not in the reference's interface)
  inline const float* fc() const { return fc_org; }
  inline const float* cc() const { return cc_org; }
  inline const int* wh() const { return wh_org; }
  inline int getpadding() const { return padding; }
  inline int getnoscales() const { return noscales; }

 private:
  const int noscales;
  float fc_org[2], cc_org[2];
  int wh_org[2];
  const int padding;
  float lv[8 * ICT_MAX_LEVELS];   // fx fy cx cy swo sho sw sh per level
};

}  // namespace CTR
#endif
