// pose.h — PoseClass with the reference's interface (pose.h:15-61).  Holds the current pose on the host; during
// OdometerClass::TrackPose the pose lives on the device and is written back here when the call returns.
#ifndef ICT_HOST_POSE_HEADER
#define ICT_HOST_POSE_HEADER

#include "camera.h"
#include "utilities.h"

namespace CTR {

class PoseClass {
 public:
  PoseClass(const CamClass* camobj_in, const optparam* op_in);
  ~PoseClass() {}

  void setpose_se3(const double* p_in, const Eigen::Vector3d meanshift_in, const double varval_in);
  void addpose_se3(const float* p_in);
  void subpose_se3(const float* p_in);
  void getPose_se3(double* p_out) const;
  void project_pt(const float* pt3d, float* pt2d, int nopoints, int sc) const;
  void project_pt_save_rotated(const float* pt3d, float* pt3d_rot, float* pt2d, int nopoints, int sc) const;

  const CamClass* camobj;

 private:
  void project(const float* pt3d, float* pt3d_rot, float* pt2d, int nopoints, int sc) const;
  Eigen::Vector3d meanshift;
  double varval;
  float cpos_G[3 * 4];
  float cpos_p[6];
  const optparam* op;
};

}  // namespace CTR
#endif
