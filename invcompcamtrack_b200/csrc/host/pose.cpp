// pose.cpp — host bookkeeping of the pose (pose.cpp:25-129, 307-488 of the reference restated as plain scalar
// loops; the reference's SSE point projection is what the device kernels do for the tracker itself).
#include "pose.h"

namespace CTR {

PoseClass::PoseClass(const CamClass* camobj_in, const optparam* op_in) : camobj(camobj_in), varval(0), op(op_in) {
  std::memset(cpos_G, 0, sizeof(cpos_G));
  std::memset(cpos_p, 0, sizeof(cpos_p));
  cpos_G[0] = cpos_G[5] = cpos_G[10] = 1.0f;
}

namespace {
// camera centre of [R|t]: -R^T t
template <typename T>
void centre_of(const T* G, double* c) {
  for (int k = 0; k < 3; ++k) c[k] = (double)(T)(-G[k] * G[3] - G[4 + k] * G[7] - G[8 + k] * G[11]);
}
}  // namespace

void PoseClass::setpose_se3(const double* p_in, const Eigen::Vector3d meanshift_in, const double varval_in) {
  double p[6];
  std::memcpy(p, p_in, sizeof(p));
  if (op->donorm) {   // move the camera centre into the normalised point frame (pose.cpp:31-61)
    varval = varval_in;
    meanshift = meanshift_in;
    double G[12], c[3];
    util_SE3_coeff_to_group(G, p);
    centre_of(G, c);
    for (int k = 0; k < 3; ++k) c[k] = (c[k] - meanshift[k]) / varval;
    for (int r = 0; r < 3; ++r) G[4 * r + 3] = -G[4 * r] * c[0] - G[4 * r + 1] * c[1] - G[4 * r + 2] * c[2];
    util_SE3_group_to_coeff(p, G);
  }
  for (int k = 0; k < 6; ++k) cpos_p[k] = static_cast<float>(p[k]);
  util_SE3_coeff_to_group(cpos_G, cpos_p);
}

void PoseClass::getPose_se3(double* p_out) const {
  float p[6];
  std::memcpy(p, cpos_p, sizeof(p));
  if (op->donorm) {   // pose.cpp:85-105
    float G[12];
    std::memcpy(G, cpos_G, sizeof(G));
    double c[3];
    centre_of(G, c);
    for (int k = 0; k < 3; ++k) c[k] = c[k] * varval + meanshift[k];
    for (int r = 0; r < 3; ++r)
      G[4 * r + 3] = (float)(-(double)G[4 * r] * c[0] - (double)G[4 * r + 1] * c[1] - (double)G[4 * r + 2] * c[2]);
    util_SE3_group_to_coeff(p, G);
  }
  for (int k = 0; k < 6; ++k) p_out[k] = static_cast<double>(p[k]);
}

void PoseClass::addpose_se3(const float* p_in) {   // additive in coefficient space, then exp (pose.cpp:116-129)
  for (int k = 0; k < 6; ++k) cpos_p[k] += p_in[k];
  util_SE3_coeff_to_group(cpos_G, cpos_p);
}

void PoseClass::subpose_se3(const float* p_in) {
  for (int k = 0; k < 6; ++k) cpos_p[k] -= p_in[k];
  util_SE3_coeff_to_group(cpos_G, cpos_p);
}

void PoseClass::project(const float* pt3d, float* pt3d_rot, float* pt2d, int nopoints, int sc) const {
  const int M = op->maxpttrack;
  const float fx = camobj->getfx(sc), fy = camobj->getfy(sc), cx = camobj->getcx(sc), cy = camobj->getcy(sc);
  const float* G = cpos_G;
  if (nopoints % SSEMULTIPL) nopoints += SSEMULTIPL - nopoints % SSEMULTIPL;   // pose.cpp:327-329
  for (int i = 0; i < nopoints; ++i) {
    const float X = pt3d[i], Y = pt3d[i + M], Z = pt3d[i + 2 * M];
    const float a = G[0] * X + G[1] * Y + G[2] * Z + G[3];
    const float b = G[4] * X + G[5] * Y + G[6] * Z + G[7];
    const float c = G[8] * X + G[9] * Y + G[10] * Z + G[11];
    if (pt3d_rot) { pt3d_rot[i] = a; pt3d_rot[i + M] = b; pt3d_rot[i + 2 * M] = c; }
    pt2d[i] = (a / c) * fx + cx;
    pt2d[i + M] = (b / c) * fy + cy;
  }
}

void PoseClass::project_pt(const float* pt3d, float* pt2d, int nopoints, int sc) const {
  project(pt3d, nullptr, pt2d, nopoints, sc);
}

void PoseClass::project_pt_save_rotated(const float* pt3d, float* pt3d_rot, float* pt2d, int nopoints, int sc) const {
  project(pt3d, pt3d_rot, pt2d, nopoints, sc);
}

}  // namespace CTR
