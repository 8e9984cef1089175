// oracle/ref_capi.cpp — C wrapper around the UNMODIFIED reference classes (test infrastructure, NOT the product).
//
// Compiled by oracle/Makefile together with /root/reference/{utilities,camera,pose,odometer}.cpp (where they lie;
// nothing is copied) against the stand-in headers in oracle/shim/ into oracle/_ref/libictrack_ref.so.  Used to
// pin oracle/ictrack_oracle.c (tests/test_oracle_vs_ref.py) and to generate tests/golden/*.npz.
// Exists only in the build container: /root/reference is absent on the GPU box, the prebuilt .so travels.
#include "utilities.h"
#include "camera.h"
#include "pose.h"
#include "odometer.h"

#include <cstddef>
#include <cstdint>
#include "../include/ictrack.h"

using namespace CTR;

static_assert(sizeof(ict_optparam) == sizeof(optparam), "ict_optparam must be bit-compatible with CTR::optparam");
static_assert(offsetof(ict_optparam, donorm) == offsetof(optparam, donorm), "optparam layout");
static_assert(offsetof(ict_optparam, dopatchnorm) == offsetof(optparam, dopatchnorm), "optparam layout");
static_assert(offsetof(ict_optparam, maxiter) == offsetof(optparam, maxiter), "optparam layout");
static_assert(offsetof(ict_optparam, normdp_ratio) == offsetof(optparam, normdp_ratio), "optparam layout");
static_assert(offsetof(ict_optparam, verbosity) == offsetof(optparam, verbosity), "optparam layout");

struct ict_ref_odom {
  CamClass* cam;
  PoseClass* pose;
  OdometerClass* odom;
};

extern "C" {

ict_ref_odom* ict_ref_odom_create(const ict_optparam* op, const float* fc, const float* cc, const int* wh) {
  const optparam* rop = reinterpret_cast<const optparam*>(op);
  ict_ref_odom* o = new ict_ref_odom;
  o->cam = new CamClass(rop->lv_f + 1, fc, cc, wh, rop->psz);   // run_io_reprojection_test.cpp:189
  o->pose = new PoseClass(o->cam, rop);
  o->odom = new OdometerClass(o->pose, rop);
  return o;
}
void ict_ref_odom_destroy(ict_ref_odom* o) {
  if (!o) return;
  delete o->odom; delete o->pose; delete o->cam; delete o;
}
void ict_ref_set3dpoints(ict_ref_odom* o, double* pt_in, int n) { o->odom->Set3Dpoints(pt_in, n); }
void ict_ref_setpose(ict_ref_odom* o, const double* p_in, const float** img_ref, const float** img_ref_dx,
                     const float** img_ref_dy, const float** img_new) {
  o->odom->SetPose(p_in, img_ref, img_ref_dx, img_ref_dy, img_new);
}
void ict_ref_trackpose(ict_ref_odom* o, double* p_out) { o->odom->TrackPose(p_out); }
const float* ict_ref_get2dpoints(ict_ref_odom* o) { return o->odom->Get2DPoints(); }

// (H, J^T r, delta_p) of every SolveLinSystem() call, recorded by the shim's fullPivLu().solve()
void ict_ref_set_solve_trace(float* buf, int cap) {
  Eigen::shim::SolveTrace& t = Eigen::shim::solve_trace();
  t.buf = buf; t.cap = cap; t.n = 0;
}
int ict_ref_solve_trace_count(void) { return Eigen::shim::solve_trace().n; }

void ict_ref_camera_levels(int noscales, const float* fc, const float* cc, const int* wh, int padding, float* out) {
  CamClass cam(noscales, fc, cc, wh, padding);
  for (int l = 0; l < noscales; ++l) {
    float* o = out + 8 * l;
    o[0] = cam.getfx(l); o[1] = cam.getfy(l); o[2] = cam.getcx(l); o[3] = cam.getcy(l);
    o[4] = cam.getswo(l); o[5] = cam.getsho(l); o[6] = cam.getsw(l); o[7] = cam.getsh(l);
  }
}

void ict_ref_pyramid_build(const float* img, int w, int h, int lv_f, int pad, float* out_I, float* out_dx, float* out_dy) {
  cv::Mat src(h, w, CV_32F);
  std::memcpy(src.data, img, sizeof(float) * (size_t)w * h);
  std::vector<cv::Mat> pi(lv_f + 1), px(lv_f + 1), py(lv_f + 1);
  std::vector<const float*> qi(lv_f + 1), qx(lv_f + 1), qy(lv_f + 1);
  util_constructpyramide(src, pi.data(), px.data(), py.data(), qi.data(), qx.data(), qy.data(), lv_f, 1, pad);
  size_t off = 0;
  for (int l = 0; l <= lv_f; ++l) {
    const size_t n = (size_t)pi[l].rows * pi[l].cols;
    std::memcpy(out_I + off, qi[l], sizeof(float) * n);
    std::memcpy(out_dx + off, qx[l], sizeof(float) * n);
    std::memcpy(out_dy + off, qy[l], sizeof(float) * n);
    off += n;
  }
}

void ict_ref_se3_exp_f(float* G, const float* p) { util_SE3_coeff_to_group<float>(G, p); }
void ict_ref_se3_exp_d(double* G, const double* p) { util_SE3_coeff_to_group<double>(G, p); }
void ict_ref_se3_log_f(float* p, const float* G) { util_SE3_group_to_coeff<float>(p, G); }
void ict_ref_se3_log_d(double* p, const double* G) { util_SE3_group_to_coeff<double>(p, G); }

void ict_ref_getpatch(const float* img, const float* mid, float* out, const ict_optparam* op, int width) {
  const optparam* rop = reinterpret_cast<const optparam*>(op);
  Eigen::Map<MatrixXfTr, Eigen::Aligned> m(out, rop->psz, rop->psz);
  util_getPatch(img, mid, &m, rop, width);
}
void ict_ref_getpatch_grad(const float* img, const float* dx, const float* dy, const float* mid, float* out,
                           float* out_dx, float* out_dy, const ict_optparam* op, int width) {
  const optparam* rop = reinterpret_cast<const optparam*>(op);
  Eigen::Map<MatrixXfTr, Eigen::Aligned> m(out, rop->psz, rop->psz), mx(out_dx, rop->psz, rop->psz),
      my(out_dy, rop->psz, rop->psz);
  util_getPatch_grad(img, dx, dy, mid, &m, &mx, &my, rop, width);
}

// Batch over independent tracks == the sid loop of run_track_nposes.cpp:193, one reference odometer per thread.
int ict_ref_track_batch(const ict_optparam* op, const float* fc, const float* cc, const int* wh,
                        const float* const* planes_I, const float* const* planes_dx, const float* const* planes_dy,
                        int T, const int64_t* pt_off, const double* pts, const int* ref_frame, const int* new_frame,
                        const double* p_in, double* p_out, int nthreads) {
  int64_t off[ICT_MAX_LEVELS];
  int64_t tot = 0;
  for (int l = 0; l <= op->lv_f; ++l) {
    off[l] = tot;
    tot += (int64_t)((wh[0] >> l) + 2 * op->psz) * ((wh[1] >> l) + 2 * op->psz);
  }
  if (nthreads <= 0) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    ict_ref_odom* o = ict_ref_odom_create(op, fc, cc, wh);
    std::vector<double> buf;
    const float *ri[ICT_MAX_LEVELS], *rx[ICT_MAX_LEVELS], *ry[ICT_MAX_LEVELS], *ni[ICT_MAX_LEVELS];
#pragma omp for schedule(dynamic, 4)
    for (int t = 0; t < T; ++t) {
      const int n = (int)(pt_off[t + 1] - pt_off[t]);
      buf.assign(pts + 3 * pt_off[t], pts + 3 * pt_off[t] + 3 * (size_t)n);
      for (int l = 0; l <= op->lv_f; ++l) {
        ri[l] = planes_I[ref_frame[t]] + off[l];
        rx[l] = planes_dx[ref_frame[t]] + off[l];
        ry[l] = planes_dy[ref_frame[t]] + off[l];
        ni[l] = planes_I[new_frame[t]] + off[l];
      }
      o->odom->Set3Dpoints(buf.data(), n);
      o->odom->SetPose(p_in + 6 * (size_t)t, ri, rx, ry, ni);
      o->odom->TrackPose(p_out + 6 * (size_t)t);
    }
    ict_ref_odom_destroy(o);
  }
  return 0;
}

}  // extern "C"
