// utilities.h — host-side mirror of the reference's utilities.h for the tracking path (namespace CTR, same names,
// same argument meaning).  The pixel work behind these calls runs on the GPU through libictrack.so.
//   optparam                  == CTR::optparam, utilities.h:46-61 (typedef of the bit-compatible ict_optparam)
//   util_constructpyramide    == utilities.h:63 / utilities.cpp:14-52: builds the pyramid ON THE DEVICE, keeps it there
//                                and also fills the caller's cv::Mat arrays + pointer tables with host copies
//   util_SE3_coeff_to_group / util_SE3_group_to_coeff == utilities.h:84-241 (host scalar helpers of PoseClass)
//   util_getPatch / util_getPatch_grad == utilities.h:74-79 / utilities.cpp:55-189: the plane pointers must be level planes
//                                made by util_constructpyramide (their device twins are looked up by pointer); one
//                                GPU round trip per call — the tracker itself never needs them (its patches stay on
//                                the device; the NCC scoring of run_track_nposes.cpp:271-355 is ict_ncc_score())
#ifndef ICT_HOST_UTIL_HEADER
#define ICT_HOST_UTIL_HEADER

#include <cmath>
#include <cstring>

#include "ict_compat.h"
#include "ictrack.h"

#define SSEMULTIPL 4
#define LIEALG_SIGTHRESH 1e-4
#define LIEALG_EPSILON 1e-10

namespace CTR {

typedef ict_optparam optparam;

void util_constructpyramide(const cv::Mat& img_ao_fmat, cv::Mat* img_ao_fmat_pyr, cv::Mat* img_ao_dx_fmat_pyr,
                            cv::Mat* img_ao_dy_fmat_pyr, const float** img_ao_pyr, const float** img_ao_dx_pyr,
                            const float** img_ao_dy_pyr, const int lv_f, const bool getgrad, const int imgpadding);

typedef Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> MatrixXfTr;   // utilities.h:33

void util_getPatch(const float* img, const float* mid_in, Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_in_e,
                   const optparam* op, const int width);
void util_getPatch_grad(const float* img, const float* img_dx, const float* img_dy, const float* mid_in,
                        Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_in_e, Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_dx_in_e,
                        Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_dy_in_e, const optparam* op, const int width);

// Device twin of a host pyramid made by util_constructpyramide, keyed by the level-0 intensity plane pointer.
// Returns false for planes this library did not build (OdometerClass then uploads them).
bool util_find_device_pyramid(const float* level0_plane, ict_frames** store, int* frame);
// Drops the device twins (and the registry) — call when the cv::Mat arrays go away.
void util_release_device_pyramids();

// exp: se(3) coefficients p = [u, w] -> 3x4 row-major [R | V u] (Eade, "Lie Groups for Computer Vision")
template <typename T>
void util_SE3_coeff_to_group(T* G, const T* p) {
  const T w[3] = {p[3], p[4], p[5]};
  const T ww[3] = {w[0] * w[0], w[1] * w[1], w[2] * w[2]};
  const T s = (T)std::sqrt((double)(T)(ww[0] + ww[1] + ww[2]));
  const T s2 = s * s, s3 = s * s * s;
  T a, b, c;
  if (s > LIEALG_SIGTHRESH) {
    const double sn = std::sin((double)s), cs = std::cos((double)s);
    a = (T)(sn / (double)s);
    b = (T)((1 - cs) / (double)s2);
    c = (T)(((double)s - sn) / (double)s3);
  } else {
    a = 1 - s2 / 6 * (1 - s2 / 20 * (1 - s2 / 42));
    b = (T)(.5 * (double)(T)(1 - s2 / 12 * (1 - s2 / 30 * (1 - s2 / 56))));
    c = (1 - s2 / 20 * (1 - s2 / 42 * (1 - s2 / 72))) / 6;
  }
  const T xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  // R = I + a [w]x + b [w]x^2
  G[0] = 1 - ww[1] * b - ww[2] * b;  G[1] = xy * b - w[2] * a;         G[2] = w[1] * a + xz * b;
  G[4] = w[2] * a + xy * b;          G[5] = 1 - ww[0] * b - ww[2] * b; G[6] = yz * b - w[0] * a;
  G[8] = xz * b - w[1] * a;          G[9] = w[0] * a + yz * b;         G[10] = 1 - ww[0] * b - ww[1] * b;
  // t = V u, V = I + b [w]x + c [w]x^2
  G[3] = (1 - (ww[1] + ww[2]) * c) * p[0] + (xy * c - w[2] * b) * p[1] + (w[1] * b + xz * c) * p[2];
  G[7] = (w[2] * b + xy * c) * p[0] + (1 - (ww[0] + ww[2]) * c) * p[1] + (yz * c - w[0] * b) * p[2];
  G[11] = (xz * c - w[1] * b) * p[0] + (w[0] * b + yz * c) * p[1] + (1 - (ww[0] + ww[1]) * c) * p[2];
}

// log: 3x4 row-major -> coefficients
template <typename T>
void util_SE3_group_to_coeff(T* p, const T* G) {
  const T tr = G[0] + G[5] + G[10];
  const T th = (T)std::acos((double)(T)(0.5f * (tr - 1)));
  T W[3] = {0, 0, 0};   // (w_x, w_y, w_z)
  if (!(th < LIEALG_EPSILON)) {
    const T k = (T)((double)th / ((double)2.0f * std::sin((double)th)));
    W[0] = -(k * (G[6] - G[9]));
    W[1] = k * (G[2] - G[8]);
    W[2] = -(k * (G[1] - G[4]));
  }
  p[3] = W[0]; p[4] = W[1]; p[5] = W[2];
  // [w]x and its square
  const T K[9] = {0, -W[2], W[1], W[2], 0, -W[0], -W[1], W[0], 0};
  T K2[9];
  K2[0] = -(W[2] * W[2]) - W[1] * W[1]; K2[1] = -W[1] * -W[0]; K2[2] = -W[2] * -W[0];
  K2[3] = K2[1]; K2[4] = -(W[2] * W[2]) - W[0] * W[0]; K2[5] = -(-W[2] * W[1]);
  K2[6] = K2[2]; K2[7] = K2[5]; K2[8] = -(W[1] * W[1]) - W[0] * W[0];
  T h;
  if (th < LIEALG_SIGTHRESH)
    h = 1.0f / 12.0f;
  else
    h = (T)(((double)1.0f - (double)th / ((double)2.0f * std::tan((double)(T)(th / 2.0f)))) / (double)(T)(th * th));
  T Vi[9];
  for (int k = 0; k < 9; ++k) Vi[k] = ((k % 4 == 0) ? (T)1.0f : -0.5f * K[k]) + h * K2[k];
  for (int r = 0; r < 3; ++r) p[r] = Vi[3 * r] * G[3] + Vi[3 * r + 1] * G[7] + Vi[3 * r + 2] * G[11];
}

}  // namespace CTR
#endif
