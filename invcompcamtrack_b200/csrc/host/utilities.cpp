// utilities.cpp — host side of util_constructpyramide: the pyramid is built by the CUDA kernels of libictrack.so and
// stays resident on the device; the caller's cv::Mat arrays receive host copies so that code written against the
// reference (which reads the planes through raw float pointers) keeps working.
#include "utilities.h"

#include <cstdio>
#include <map>
#include <vector>

namespace CTR {

namespace {
struct DevicePyramid { ict_frames* store; int frame; };
std::map<const float*, DevicePyramid>& registry() {
  static std::map<const float*, DevicePyramid> r;
  return r;
}
// every intensity level plane of a host pyramid -> (device store, frame, level), for util_getPatch*
struct DevicePlane { ict_frames* store; int frame, level; };
std::map<const float*, DevicePlane>& plane_registry() {
  static std::map<const float*, DevicePlane> r;
  return r;
}
}  // namespace

void util_getPatch_grad(const float* img, const float* img_dx, const float* img_dy, const float* mid_in,
                        Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_in_e, Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_dx_in_e,
                        Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_dy_in_e, const optparam* op, const int width) {
  (void)img_dx; (void)img_dy; (void)width;     // the device twin carries all three planes and its own row pitch
  auto it = plane_registry().find(img);
  if (it == plane_registry().end()) {
    std::printf("util_getPatch: the plane was not made by util_constructpyramide\n");
    return;
  }
  if (ict_get_patches(it->second.store, it->second.frame, it->second.level, op, 1, mid_in, tmp_in_e->data(),
                      tmp_dx_in_e ? tmp_dx_in_e->data() : nullptr, tmp_dy_in_e ? tmp_dy_in_e->data() : nullptr) != ICT_OK)
    std::printf("util_getPatch: %s\n", ict_last_error());
}

void util_getPatch(const float* img, const float* mid_in, Eigen::Map<MatrixXfTr, Eigen::Aligned>* tmp_in_e,
                   const optparam* op, const int width) {
  util_getPatch_grad(img, nullptr, nullptr, mid_in, tmp_in_e, nullptr, nullptr, op, width);
}

bool util_find_device_pyramid(const float* level0_plane, ict_frames** store, int* frame) {
  auto it = registry().find(level0_plane);
  if (it == registry().end()) return false;
  *store = it->second.store;
  *frame = it->second.frame;
  return true;
}

void util_release_device_pyramids() {
  for (auto& kv : registry()) ict_frames_destroy(kv.second.store);
  registry().clear();
  plane_registry().clear();
}

void util_constructpyramide(const cv::Mat& img, cv::Mat* pyr, cv::Mat* pyr_dx, cv::Mat* pyr_dy, const float** p,
                            const float** pdx, const float** pdy, const int lv_f, const bool getgrad,
                            const int imgpadding) {
  const int w = img.cols, h = img.rows;
  int64_t off[ICT_MAX_LEVELS];
  int sw[ICT_MAX_LEVELS], sh[ICT_MAX_LEVELS];
  const int64_t total = ict_pyramid_layout(w, h, lv_f, imgpadding, off, sw, sh);
  ict_frames* fs = total > 0 ? ict_frames_create(1, w, h, lv_f, imgpadding) : nullptr;
  if (!fs) {
    std::printf("util_constructpyramide: %s\n", total > 0 ? ict_last_error() : "image size must be divisible by 2^lv_f");
    return;   // the reference has no error path either; callers see null planes
  }
  const int rc = img.type() == CV_32F ? ict_frames_upload(fs, 0, 1, (const float*)img.data)
                                      : ict_frames_upload_u8(fs, 0, 1, img.data);
  std::vector<float> I((size_t)total), dx((size_t)total), dy((size_t)total);
  if (rc != ICT_OK || ict_frames_download(fs, 0, I.data(), dx.data(), dy.data()) != ICT_OK) {
    std::printf("util_constructpyramide: %s\n", ict_last_error());
    ict_frames_destroy(fs);
    return;
  }
  for (int l = 0; l <= lv_f; ++l) {
    const size_t n = (size_t)sw[l] * sh[l];
    pyr[l].create(sh[l], sw[l], CV_32F);
    std::memcpy(pyr[l].data, I.data() + off[l], n * sizeof(float));
    p[l] = (const float*)pyr[l].data;
    plane_registry()[p[l]] = DevicePlane{fs, 0, l};
    if (getgrad) {
      pyr_dx[l].create(sh[l], sw[l], CV_32F);
      pyr_dy[l].create(sh[l], sw[l], CV_32F);
      std::memcpy(pyr_dx[l].data, dx.data() + off[l], n * sizeof(float));
      std::memcpy(pyr_dy[l].data, dy.data() + off[l], n * sizeof(float));
      pdx[l] = (const float*)pyr_dx[l].data;
      pdy[l] = (const float*)pyr_dy[l].data;
    }
  }
  registry()[p[0]] = DevicePyramid{fs, 0};
}

}  // namespace CTR
