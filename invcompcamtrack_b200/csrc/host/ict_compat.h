// ict_compat.h — the two third-party types that appear in the reference's signatures (Eigen::Vector3d in
// pose.h:23, cv::Mat in utilities.h:63).  When the real libraries are installed their headers are used; this image
// has neither, so the smallest same-named stand-ins are defined instead (host bookkeeping only, no arithmetic).
#pragma once
#include <memory>
#include <vector>

#if defined(__has_include) && __has_include(<Eigen/Core>) && !defined(ICT_NO_EIGEN)
#include <Eigen/Core>
#else
namespace Eigen {
enum { Aligned = 1, Dynamic = -1, RowMajor = 1 };
// the reference's patch buffers are Eigen::Map<MatrixXfTr, Eigen::Aligned> over caller-owned floats (odometer.h:70-99):
// only data() is needed on this side
template <typename M, int A = 0>
struct Map {
  float* p;
  int n;
  Map(float* d, int rows, int cols = 1) : p(d), n(rows * cols) {}
  float* data() { return p; }
  const float* data() const { return p; }
  int size() const { return n; }
};
template <typename S, int R, int C, int O = 0>
struct Matrix {};
struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  double* data() { return v; }
  const double* data() const { return v; }
};
}  // namespace Eigen
#endif

#if defined(__has_include) && __has_include(<opencv2/core/core.hpp>) && !defined(ICT_NO_OPENCV)
#include <opencv2/core/core.hpp>
#else
#define ICT_COMPAT_MAT 1
#ifndef CV_8U
#define CV_8U 0
#define CV_32F 5
#endif
namespace cv {
// rows x cols single-channel matrix, CV_8U or CV_32F, reference-counted storage
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), type_(CV_8U) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    st_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * (type == CV_32F ? 4 : 1));
    data = st_->data();
  }
  int type() const { return type_; }
  bool empty() const { return data == nullptr; }
  int rows, cols;
  unsigned char* data;
 private:
  int type_;
  std::shared_ptr<std::vector<unsigned char>> st_;
};
}  // namespace cv
#endif
