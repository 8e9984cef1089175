// oracle/shim/opencv2/core/core.hpp — STAND-IN for the subset of OpenCV the reference's tracking path uses
// (test infrastructure, NOT the product, NOT OpenCV).
//
// OpenCV is an unpinned external dependency of the reference (CMakeLists.txt:10) whose C++ headers/libs are
// not installed in this image.  This header lets the UNMODIFIED reference sources compile (oracle/Makefile).
// The three image operations of util_constructpyramide (utilities.cpp:24-46) are restated from OpenCV's
// documented behaviour and PINNED bit-exact against Python cv2 4.13 on uint8-valued images by
// tests/test_oracle_golden.py (fixtures tests/golden/pyramid_*.npz made by tests/golden/make_golden.py):
//   resize(.5,.5,INTER_LINEAR)  exact factor 2 -> OpenCV's area-fast path, ((a+b)+(c+d))*0.25
//   Sobel(ksize=1)              [-1 0 1], BORDER_REFLECT_101
//   copyMakeBorder              BORDER_REPLICATE / BORDER_CONSTANT
// imread reads binary PGM (P5) only.
#ifndef ICT_SHIM_OPENCV_CORE
#define ICT_SHIM_OPENCV_CORE

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 CV_8U
#define CV_32FC1 CV_32F
#define CV_LOAD_IMAGE_GRAYSCALE 0

namespace cv {

typedef unsigned char uchar;
enum { INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { IMREAD_GRAYSCALE = 0 };

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

class Mat {
 public:
  Mat() : rows(0), cols(0), data(0), type_(CV_8U) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    st_.reset(new std::vector<uchar>((size_t)r * c * elem()));
    data = st_->empty() ? 0 : st_->data();
  }
  int type() const { return type_; }
  size_t elem() const { return type_ == CV_32F ? 4 : 1; }
  Mat clone() const { Mat m(rows, cols, type_); if (data) std::memcpy(m.data, data, (size_t)rows * cols * elem()); return m; }
  void convertTo(Mat& dst, int type) const {
    Mat src = *this;                                   // keeps the storage alive if dst aliases *this
    Mat out(rows, cols, type);
    const size_t n = (size_t)rows * cols;
    if (src.type_ == type) { if (n) std::memcpy(out.data, src.data, n * elem()); }
    else if (src.type_ == CV_8U && type == CV_32F) { for (size_t i = 0; i < n; ++i) ((float*)out.data)[i] = (float)src.data[i]; }
    else { std::fprintf(stderr, "shim cv::Mat::convertTo: unsupported conversion\n"); }
    dst = out;
  }
  float* f() const { return (float*)data; }
  int rows, cols;
  uchar* data;
 private:
  int type_;
  std::shared_ptr<std::vector<uchar> > st_;
};

inline void resize(const Mat& src_in, Mat& dst, Size, double fx, double fy, int /*interpolation*/) {
  Mat src = src_in;
  if (fx != .5 || fy != .5 || src.type() != CV_32F || (src.cols & 1) || (src.rows & 1))
    std::fprintf(stderr, "shim cv::resize: only exact 1/2 downscale of even-sized CV_32F images\n");
  Mat out(src.rows / 2, src.cols / 2, CV_32F);
  for (int y = 0; y < out.rows; ++y)
    for (int x = 0; x < out.cols; ++x) {
      const float* r0 = src.f() + (size_t)(2 * y) * src.cols + 2 * x;
      const float* r1 = r0 + src.cols;
      out.f()[(size_t)y * out.cols + x] = ((r0[0] + r0[1]) + (r1[0] + r1[1])) * 0.25f;
    }
  dst = out;
}

inline void Sobel(const Mat& src_in, Mat& dst, int /*ddepth*/, int dx, int dy, int /*ksize=1*/, double, double, int) {
  Mat src = src_in;
  Mat out(src.rows, src.cols, CV_32F);
  const int w = src.cols, h = src.rows;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float g;
      if (dx == 1) {
        int xp = x + 1 < w ? x + 1 : (w >= 2 ? w - 2 : 0), xm = x - 1 >= 0 ? x - 1 : (w >= 2 ? 1 : 0);   // reflect-101
        g = src.f()[(size_t)y * w + xp] - src.f()[(size_t)y * w + xm];
      } else {
        int yp = y + 1 < h ? y + 1 : (h >= 2 ? h - 2 : 0), ym = y - 1 >= 0 ? y - 1 : (h >= 2 ? 1 : 0);
        g = src.f()[(size_t)yp * w + x] - src.f()[(size_t)ym * w + x];
      }
      out.f()[(size_t)y * w + x] = g;
    }
  (void)dy;
  dst = out;
}

inline void copyMakeBorder(const Mat& src_in, Mat& dst, int top, int bottom, int left, int right, int type, double value = 0) {
  Mat src = src_in;
  Mat out(src.rows + top + bottom, src.cols + left + right, CV_32F);
  for (int Y = 0; Y < out.rows; ++Y)
    for (int X = 0; X < out.cols; ++X) {
      int y = Y - top, x = X - left;
      float v;
      if (type == BORDER_REPLICATE) {
        int yc = y < 0 ? 0 : (y >= src.rows ? src.rows - 1 : y), xc = x < 0 ? 0 : (x >= src.cols ? src.cols - 1 : x);
        v = src.f()[(size_t)yc * src.cols + xc];
      } else {
        v = (y >= 0 && y < src.rows && x >= 0 && x < src.cols) ? src.f()[(size_t)y * src.cols + x] : (float)value;
      }
      out.f()[(size_t)Y * out.cols + X] = v;
    }
  dst = out;
}

inline Mat imread(const std::string& name, int /*flags*/) {
  Mat m;
  FILE* f = std::fopen(name.c_str(), "rb");
  if (!f) return m;
  char magic[3] = {0, 0, 0};
  int w = 0, h = 0, maxv = 0;
  if (std::fscanf(f, "%2s", magic) == 1 && std::strcmp(magic, "P5") == 0) {
    int c = std::fgetc(f);
    while (c == '#' || c == ' ' || c == '\n' || c == '\r' || c == '\t') {
      if (c == '#') while (c != '\n' && c != EOF) c = std::fgetc(f);
      c = std::fgetc(f);
    }
    std::ungetc(c, f);
    if (std::fscanf(f, "%d %d %d", &w, &h, &maxv) == 3 && maxv == 255) {
      std::fgetc(f);
      m.create(h, w, CV_8U);
      if (std::fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) m = Mat();
    }
  }
  std::fclose(f);
  return m;
}

}  // namespace cv
#endif
