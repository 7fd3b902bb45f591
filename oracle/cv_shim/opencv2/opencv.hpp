// Minimal stand-in for <opencv2/opencv.hpp> -- TEST INFRASTRUCTURE ONLY.
//
// OpenCV's C++ headers/libraries are not installed in this image.  This shim provides exactly the
// slice of the cv:: API that the reference's src/orb_cpu.cpp uses, so that file (and
// src/orb_pattern.cpp) can be compiled UNMODIFIED from /root/reference into oracle/_ref/ and
// serve as a witness for the oracle's single-level mode.  It is written from the OpenCV API
// documentation, not from OpenCV sources.  Arithmetic primitives (resize / GaussianBlur /
// integral) forward to the oracle's restatements, which are themselves checked against Python
// cv2 4.13 in tests/test_oracle_cv2.py.
//
// One deliberate extension: every Mat allocation carries 4 zeroed slack rows after its last row.
// The reference's BRIEF bound check (src/orb_cpu.cpp:240-245) lets sum5x5 (:190-201) index up to
// two columns / rows past the integral image; with the slack those reads are *defined*
// (decision D7 of SURVEY.md 8(c)): column overruns wrap into the next row exactly as the flat
// address arithmetic of cv::Mat::at does, row overruns read zeros.
#ifndef ORB_ORACLE_CV_SHIM_HPP
#define ORB_ORACLE_CV_SHIM_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../orb_oracle.h"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_8UC1 CV_8U
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

namespace cv {

struct Point {
  int x, y;
  Point() : x(0), y(0) {}
  Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
enum { INTER_LINEAR = 1 };

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  // non-owning view of caller memory (cv::Mat(rows, cols, type, ptr, step))
  Mat(int r, int c, int type, void* ptr, size_t step_) : rows(r), cols(c), step(step_), data((uchar*)ptr), type_(type) {}

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * elem(type);
    buf_ = std::make_shared<std::vector<uchar>>(step * (size_t)(r + 4), (uchar)0);   // 4 zero slack rows
    data = buf_->data();
  }
  static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
  int type() const { return type_; }
  int channels() const { return 1; }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == nullptr; }

  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }

  Mat operator()(const Rect& r) const {   // ROI view sharing the buffer
    Mat m;
    m.rows = r.height; m.cols = r.width; m.step = step; m.type_ = type_; m.buf_ = buf_;
    m.data = data + (size_t)r.y * step + (size_t)r.x * elem(type_);
    return m;
  }

 private:
  static size_t elem(int type) { return type == CV_8U ? 1 : 4; }
  int type_ = CV_8U;
  std::shared_ptr<std::vector<uchar>> buf_;
};

inline void minMaxLoc(const Mat& m, double* minVal, double* maxVal) {
  CV_Assert(m.type() == CV_32F);
  double lo = 1e300, hi = -1e300;
  for (int y = 0; y < m.rows; y++)
    for (int x = 0; x < m.cols; x++) {
      double v = m.at<float>(y, x);
      if (v < lo) lo = v;
      if (v > hi) hi = v;
    }
  if (minVal) *minVal = lo;
  if (maxVal) *maxVal = hi;
}

inline void integral(const Mat& src, Mat& sum) {
  CV_Assert(src.type() == CV_8U);
  std::vector<int32_t> flat((size_t)(src.cols + 1) * (src.rows + 4));
  orc_integral_flat(src.data, src.cols, src.rows, src.step, flat.data());
  sum.create(src.rows + 1, src.cols + 1, CV_32S);   // create() zeroes rows+4 rows
  std::memcpy(sum.data, flat.data(), flat.size() * sizeof(int32_t));
}

inline void resize(const Mat& src, Mat& dst, Size dsize, double, double, int) {
  Mat out(dsize.height, dsize.width, CV_8U);
  orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, out.data, out.cols, out.rows, out.step);
  dst = out;
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double) {
  CV_Assert(ksize.width == 5 && ksize.height == 5);
  Mat out(src.rows, src.cols, CV_8U);
  orc_gauss5x5_u8(src.data, src.cols, src.rows, src.step, out.data, out.step);
  dst = out;
}

inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

}  // namespace cv
#endif
