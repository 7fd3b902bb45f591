// Mock of the slice of <opencv2/core.hpp> that include/orb_feature2d.hpp touches -- TEST SCAFFOLDING ONLY.
// OpenCV C++ is not installed in this image; this header follows the documented OpenCV 4 signatures
// (cv::Mat, cv::InputArray / cv::OutputArray, cv::KeyPoint, cv::Ptr, cv::noArray) closely enough that the adapter is
// compiled and run against the same calls a real OpenCV build would resolve.
#ifndef MOCK_OPENCV_CORE_HPP
#define MOCK_OPENCV_CORE_HPP
#include <cstddef>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {
typedef std::string String;
typedef unsigned char uchar;
enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct TermCriteria {
  enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
  int type = 0, maxCount = 0; double epsilon = 0;
  TermCriteria() {}
  TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};
enum ImreadModes { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };

class KeyPoint {
 public:
  KeyPoint() {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ptr, size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : (size_t)c), data((uchar*)ptr), type_(type) {}
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = (size_t)c * (type == CV_8U ? 1 : 4);
    buf_ = std::make_shared<std::vector<uchar>>(step * (size_t)(r > 0 ? r : 0));
    data = buf_->data();
  }
  int type() const { return type_; }
  int channels() const { return 1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * (type_ == CV_8U ? 1 : 4); }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
 private:
  int type_ = CV_8U;
  std::shared_ptr<std::vector<uchar>> buf_;
};

class _InputArray {
 public:
  _InputArray() {}
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  bool empty() const { return !m_ || m_->empty(); }
 protected:
  Mat* m_ = nullptr;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) { m_ = &m; }
  bool needed() const { return m_ != nullptr; }
  void create(int rows, int cols, int type) const { if (m_) m_->create(rows, cols, type); }
  void release() const { if (m_) m_->release(); }
  Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline const _OutputArray& noArray() { static _OutputArray none; return none; }
}  // namespace cv
#endif
