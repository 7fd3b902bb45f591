// orb_cv_min.hpp -- the few cv:: names include/orb.hpp needs, for builds WITHOUT OpenCV.
// With OpenCV installed include <opencv2/core.hpp> before orb.hpp and this file is not used.
// A cv::Mat here is a non-owning view (or, after create(), a tightly packed owned buffer) (rows, cols, step in bytes, data, type) of single-channel elements
// (8-bit pixels; 32-bit float score maps for NMS(); 32-bit integral images for RotatedBRIEFCPU::sum5x5).
#ifndef ORB_CV_MIN_HPP
#define ORB_CV_MIN_HPP
#include <cstddef>
#include <memory>
#ifndef CV_8UC1
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32SC1 4
#define CV_32F 5
#define CV_32FC1 5
namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type, void* ptr, size_t step_ = 0)
      : rows(r), cols(c), step(step_ ? step_ : (size_t)c * (type == CV_8UC1 ? 1 : 4)), data((unsigned char*)ptr), type_(type) {}
  // (re)allocate an owned, tightly packed matrix (cv::Mat::create): what the filter wrappers do with their dst argument
  void create(int r, int c, int type) {
    const size_t es = type == CV_8UC1 ? 1 : 4;
    if (r == rows && c == cols && type == type_ && data) return;
    owned_.reset(new unsigned char[(size_t)r * c * es]);
    rows = r; cols = c; step = (size_t)c * es; data = owned_.get(); type_ = type;
  }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  int type() const { return type_; }
  int channels() const { return 1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
 private:
  int type_ = CV_8UC1;
  std::shared_ptr<unsigned char[]> owned_;
};
}  // namespace cv
#endif
#endif
