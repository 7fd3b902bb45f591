// orb_cv_min.hpp -- the few cv:: names include/orb.hpp needs, for builds WITHOUT OpenCV.
// With OpenCV installed include <opencv2/core.hpp> before orb.hpp and this file is not used.
// A cv::Mat here is only a non-owning view (rows, cols, step in bytes, data, type) of single-channel elements
// (8-bit pixels; 32-bit float score maps for NMS(); 32-bit integral images for RotatedBRIEFCPU::sum5x5).
#ifndef ORB_CV_MIN_HPP
#define ORB_CV_MIN_HPP
#include <cstddef>
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
  int type() const { return type_; }
  int channels() const { return 1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
 private:
  int type_ = CV_8UC1;
};
}  // namespace cv
#endif
#endif
