// Mock of cv::Feature2D (OpenCV 4 <opencv2/features2d.hpp>) -- TEST SCAFFOLDING ONLY; same virtual signatures.
#ifndef MOCK_OPENCV_FEATURES2D_HPP
#define MOCK_OPENCV_FEATURES2D_HPP
#include "core.hpp"
namespace cv {
class Feature2D {
 public:
  virtual ~Feature2D() {}
  virtual void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray()) {
    detectAndCompute(image, mask, keypoints, noArray(), false);
  }
  virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) {
    detectAndCompute(image, noArray(), keypoints, descriptors, true);
  }
  virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false) {}
  virtual int descriptorSize() const { return 0; }
  virtual int descriptorType() const { return CV_32F; }
  virtual int defaultNorm() const { return NORM_L2; }
  virtual String getDefaultName() const { return "Feature2D"; }
};
}  // namespace cv
#endif
