// orb_feature2d.hpp -- cv::Feature2D adapter over the B200 ORB extractor.
//
// The reference's VO programs hold a cv::Ptr<cv::ORB> / cv::Ptr<cv::SIFT> and call
//   detector->detectAndCompute(img, cv::noArray(), std::vector<cv::KeyPoint>&, cv::Mat& descriptors)
//   detector->detect(img, keypoints)
// (src/feature_matching.cpp:56,164; src/feature_tracking.cpp:59,201-202) and feed the descriptors to a cv::DescriptorMatcher.
// This class lets them hold `cv::Ptr<cv::Feature2D> orb = orb_b200::ORBFeature2D::create(3000);` instead.
// It needs OpenCV's C++ headers (<opencv2/features2d.hpp>); the image this repository is built in has none, so the unit test
// compiles it against tests/cpp/mock_opencv (same virtual signatures).  Semantics: the extractor of include/orb_b200.h
// (pyramid, FAST-9 + NMS, Harris top-N per level, orientation, rotated BRIEF), converted to OpenCV's conventions:
//   KeyPoint.pt = level-0 coordinates (int(x * scale), reference src/orb.cpp:94-98), .size = patch * scale^level,
//   .angle = orientation in degrees in [0, 360), .response = Harris response, .octave = pyramid level;
//   descriptors = CV_8U, N x 32 (what FlannBasedMatcher with LshIndexParams / BFMatcher(NORM_HAMMING) expect).
// Masks and useProvidedKeypoints are not supported (the reference never passes them: it always hands cv::noArray()).
#ifndef ORB_FEATURE2D_HPP
#define ORB_FEATURE2D_HPP

#include <opencv2/features2d.hpp>

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "orb_b200.h"

namespace orb_b200 {

class ORBFeature2D : public cv::Feature2D {
 public:
  // same leading arguments as cv::ORB::create(nfeatures, scaleFactor, nlevels, ...): drop-in for `cv::ORB::create(3000)`
  static cv::Ptr<ORBFeature2D> create(int nfeatures = 500, float scaleFactor = 1.2f, int nlevels = 8, int fastThreshold = 20,
                                      int patchSize = 31) {
    return cv::makePtr<ORBFeature2D>(nfeatures, scaleFactor, nlevels, fastThreshold, patchSize);
  }

  ORBFeature2D(int nfeatures, float scaleFactor, int nlevels, int fastThreshold, int patchSize) {
    orb_default_params(&p_);
    p_.nfeatures = nfeatures; p_.scale_factor = scaleFactor; p_.nlevels = nlevels;
    p_.fast_threshold = fastThreshold; p_.orient_patch = patchSize;
    p_.keep_side_arrays = 1;                       // level id and response become KeyPoint.octave / .response
    p_.max_width = 0; p_.max_height = 0;           // context is created on the first image
  }
  ORBFeature2D(const ORBFeature2D&) = delete;
  ORBFeature2D& operator=(const ORBFeature2D&) = delete;
  ~ORBFeature2D() override { orb_destroy(ctx_); }

  void detectAndCompute(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                        cv::OutputArray descriptors, bool useProvidedKeypoints = false) override {
    if (useProvidedKeypoints) throw std::runtime_error("orb_b200::ORBFeature2D: useProvidedKeypoints is not supported");
    if (!mask.empty()) throw std::runtime_error("orb_b200::ORBFeature2D: masks are not supported");
    const cv::Mat img = image.getMat();
    if (img.empty() || img.type() != CV_8UC1) throw std::runtime_error("orb_b200::ORBFeature2D: image must be CV_8UC1");
    ensure(img.cols, img.rows);
    int cap = 0;
    for (int l = 0; l < p_.nlevels; l++) cap += orb_level_quota(ctx_, l);
    if (cap < 1) cap = 1;
    std::vector<orb_keypoint> k(cap), lxy(cap);
    std::vector<float> ang(cap), resp(cap);
    std::vector<int32_t> lvl(cap);
    std::vector<orb_descriptor> d(cap);
    int n = 0;
    check(orb_detect_and_compute(ctx_, img.data, img.cols, img.rows, img.step, cap, k.data(), ang.data(), d.data(), &n, nullptr));
    check(orb_get_side_arrays(ctx_, 0, n, lxy.data(), lvl.data(), resp.data()));
    keypoints.clear();
    keypoints.reserve(n);
    for (int i = 0; i < n; i++) {
      float deg = ang[i] * (180.0f / 3.14159265358979323846f);
      if (deg < 0.f) deg += 360.f;
      if (deg >= 360.f) deg -= 360.f;
      const float size = (float)p_.orient_patch * std::pow(p_.scale_factor, (float)lvl[i]);
      keypoints.push_back(cv::KeyPoint((float)k[i].x, (float)k[i].y, size, deg, resp[i], lvl[i]));
    }
    if (descriptors.needed()) {
      descriptors.create(n, 32, CV_8U);
      cv::Mat out = descriptors.getMatRef();
      for (int i = 0; i < n; i++) std::memcpy(out.ptr<unsigned char>(i), d[i].data, 32);
    }
  }

  void detect(cv::InputArray image, std::vector<cv::KeyPoint>& keypoints, cv::InputArray mask = cv::noArray()) override {
    detectAndCompute(image, mask, keypoints, cv::noArray(), false);
  }

  int descriptorSize() const override { return 32; }
  int descriptorType() const override { return CV_8U; }
  int defaultNorm() const override { return cv::NORM_HAMMING; }
  cv::String getDefaultName() const override { return "Feature2D.ORB_B200"; }

 private:
  void ensure(int w, int h) {
    if (ctx_ && w <= p_.max_width && h <= p_.max_height) return;
    orb_destroy(ctx_);
    ctx_ = nullptr;
    if (w > p_.max_width) p_.max_width = w;
    if (h > p_.max_height) p_.max_height = h;
    if (orb_create(&p_, &ctx_) != ORB_OK) throw std::runtime_error(std::string("orb_b200: orb_create failed: ") + orb_last_error(nullptr));
  }
  void check(int rc) const {
    if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error(ctx_));
  }
  orb_params p_;
  orb_ctx* ctx_ = nullptr;
};

}  // namespace orb_b200
#endif
