// orb_vo_frontend.hpp -- the two OpenCV calls on either side of the detector in the reference's VO loops, over the C ABI of
// include/orb_b200.h:
//   cv::imread(path, cv::IMREAD_GRAYSCALE)                        src/feature_matching.cpp:55,59; src/feature_tracking.cpp:56,196
//   cv::calcOpticalFlowPyrLK(img_1, img_2, points1, points2, status, err, winSize, 3, termcrit, 0, 0.001)
//                                                                 src/feature_tracking.cpp:174-180
// Same argument lists, so the call sites change by a namespace: orb_b200::imread / orb_b200::calcOpticalFlowPyrLK.
// Needs OpenCV's C++ headers for cv::Mat, cv::Point2f, cv::Size, cv::TermCriteria; this repository's image has none, so the
// unit test compiles it against tests/cpp/mock_opencv.
#ifndef ORB_VO_FRONTEND_HPP
#define ORB_VO_FRONTEND_HPP

#include <opencv2/core.hpp>
#include <opencv2/imgcodecs.hpp>
#include <opencv2/video/tracking.hpp>

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "orb_b200.h"

namespace orb_b200 {

// == cv::imread(path, cv::IMREAD_GRAYSCALE): CV_8UC1, empty Mat when the file cannot be read or decoded (OpenCV's convention)
inline cv::Mat imread(const std::string& path, int flags = cv::IMREAD_GRAYSCALE) {
  if (flags != cv::IMREAD_GRAYSCALE) throw std::invalid_argument("orb_b200::imread: only IMREAD_GRAYSCALE (what the reference uses)");
  std::vector<unsigned char> head(64);
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return cv::Mat();
  const size_t got = std::fread(head.data(), 1, head.size(), f);
  std::fclose(f);
  orb_image_info info;
  if (orb_png_info(head.data(), got, &info) != ORB_OK) return cv::Mat();
  cv::Mat img(info.height, info.width, CV_8UC1);
  int w = 0, h = 0;
  if (orb_imread_gray8(path.c_str(), img.ptr<unsigned char>(0), img.step, info.width, info.height, &w, &h) != ORB_OK) return cv::Mat();
  return img;
}

namespace detail {
// one context per thread, grown on demand (the tracker needs no ORB arena to speak of: max_batch 1, one level)
struct TrackerContext {
  orb_ctx* ctx = nullptr;
  int w = 0, h = 0;
  orb_ctx* get(int width, int height) {
    if (ctx && width <= w && height <= h) return ctx;
    if (ctx) orb_destroy(ctx);
    ctx = nullptr;
    orb_params p;
    orb_default_params(&p);
    p.nfeatures = 1; p.nlevels = 1; p.max_width = width; p.max_height = height; p.max_batch = 1;
    if (orb_create(&p, &ctx) != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error(nullptr));
    w = width; h = height;
    return ctx;
  }
  ~TrackerContext() { if (ctx) orb_destroy(ctx); }
};
}  // namespace detail

// == cv::calcOpticalFlowPyrLK with plain images (no precomputed pyramids), flags 0
inline void calcOpticalFlowPyrLK(const cv::Mat& prevImg, const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts,
                                 std::vector<cv::Point2f>& nextPts, std::vector<unsigned char>& status, std::vector<float>& err,
                                 cv::Size winSize = cv::Size(21, 21), int maxLevel = 3,
                                 cv::TermCriteria criteria = cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01),
                                 int flags = 0, double minEigThreshold = 1e-4) {
  if (flags != 0) throw std::invalid_argument("orb_b200::calcOpticalFlowPyrLK: flags are not supported (the reference passes 0)");
  if (winSize.width != winSize.height) throw std::invalid_argument("orb_b200::calcOpticalFlowPyrLK: square windows only");
  if (prevImg.empty() || nextImg.empty() || prevImg.rows != nextImg.rows || prevImg.cols != nextImg.cols ||
      prevImg.type() != CV_8UC1 || nextImg.type() != CV_8UC1 || prevImg.step != nextImg.step)
    throw std::invalid_argument("orb_b200::calcOpticalFlowPyrLK: two CV_8UC1 images of one size and step");
  static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "Point2f is two floats");
  const int n = (int)prevPts.size();
  nextPts.resize(n); status.resize(n); err.resize(n);
  if (!n) return;
  // cv::TermCriteria: a missing COUNT means 30 iterations, a missing EPS means 0.01 (lkpyramid.cpp; tests/test_lk.py pins
  // the latter against cv2: COUNT-only criteria give exactly the tracks of COUNT+EPS with epsilon 0.01)
  const int max_iter = (criteria.type & cv::TermCriteria::COUNT) ? criteria.maxCount : 30;
  const double eps = (criteria.type & cv::TermCriteria::EPS) ? criteria.epsilon : 0.01;
  static thread_local detail::TrackerContext holder;
  orb_ctx* ctx = holder.get(prevImg.cols, prevImg.rows);
  const int rc = orb_lk_track(ctx, prevImg.ptr<unsigned char>(0), nextImg.ptr<unsigned char>(0), prevImg.cols, prevImg.rows, prevImg.step,
                              reinterpret_cast<const float*>(prevPts.data()), n, winSize.width, maxLevel, max_iter, eps,
                              (float)minEigThreshold, reinterpret_cast<float*>(nextPts.data()), status.data(), err.data());
  if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200::calcOpticalFlowPyrLK: ") + orb_last_error(ctx));
}

}  // namespace orb_b200
#endif
