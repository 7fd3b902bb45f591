// orb_stages.hpp -- the stage-level free functions of the reference's host <-> CUDA seam, over the C ABI:
//   int  Fast(image, keypoints, threshold, n, nms_window, nfeatures)                 reference include/Fast.cuh:5
//   void Orientations(image, keypoints, orientations, patch_size)                    reference include/Fast.cuh:6
//   void Brief(image, keypoints, orientations, descriptors, n_bits, patch_size)      reference include/Brief.cuh:5
//   void HarrisScore(image, keypoints, harris_scores, corner_window, k)              reference include/HarrisScore.cuh:5
//   void NMS(score_map, keypoints, nms_window, nfeatures, threshold)                 reference include/NMS.cuh:5
// Same names and argument lists, so a translation unit that includes the reference's Fast.cuh / Brief.cuh / HarrisScore.cuh
// (src/orb.cpp:24,31,42,65) includes this header instead.  Each call forwards to one C-ABI stage entry point
// (orb_fast_detect / orb_orientations / orb_brief / orb_harris); contexts are per thread and per parameter set and grow
// with the image.  Differences from the reference, all documented in DESIGN.md: Fast() returns the first `nfeatures` NMS
// survivors in raster order (the reference's order is the arrival order of a global atomic); Brief() ignores n_bits and
// patch_size exactly like the reference does (src/cuda/Brief.cu:97-137: always 256 bits, pattern 31); HarrisScore() computes
// the response the reference intends (decision D5: Sobel 3x3 reflect-101, 7x7 Gaussian window sigma 1.7).  Its last parameter
// exists twice: the reference's exact `int k` signature (include/HarrisScore.cuh:5), which uses the integer it is given, and
// a floating-point overload.  At the reference's call site `HarrisScore(pyramid[l], kps, scores, 7, 0.04)` (src/orb.cpp:65) the
// floating-point overload WINS overload resolution (exact match for the double literal), so k = 0.04 as the call site means;
// the reference's own declaration would truncate it to 0 -- recorded deviation, DESIGN.md D5.  NMS() returns the first
// `nfeatures` survivors in raster order, like Fast().
// conv2d(), GaussianBlur*(), SobelCUDA() are internal steps of HarrisScore in this implementation and have no stand-alone
// entry point.
#ifndef ORB_STAGES_HPP
#define ORB_STAGES_HPP

#include <map>
#include <memory>
#include <tuple>
#include <type_traits>

#include "orb.hpp"

namespace orb_b200_detail {
// one Handle per (threshold, n, nms_window, patch) and thread
inline Handle& stage_handle(int threshold, int n, int nms_window, int patch_size, float harris_k = 0.04f) {
    typedef std::tuple<int, int, int, int, float> Key;
    static thread_local std::map<Key, std::unique_ptr<Handle>> cache;
    std::unique_ptr<Handle>& h = cache[Key(threshold, n, nms_window, patch_size, harris_k)];
    if (!h) {
        orb_params p = make_params(8192, 1.2f, 1, threshold, n, nms_window, patch_size, ORB_SELECT_RASTER_FIRST_N);
        p.harris_k = harris_k;
        h.reset(new Handle(p));
    }
    return *h;
}
}  // namespace orb_b200_detail

inline int Fast(const cv::Mat& image, std::vector<Keypoint>& keypoints, int threshold, int n, int nms_window, int nfeatures) {
    orb_b200_detail::check_image(image);
    orb_b200_detail::Handle& h = orb_b200_detail::stage_handle(threshold, n, nms_window, 31);
    keypoints.resize(nfeatures > 0 ? nfeatures : 0);
    int count = 0;
    if (nfeatures > 0)
        h.check(orb_fast_detect(h.get(image.cols, image.rows), image.data, image.cols, image.rows, image.step, nfeatures,
                                reinterpret_cast<orb_keypoint*>(keypoints.data()), &count));
    keypoints.resize(count);
    return count;
}

inline void Orientations(const cv::Mat& image, const std::vector<Keypoint>& keypoints, std::vector<float>& orientations, int patch_size) {
    orb_b200_detail::check_image(image);
    orb_b200_detail::Handle& h = orb_b200_detail::stage_handle(20, 9, 3, patch_size);
    orientations.resize(keypoints.size());
    h.check(orb_orientations(h.get(image.cols, image.rows), image.data, image.cols, image.rows, image.step,
                             reinterpret_cast<const orb_keypoint*>(keypoints.data()), (int)keypoints.size(), orientations.data()));
}

inline void Brief(const cv::Mat& image, const std::vector<Keypoint>& keypoints, const std::vector<float>& orientations,
                  std::vector<ORBDescriptor>& descriptors, int /*n_bits*/, int /*patch_size*/) {
    orb_b200_detail::check_image(image);
    if (orientations.size() != keypoints.size()) throw std::runtime_error("orb_b200: Brief needs one orientation per keypoint");
    orb_b200_detail::Handle& h = orb_b200_detail::stage_handle(20, 9, 3, 31);
    descriptors.resize(keypoints.size());
    h.check(orb_brief(h.get(image.cols, image.rows), image.data, image.cols, image.rows, image.step,
                      reinterpret_cast<const orb_keypoint*>(keypoints.data()), orientations.data(), (int)keypoints.size(),
                      reinterpret_cast<orb_descriptor*>(descriptors.data())));
}

namespace orb_b200_detail {
inline void harris_score(const cv::Mat& image, std::vector<Keypoint>& keypoints, std::vector<float>& harris_scores, int corner_window, float k);
}
// the reference's exact signature (include/HarrisScore.cuh:5): k is used as the integer it is
inline void HarrisScore(const cv::Mat& image, std::vector<Keypoint>& keypoints, std::vector<float>& harris_scores,
                        int corner_window, int k) {
    orb_b200_detail::harris_score(image, keypoints, harris_scores, corner_window, (float)k);
}
// floating-point k: what `HarrisScore(..., 7, 0.04)` (src/orb.cpp:65) resolves to
template <typename T, typename std::enable_if<std::is_floating_point<T>::value, int>::type = 0>
inline void HarrisScore(const cv::Mat& image, std::vector<Keypoint>& keypoints, std::vector<float>& harris_scores,
                        int corner_window, T k) {
    orb_b200_detail::harris_score(image, keypoints, harris_scores, corner_window, (float)k);
}

inline void orb_b200_detail::harris_score(const cv::Mat& image, std::vector<Keypoint>& keypoints, std::vector<float>& harris_scores,
                                          int corner_window, float k) {
    orb_b200_detail::check_image(image);
    if (corner_window != 7) throw std::runtime_error("orb_b200: HarrisScore supports the 7x7 window of the reference's call site");
    orb_b200_detail::Handle& h = orb_b200_detail::stage_handle(20, 9, 3, 31, k);
    harris_scores.resize(keypoints.size());
    h.check(orb_harris(h.get(image.cols, image.rows), image.data, image.cols, image.rows, image.step,
                       reinterpret_cast<const orb_keypoint*>(keypoints.data()), (int)keypoints.size(), harris_scores.data()));
}


// NMS over a caller's CV_32F score map (reference include/NMS.cuh:5; src/cuda/NMS.cu:130-164)
inline void NMS(const cv::Mat& input, std::vector<Keypoint>& keypoints, int nms_window, int nfeatures, float threshold) {
    if (input.empty() || input.type() != CV_32FC1) throw std::runtime_error("orb_b200: NMS needs a CV_32FC1 score map");
    orb_b200_detail::Handle& h = orb_b200_detail::stage_handle(20, 9, 3, 31);
    keypoints.resize(nfeatures > 0 ? nfeatures : 0);
    int count = 0;
    if (nfeatures > 0)
        h.check(orb_nms_scores(h.get(input.cols, input.rows), reinterpret_cast<const float*>(input.data), input.cols, input.rows, input.step,
                               nms_window, nfeatures, threshold, reinterpret_cast<orb_keypoint*>(keypoints.data()), &count));
    keypoints.resize(count);
}

#endif
