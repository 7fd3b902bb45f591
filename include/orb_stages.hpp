// orb_stages.hpp -- the stage-level free functions of the reference's host <-> CUDA seam, over the C ABI:
//   int  Fast(image, keypoints, threshold, n, nms_window, nfeatures)                 reference include/Fast.cuh:5
//   void Orientations(image, keypoints, orientations, patch_size)                    reference include/Fast.cuh:6
//   void Brief(image, keypoints, orientations, descriptors, n_bits, patch_size)      reference include/Brief.cuh:5
//   void HarrisScore(image, keypoints, harris_scores, corner_window, k)              reference include/HarrisScore.cuh:5
//   void NMS(score_map, keypoints, nms_window, nfeatures, threshold)                 reference include/NMS.cuh:5
//   int  conv2d(image, dst, kernel, kernel_size)                                     reference include/Convolution.cuh:5
//   void GaussianBlur(image, dst), GaussianBlur1D(image, dst)                        reference include/GaussianBlur.cuh:3-4
//   void GaussianBlurCUDA(image, dst, kernel_size)                                   reference include/GaussianBlur.hpp:6
//   void SobelCUDA(image, dst, dir)                                                  reference include/Sobel.hpp:6
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
// conv2d(), GaussianBlur*(), SobelCUDA() keep the reference's arithmetic (u8 promoted to float, one FMA per tap row by row,
// cv::Mat::convertTo(CV_8U) at the end, BORDER_REFLECT_101 where the reference pads); the Harris response itself does not go
// through them (k_harris evaluates Sobel and the 7x7 window at the candidates only).
#ifndef ORB_STAGES_HPP
#define ORB_STAGES_HPP

#include <cmath>
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


// ---- the reference's stand-alone filters ---------------------------------------------------------------------------------
namespace orb_b200_detail {
inline void filter_u8(const cv::Mat& image, cv::Mat& dst, const float* kernel, int ksize, bool reflect, float divisor) {
    check_image(image);
    const int ow = reflect ? image.cols : image.cols - ksize + 1, oh = reflect ? image.rows : image.rows - ksize + 1;
    if (ow < 1 || oh < 1) throw std::runtime_error("orb_b200: image smaller than the kernel");
    Handle& h = stage_handle(20, 9, 3, 31);
    cv::Mat out;
    out.create(oh, ow, CV_8UC1);                        // (dst may alias image)
    h.check(orb_conv2d_u8(h.get(image.cols, image.rows), image.data, image.cols, image.rows, image.step, kernel, ksize, reflect ? 1 : 0,
                          divisor, out.data, out.step));
    dst = out;
}
}  // namespace orb_b200_detail

// valid-mode K x K correlation of a (pre-padded) image -> CV_8U (reference src/cuda/Convolution.cu:57-103; its header declares
// `int`, its definition `void`: 0 is returned)
inline int conv2d(const cv::Mat& image, cv::Mat& dst, float* kernel, int kernel_size) {
    orb_b200_detail::filter_u8(image, dst, kernel, kernel_size, false, 0.0f);
    return 0;
}
// 5x5 Gaussian {1 4 7 4 1; 4 16 26 16 4; 7 26 41 26 7; ...} / 273, BORDER_REFLECT_101 (reference src/cuda/GaussianBlur.cu:21-130)
inline void GaussianBlur(const cv::Mat& image, cv::Mat& dst) {
    static const float k[25] = {1, 4, 7, 4, 1, 4, 16, 26, 16, 4, 7, 26, 41, 26, 7, 4, 16, 26, 16, 4, 1, 4, 7, 4, 1};
    orb_b200_detail::filter_u8(image, dst, k, 5, true, 273.0f);
}
// separable [1 4 6 4 1] / 16, BORDER_REFLECT_101 (reference src/cuda/GaussianBlur1D.cu:108-166)
inline void GaussianBlur1D(const cv::Mat& image, cv::Mat& dst) {
    orb_b200_detail::check_image(image);
    orb_b200_detail::Handle& h = orb_b200_detail::stage_handle(20, 9, 3, 31);
    cv::Mat out;
    out.create(image.rows, image.cols, CV_8UC1);
    h.check(orb_gaussian_blur_1d(h.get(image.cols, image.rows), image.data, image.cols, image.rows, image.step, out.data, out.step));
    dst = out;
}
// createGaussianKernel(kernel_size) + BORDER_REFLECT_101 + conv2d (reference src/GaussianBlur.cpp:7-49)
inline void GaussianBlurCUDA(const cv::Mat& image, cv::Mat& dst, int kernel_size) {
    if (kernel_size < 1 || kernel_size % 2 == 0) throw std::runtime_error("orb_b200: kernel size must be odd");
    std::vector<float> kernel((size_t)kernel_size * kernel_size);
    const float sigma = 0.3f * ((kernel_size - 1) * 0.5f) + 0.8f;       // the reference's heuristic (src/GaussianBlur.cpp:15-16)
    const int half = kernel_size / 2;
    float sum = 0.0f;
    for (int y = -half; y <= half; ++y)
        for (int x = -half; x <= half; ++x) {
            const float value = std::exp(-(x * x + y * y) / (2 * sigma * sigma));
            kernel[(size_t)(y + half) * kernel_size + (x + half)] = value;
            sum += value;
        }
    for (float& v : kernel) v /= sum;
    orb_b200_detail::filter_u8(image, dst, kernel.data(), kernel_size, true, 0.0f);
}
// 3x3 Sobel, dir 0 = x, else y, BORDER_REFLECT_101, result CV_8U (negative responses saturate to 0, as in the reference:
// src/Sobel.cpp:6-31 -> conv2d's convertTo(CV_8U))
inline void SobelCUDA(const cv::Mat& image, cv::Mat& dst, int dir) {
    static const float sx[9] = {-1.f, 0.f, 1.f, -2.f, 0.f, 2.f, -1.f, 0.f, 1.f}, sy[9] = {-1.f, -2.f, -1.f, 0.f, 0.f, 0.f, 1.f, 2.f, 1.f};
    orb_b200_detail::filter_u8(image, dst, dir == 0 ? sx : sy, 3, true, 0.0f);
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
