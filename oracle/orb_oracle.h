/* orb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's ORB path (WeeFav/Visual-Odometry-GPU), used as the parity
 * checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (visual-odometry-gpu_b200/) never does.
 *
 * Parity pinning: the reference has no tests and no golden vectors.  This oracle is
 * pinned (tests/test_oracle_*.py) against
 *   (1) the reference's own src/orb_cpu.cpp + src/orb_pattern.cpp compiled UNMODIFIED against a
 *       minimal opencv2 shim (oracle/cv_shim, recipe oracle/Makefile -> oracle/_ref/liborbcpu_ref.so)
 *       for the single-level mode the reference ships (ORBCPU::detectAndCompute);
 *   (2) Python cv2 4.13 for the three OpenCV primitives (resize INTER_LINEAR, GaussianBlur 5x5,
 *       integral);
 *   (3) the committed fixtures under tests/golden/ generated from (1)+(2) by tools/gen_golden.py.
 * The multi-level composition (decisions D1-D10 of SURVEY.md 8(c)) is ours; the reference's GPU
 * facade it follows (src/orb.cpp:58-109) does not build as shipped.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t x, y; } orc_keypoint;          /* reference include/orb.hpp:4   */
typedef struct { uint8_t data[32]; } orc_descriptor;    /* reference include/orb.hpp:6-8 */

typedef struct {
  int32_t nfeatures; float scale_factor; int32_t nlevels;     /* include/orb.hpp:36       */
  int32_t fast_threshold, fast_n, nms_window, orient_patch;   /* include/orb.hpp:12       */
  int32_t select_policy;   /* 0 raster-first-N (orb_cpu.cpp:110), 1 Harris top-N (orb.cpp:62-86) */
  int32_t blur_levels;     /* 1: orb_cpu.cpp:288-289, 0: orb.cpp:119                       */
  float   harris_k;
} orc_params;

/* geometry: reference src/orb_cpu.cpp:284-285, src/orb.cpp:62 */
void orc_level_size(int W, int H, float f, int level, int* w, int* h);
float orc_level_scale(float f, int level);
int  orc_level_quota(int nfeatures, float f, int nlevels, int level);

/* OpenCV primitives restated (cv::resize INTER_LINEAR, cv::GaussianBlur 5x5 sigma 0, cv::integral) */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sp, uint8_t* dst, int dw, int dh, size_t dp);
void orc_gauss5x5_u8(const uint8_t* src, int w, int h, size_t sp, uint8_t* dst, size_t dp);
/* (h+4) x (w+1) int32: rows 0..h are cv::integral, rows h+1..h+3 are zero (decision D7) */
void orc_integral_flat(const uint8_t* src, int w, int h, size_t sp, int32_t* out);
/* resize tables exactly as the restated resize uses them (for the product's table test) */
void orc_resize_table(int src, int dst, int32_t* ofs, int16_t* a0, int16_t* a1);

/* pyramid level l of an image (l = 0 copies) */
void orc_build_level(const uint8_t* img, int W, int H, size_t pitch, const orc_params* p, int level,
                     uint8_t* dst /* w*h, pitch w */);

/* stages: reference src/orb_cpu.cpp:23-103, :105-134, :139-183, :203-258 */
void orc_fast_scores(const uint8_t* img, int w, int h, size_t pitch, int thr, int n, float* scores /* w*h */);
int  orc_nms(const float* scores, int w, int h, int nms_window, int cap, orc_keypoint* kps);
void orc_harris_weights(float* w49);  /* createGaussianKernel(7): reference src/GaussianBlur.cpp:7-37 */
/* the reference's filter wrappers (src/cuda/Convolution.cu, src/cuda/GaussianBlur.cu, src/cuda/GaussianBlur1D.cu, src/GaussianBlur.cpp, src/Sobel.cpp) */
void orc_gaussian_kernel(int ksize, float* kernel);
void orc_conv2d_u8(const uint8_t* img, int w, int h, size_t pitch, const float* kernel, int ksize, int reflect, float divisor, uint8_t* out);
void orc_gaussian_blur_1d(const uint8_t* img, int w, int h, size_t pitch, uint8_t* out);
void orc_harris(const uint8_t* img, int w, int h, size_t pitch, const orc_keypoint* kps, int n, float k, float* out);
void orc_orientations(const uint8_t* img, int w, int h, size_t pitch, const orc_keypoint* kps, int n, int patch, float* out);
void orc_brief(const uint8_t* img, int w, int h, size_t pitch, const orc_keypoint* kps, const float* angles, int n,
               orc_descriptor* out);
/* per keypoint diagnostics of the BRIEF bound rule: bit0 = some test skipped, bit1 = some box read
 * outside the (h+1)x(w+1) integral (the reads the reference leaves undefined) */
void orc_brief_flags(int w, int h, const orc_keypoint* kps, const float* angles, int n, uint8_t* flags);

/* whole path.  Outputs (any side pointer may be NULL): records in output order (levels 0..L-1,
 * raster order inside a level); n_per_level[nlevels]; level-space xy; level id; Harris response
 * (0 when policy is raster-first-N).  Returns the number of records (<= cap). */
int  orc_detect_and_compute(const uint8_t* img, int W, int H, size_t pitch, const orc_params* p, int cap,
                            orc_keypoint* kps, float* angles, orc_descriptor* desc, int* n_per_level,
                            orc_keypoint* level_xy, int32_t* level_id, float* response);
/* frame-parallel driver for CPU-baseline timing: n_threads std::threads, contiguous frame blocks.
 * n_out[n_frames]; record buffers are [n_frames][cap] and may be NULL (results discarded). */
int  orc_detect_and_compute_batch(const uint8_t* frames, int n_frames, size_t frame_stride, int W, int H, size_t pitch,
                                  const orc_params* p, int cap, orc_keypoint* kps, float* angles, orc_descriptor* desc,
                                  int* n_out, int n_threads);

/* exact brute-force Hamming 2-NN (the deterministic counterpart of flann->knnMatch(des1, des2, matches, 2),
 * reference src/feature_matching.cpp:168): out[i] = {idx1, dist1, idx2, dist2}, ties to the lower train index,
 * absent neighbour = {-1, INT32_MAX}; keep[i] = ratio test of src/feature_matching.cpp:178 (may be NULL) */
void orc_match_knn2(const orc_descriptor* query, int nq, const orc_descriptor* train, int nt, int32_t* out4, float ratio,
                    uint8_t* keep);

/* pyramidal Lucas-Kanade == cv::calcOpticalFlowPyrLK(prev, next, prev_pts, next_pts, status, err, Size(win, win), max_level,
 * TermCriteria(COUNT+EPS, max_iter, eps), 0, min_eig), the call of reference src/feature_tracking.cpp:174-180; OpenCV 4.x
 * lkpyramid.cpp restated (scalar path, fixed summation order).  prev_pts / next_pts are [n][2] floats. */
void orc_lk_track(const uint8_t* prev, const uint8_t* next, int w, int h, size_t pitch, const float* prev_pts, int n, int win,
                  int max_level, int max_iter, double eps, float min_eig, float* next_pts, uint8_t* status, float* err);
int  orc_lk_levels(int w, int h, int win, int max_level);
void orc_lk_pyr_down(const uint8_t* src, int w, int h, size_t pitch, uint8_t* dst /* ((w+1)/2) x ((h+1)/2), packed */);

/* helpers for property tests */
long orc_lround_f(float v);                       /* std::lround(float)                   */
float orc_atan2f(float y, float x);
float orc_cosf(float a);
float orc_sinf(float a);
const int8_t* orc_pattern(void);                  /* 1024 int8, include/orb_brief_pattern.h */

#ifdef __cplusplus
}
#endif
#endif
