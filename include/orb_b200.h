/* orb_b200.h -- C ABI of the B200-native ORB extractor (the only layer that touches CUDA).
 *
 * This is the drop-in boundary for the reference's hand-written ORB path
 * (WeeFav/Visual-Odometry-GPU).  Every entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  The reference crosses from C++ into
 * .cu files with cv::Mat / std::vector (include/Fast.cuh:5-6, include/Brief.cuh:5,
 * include/HarrisScore.cuh:5, include/NMS.cuh:5); here that seam is plain pointers and sizes,
 * so the C++ facade (include/orb.hpp in this repo), ctypes, cgo or JNI can all bind it.
 *
 * Conventions
 *  - all functions return ORB_OK (0) or a negative orb_status; no exit(), no exceptions
 *    (the reference's cudaCheckErrors macro prints and exit(1)s, src/cuda/Fast.cu:8-18);
 *  - images are 8-bit single channel, row-major, `pitch` bytes between rows (cv::Mat::step);
 *  - a context owns every device allocation (arena sized at create from max_* fields),
 *    its stream and its resize tables; nothing is allocated on the hot path
 *    (the reference cudaMallocs per call and never frees);
 *  - a context is not thread-safe; distinct contexts are independent (one per device/thread);
 *  - there is NO CPU fallback: without a CUDA device orb_create fails with ORB_E_CUDA.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_B200_ABI_VERSION 1

/* == Keypoint, reference include/orb.hpp:4 (x = column, y = row, level-0 coordinates) */
typedef struct { int32_t x, y; } orb_keypoint;
/* == ORBDescriptor, reference include/orb.hpp:6-8; bit i lives at data[i>>3] bit (i&7) */
typedef struct { uint8_t data[32]; } orb_descriptor;

typedef enum {
  ORB_OK = 0,
  ORB_E_INVALID = -1,      /* bad argument / unsupported parameter value          */
  ORB_E_CUDA = -2,         /* CUDA runtime error (see orb_last_error)             */
  ORB_E_CAPACITY = -3,     /* image / batch larger than the context was built for */
  ORB_E_OVERFLOW = -4,     /* more corner candidates than the arena holds         */
  ORB_E_NOMEM = -5,
  ORB_E_IO = -6,           /* a frame file could not be read                      */
  ORB_E_FORMAT = -7        /* a frame file is not a PNG this library decodes      */
} orb_status;

/* selection policy for the per-level keypoint cap */
#define ORB_SELECT_RASTER_FIRST_N 0 /* first N NMS survivors in raster order: reference src/orb_cpu.cpp:110 */
#define ORB_SELECT_HARRIS_TOP_N   1 /* top quota_l by Harris response:       reference src/orb.cpp:62-86    */

typedef struct {
  /* ORB ctor, reference include/orb.hpp:36 */
  int32_t nfeatures;        /* total keypoint budget (per frame)                           */
  float   scale_factor;     /* pyramid scale, 1.2f                                         */
  int32_t nlevels;          /* pyramid levels, 8                                           */
  /* OrientedFAST ctor, reference include/orb.hpp:12 / include/orb_cpu.hpp:6 */
  int32_t fast_threshold;   /* 20 (orb.hpp) or 50 (orb_cpu.hpp)                            */
  int32_t fast_n;           /* contiguous arc length, 9                                    */
  int32_t nms_window;       /* 3 (radius 1); 1 disables NMS                                */
  int32_t orient_patch;     /* 31 (orb.hpp) or 9 (orb_cpu.hpp)                             */
  int32_t select_policy;    /* ORB_SELECT_*                                                */
  int32_t blur_levels;      /* 1: resize+GaussianBlur5x5 (src/orb_cpu.cpp:288-289); 0: resize only (src/orb.cpp:119) */
  float   harris_k;         /* 0.04f (intended value at src/orb.cpp:65)                    */
  /* device / capacity (no reference counterpart: the reference allocates per call) */
  int32_t device;           /* CUDA device ordinal                                         */
  int32_t max_width, max_height;
  int32_t max_batch;        /* frames per orb_detect_and_compute_batch call                */
  int32_t chunk_frames;     /* frames processed per kernel wave (scratch kept L2-resident); 0 = auto */
  int32_t max_keypoints;    /* output slots per frame (>= what selection can return); 0 = nfeatures   */
  int32_t keep_side_arrays; /* 1: also record level-space xy / level id / response per output (diagnostics) */
  int32_t reserved[3];
} orb_params;

typedef struct orb_ctx orb_ctx;

/* fills *p with the reference defaults of include/orb.hpp (500 / 1.2f / 8, thr 20, n 9, nms 3,
 * patch 31), HARRIS_TOP_N, blurred pyramid, k = 0.04f, 1241x376, batch 1. */
void orb_default_params(orb_params* p);

/* replaces: ORB::ORB + OrientedFAST::OrientedFAST + RotatedBRIEF::RotatedBRIEF
 * (reference src/orb.cpp:9-20,35-38,46-56) and every per-call cudaMalloc below them. */
int  orb_create(const orb_params* p, orb_ctx** out);
void orb_destroy(orb_ctx* ctx);
const char* orb_last_error(const orb_ctx* ctx);   /* ctx may be NULL: error of the failed orb_create */
int  orb_abi_version(void);

/* run the context's work on a caller-provided cudaStream_t (passed as void*; NULL is the CUDA legacy
 * default stream, exactly as in the runtime API).  Lets a host framework order and time the kernels with
 * events on its own stream.  orb_use_own_stream goes back to the context's private non-blocking stream. */
int  orb_set_stream(orb_ctx* ctx, void* cuda_stream);
int  orb_use_own_stream(orb_ctx* ctx);

/* replaces: ORB::detectAndCompute (reference include/orb.hpp:37, src/orb.cpp:58-109; CPU twin
 * ORBCPU::detectAndCompute include/orb_cpu.hpp:31, src/orb_cpu.cpp:271-276).
 * img is a HOST pointer.  Writes at most `cap` records, levels concatenated 0..L-1, each level
 * in raster order.  n_per_level may be NULL (else nlevels ints).  Synchronous. */
int  orb_detect_and_compute(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
                            int cap, orb_keypoint* kps, float* angles, orb_descriptor* desc,
                            int* n_out, int* n_per_level);

/* batch form of the same call: n_frames images of identical shape.
 *  frames_on_device = 0: host frames are copied in; 1: `frames` is a device pointer.
 *  outputs_on_device = 0: kps/angles/desc/n_out are host buffers ([n_frames][cap] records,
 *  n_out[n_frames]); 1: they are device buffers and the call returns without synchronising.
 * Frames are independent (reference src/orb.cpp:58-109 keeps no state between calls). */
int  orb_detect_and_compute_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device,
                                  int n_frames, int w, int h, size_t pitch, size_t frame_stride,
                                  int cap, orb_keypoint* kps, float* angles, orb_descriptor* desc,
                                  int* n_out, int outputs_on_device);

/* ---- stage entry points (one image, host pointers, synchronous) ---------------------------
 * They exist so that the facade's per-stage methods and the parity tests can exercise each
 * kernel against the matching oracle stage. */

/* replaces: ORBCPU::buildPyramid / ORB::buildPyramid (reference src/orb_cpu.cpp:278-290,
 * src/orb.cpp:111-120).  Pyramid pixels of `level` of frame `frame` of the LAST batch call. */
int  orb_get_level(orb_ctx* ctx, int frame, int level, uint8_t* dst, size_t dst_pitch, int* w, int* h);
/* level geometry and per-level quota (reference src/orb_cpu.cpp:284-285, src/orb.cpp:62) */
int  orb_level_size(const orb_ctx* ctx, int w, int h, int level, int* lw, int* lh);
int  orb_level_quota(const orb_ctx* ctx, int level);

/* replaces: Fast() + d_NMS (reference include/Fast.cuh:5, src/cuda/Fast.cu:211-270,
 * src/cuda/NMS.cu:21-128) == OrientedFASTCPU::detect (src/orb_cpu.cpp:23-137).
 * FAST-n segment test, SAD score, 3x3 NMS; first `nfeatures` survivors in raster order. */
int  orb_fast_detect(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
                     int nfeatures, orb_keypoint* kps, int* n_out);
/* replaces: NMS() over a caller's score map (reference include/NMS.cuh:5, src/cuda/NMS.cu:21-164): `scores` is a host
 * float map (w x h, pitch_bytes between rows).  A pixel at least nms_window/2 inside the map is kept iff its score
 * exceeds `threshold` and no score of its window is strictly greater (ties keep both).  The reference appends survivors
 * in the arrival order of a global atomic and drops what exceeds nfeatures; here the first `nfeatures` survivors in
 * raster order are returned (deterministic). */
int  orb_nms_scores(orb_ctx* ctx, const float* scores, int w, int h, size_t pitch_bytes, int nms_window, int nfeatures,
                    float threshold, orb_keypoint* kps, int* n_out);
/* replaces: conv2d() (reference include/Convolution.cuh:5, src/cuda/Convolution.cu:57-103) and, with it, what GaussianBlurCUDA
 * (src/GaussianBlur.cpp:39-49), SobelCUDA (src/Sobel.cpp:18-31) and GaussianBlur (src/cuda/GaussianBlur.cu:73-130) do around
 * it: K x K correlation of the u8 image promoted to float (one FMA per tap, row by row), optionally after extending the image
 * by K/2 with BORDER_REFLECT_101 (output then has the input's size, else valid mode: (w-K+1) x (h-K+1)), optionally divided by
 * `divisor` (0 = no division), converted like cv::Mat::convertTo(CV_8U).  `kernel` is a host array of ksize*ksize floats. */
int  orb_conv2d_u8(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, const float* kernel, int ksize,
                   int border_reflect101, float divisor, uint8_t* out, size_t out_pitch);
/* replaces: GaussianBlur1D() (reference include/GaussianBlur.cuh:4, src/cuda/GaussianBlur1D.cu:108-166): separable [1 4 6 4 1]/16,
 * BORDER_REFLECT_101, float, convertTo(CV_8U) */
int  orb_gaussian_blur_1d(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, uint8_t* out, size_t out_pitch);
/* replaces: HarrisScore() (reference include/HarrisScore.cuh:5, src/cuda/HarrisScore.cu:42-89) */
int  orb_harris(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
                const orb_keypoint* kps, int n, float* response);
/* replaces: Orientations() (reference include/Fast.cuh:6, src/cuda/Orientations.cu:65-100)
 * == OrientedFASTCPU::compute_orientations (src/orb_cpu.cpp:139-183) */
int  orb_orientations(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
                      const orb_keypoint* kps, int n, float* angles);
/* replaces: Brief() (reference include/Brief.cuh:5, src/cuda/Brief.cu:97-137)
 * == RotatedBRIEFCPU::compute (src/orb_cpu.cpp:203-258) */
int  orb_brief(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
               const orb_keypoint* kps, const float* angles, int n, orb_descriptor* desc);

/* ---- descriptor matching (the step right after the descriptors) ----------
 * replaces: flann->knnMatch(des1, des2, matches, 2) of the VO loops (reference src/feature_matching.cpp:168,
 * src/feature_tracking.cpp:205) by an exact brute-force Hamming 2-nearest-neighbour search on the device (the reference's
 * FLANN-LSH index is approximate and randomised; ties here go to the lower train index).  Descriptors stay on the device
 * when they were produced with outputs_on_device = 1. */
typedef struct { int32_t idx1, dist1, idx2, dist2; } orb_match;   /* absent neighbour: idx -1, dist INT32_MAX */
int  orb_match_knn2(orb_ctx* ctx, const orb_descriptor* query, int nq, const orb_descriptor* train, int nt,
                    int on_device, orb_match* out /* [nq], same memory space as the inputs */);
/* all consecutive frame pairs of a batch: pair p = frame p (query, n[p] descriptors) against frame p+1 (train);
 * desc is [n_frames][cap], out is [n_frames-1][cap] (entries >= n[p] untouched) */
int  orb_match_knn2_batch(orb_ctx* ctx, const orb_descriptor* desc, const int* n, int n_frames, int cap, int on_device,
                          orb_match* out);
/* the reference's ratio test, src/feature_matching.cpp:174-182: keep[i] = has two neighbours && dist1 < ratio * dist2 */
void orb_ratio_test(const orb_match* m, int n, float ratio, uint8_t* keep);

/* ---- frame ingest: the step before the path ---------------------------------------------------
 * Replaces cv::imread(path, cv::IMREAD_GRAYSCALE) of the reference's VO loops (src/feature_matching.cpp:55,59;
 * src/feature_tracking.cpp:56,196).  Decodes non-interlaced PNG of bit depth 8 / 16, gray, gray+alpha, RGB, RGBA to
 * 8-bit gray with the results of OpenCV 4.13's imread (libpng's rgb_to_gray coefficients, 16 -> 8 by the high byte);
 * palette, interlaced and sub-byte files are rejected with ORB_E_FORMAT.  Chunk CRCs and the zlib checksum are
 * verified, corrupt files fail like they do in libpng. */
typedef struct orb_image_info { int32_t width, height, bit_depth, channels; } orb_image_info;
/* header of an encoded PNG held in memory (ctx-less; message via orb_last_error(NULL)) */
int  orb_png_info(const uint8_t* file, size_t file_bytes, orb_image_info* info);
/* host decode into a caller buffer of `pitch` bytes per row (ctx-less; w, h must equal the file's size) */
int  orb_png_decode_gray8(const uint8_t* file, size_t file_bytes, uint8_t* dst, size_t pitch, int w, int h);
/* == cv::imread(path, IMREAD_GRAYSCALE) into a caller buffer of cap_w x cap_h at `pitch`; size returned in w, h */
int  orb_imread_gray8(const char* path, uint8_t* dst, size_t pitch, int cap_w, int cap_h, int* w, int* h);
/* detect-and-compute over a list of PNG files of one size: n_threads host threads (0 = all cores) read and decode
 * the files straight into the context's pinned staging area while earlier waves of frames are copied and processed
 * on the device; outputs as orb_detect_and_compute_batch.  A frame that fails to load aborts the call with
 * ORB_E_IO / ORB_E_FORMAT and names the file in orb_last_error.
 * decode_on_device = 1 moves inflate, unfilter and the checksum tests to the GPU for 8-bit gray files (the host only
 * reads the files and uploads the compressed bytes); files of other layouts make the call fail with ORB_E_FORMAT. */
int  orb_detect_and_compute_files(orb_ctx* ctx, const char* const* paths, int n_frames, int n_threads,
                                  int decode_on_device, int cap, orb_keypoint* kps, float* angles, orb_descriptor* desc,
                                  int* n_out, int outputs_on_device);
/* test hook of the device inflate kernel: n raw deflate streams (RFC 1951, no zlib header) concatenated in `streams`
 * (stream i = bytes [offsets[i], offsets[i+1])), inflated into out[out_offsets[i] .. out_offsets[i+1]) which must be
 * their exact sizes; status[i] = 0 or 1 corrupt / 2 size mismatch / 3 truncated / 4 table overflow.  Host pointers. */
int  orb_debug_inflate(orb_ctx* ctx, const uint8_t* streams, const uint32_t* offsets, int n, uint8_t* out,
                       const uint32_t* out_offsets, int* status);
/* the decoded frame `frame` of the last orb_detect_and_compute_files call (level 0 as the kernels saw it) */
int  orb_get_ingested_frame(orb_ctx* ctx, int frame, uint8_t* dst, size_t dst_pitch, int* w, int* h);

/* ---- pyramidal Lucas-Kanade tracker: the other front end of the reference's VO loop -------------
 * == cv::calcOpticalFlowPyrLK(prev, next, prev_pts, next_pts, status, err, Size(win, win), max_level,
 *                             TermCriteria(COUNT+EPS, max_iter, eps), 0, min_eig)
 * as called at reference src/feature_tracking.cpp:174-180 with win 21, max_level 3, 30 iterations, eps 0.01, min_eig 0.001.
 * Follows OpenCV 4.x lkpyramid.cpp (pyrDown pyramid, Scharr derivatives with a zero border, 14-bit fixed-point window
 * samples, float normal equations, the same termination rules and L1 error).  Host pointers: 8-bit frames of w x h at
 * `pitch`, prev_pts / next_pts [n][2] float (x, y), status [n], err [n] (may be NULL).  win in [3, 33]. */
int  orb_lk_track(orb_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h, size_t pitch, const float* prev_pts,
                  int n, int win, int max_level, int max_iter, double eps, float min_eig, float* next_pts, uint8_t* status,
                  float* err);
/* The same for a batch whose frames stay on the device (the VO loop of reference src/feature_tracking.cpp:160-225 run over a
 * sequence): frame t is tracked into frame t + 1, t = 0 .. n_frames - 2, every pyramid built once, one launch for all points.
 * prev_pts / next_pts [n_frames - 1][cap][2], status / err [n_frames - 1][cap]; n_pts [n_frames - 1] (NULL: cap points in every
 * pair) -- all in device memory if pts_on_device, else in host memory (results are then complete on return; with device
 * pointers the call is asynchronous on the context's stream).  Entries beyond a pair's count are left untouched. */
int  orb_lk_track_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h, size_t pitch,
                        size_t frame_stride, const float* prev_pts, const int* n_pts, int cap, int pts_on_device, int win,
                        int max_level, int max_iter, double eps, float min_eig, float* next_pts, uint8_t* status, float* err);
/* highest pyramid level index the tracker uses for this frame size (OpenCV stops before a level <= the window) */
int  orb_lk_levels(int w, int h, int win, int max_level);
/* pyramid level of the last orb_lk_track call (which: 0 = prev, 1 = next), packed rows; size returned in w, h */
int  orb_lk_get_level(orb_ctx* ctx, int which, int level, uint8_t* dst, int* w, int* h);

/* ---- side arrays of the LAST orb_detect_and_compute[_batch] call (testing / diagnostics) ----
 * level-space coordinates, level id and Harris response of output record i of `frame`
 * (the reference drops them, src/orb.cpp:94-102).  Any pointer may be NULL. */
int  orb_get_side_arrays(orb_ctx* ctx, int frame, int n, orb_keypoint* level_xy,
                         int32_t* level_id, float* response);
/* all NMS survivors of (frame, level) of the last call, unordered: coordinates + response */
int  orb_get_candidates(orb_ctx* ctx, int frame, int level, int cap, orb_keypoint* xy,
                        float* response, int* n_out);
/* the 49 Harris window weights the context uses (createGaussianKernel(7), reference
 * src/GaussianBlur.cpp:7-37) */
int  orb_get_harris_weights(const orb_ctx* ctx, float* w49);
/* wait for the context's stream and report deferred errors (candidate overflow) of calls made with
 * outputs_on_device = 1 */
int  orb_synchronize(orb_ctx* ctx);
/* libm twins used on the device, evaluated on host arrays (test hook): op 0 atan2f(a,b), 1 cosf(a),
 * 2 sinf(a), 3 lround(a) -- the glibc calls of reference src/orb_cpu.cpp:178,217-218,228-232 */
int  orb_debug_eval_math(orb_ctx* ctx, int op, const float* a, const float* b, int n, float* out);
/* host logic of the batch pipeline, callable without a GPU: wave boundaries for a batch of n_frames with waves of `wave`
 * frames (ramp_up: frames staged from the host; ramp_down: results copied back to the host).  Writes up to cap begins
 * (the last one is n_frames) and returns how many there are. */
int  orb_debug_wave_schedule(int n_frames, int wave, int ramp_up, int ramp_down, int* begins, int cap);

/* Bounds-check builds (-DORB_BOUNDS_CHECK: every shared / global index of the ORB kernels is checked before use; stands in
 * for compute-sanitizer, which is closed on the B200 pool): *enabled = 1 in such a build, *failures = failed checks since the
 * library was loaded, *first_line = source line (orb_kernels.cuh) of the first one, *kernels_checked = CTAs that ran checks. */
int  orb_debug_bounds_check(orb_ctx* ctx, int* enabled, unsigned* failures, unsigned* first_line, unsigned* kernels_checked);
/* launches a kernel in which exactly one check fails (a bounds-check build then reports one more failure) */
int  orb_debug_bounds_selftest(orb_ctx* ctx);
/* per-kernel timing (bench / roofline accounting): when enabled, every kernel launch of the detect calls is
 * bracketed by CUDA events on the context's stream.  orb_get_stage_ms waits for the stream and returns, for the
 * launches since the last call, the summed device time [ms] and launch count of
 * stage 0 = pyramid kernel (resize+blur), 1 = FAST+NMS+box-sum kernel, 2 = Harris kernel, 3 = selection kernel,
 * 4 = orientation+BRIEF kernel. */
int  orb_set_profiling(orb_ctx* ctx, int enable);
int  orb_get_stage_ms(orb_ctx* ctx, float ms[5], int launches[5]);
/* number of kernel launches issued by the last detect call (for bench accounting) */
int  orb_last_launch_count(const orb_ctx* ctx);

/* == extern int bit_pattern_31_[256*4], reference include/orb_pattern.hpp:2 */
extern int bit_pattern_31_[256 * 4];

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
