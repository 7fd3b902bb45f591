// orb.hpp -- drop-in for the reference's include/orb.hpp (WeeFav/Visual-Odometry-GPU) on top of the C ABI
// of the B200-native extractor (include/orb_b200.h, liborb_b200.so).
//
// Same names, constructor defaults, method signatures and record layouts as the reference header:
//   struct Keypoint {int x, y;}                          reference include/orb.hpp:4
//   struct ORBDescriptor {uint8_t data[32];}             reference include/orb.hpp:6-8
//   class OrientedFAST(threshold=20, n=9, nms_window=3, patch_size=31)      :10-22
//   class RotatedBRIEF()                                                     :24-32
//   class ORB(nfeatures=500, scaleFactor=1.2f, nlevels=8)                    :34-49
// so src/compare.cpp (and later feature_matching.cpp / feature_tracking.cpp) compile against it unchanged and
// link liborb_b200.so instead of src/orb.cpp + src/cuda/*.cu.  Unlike the reference header this one is
// self-contained (it includes what it uses).  Header-only: every method forwards to one C-ABI call; errors
// become std::runtime_error (the reference exit(1)s, src/cuda/Fast.cu:8-18).
//
// Semantics kept from the reference: ORB::detectAndCompute APPENDS to the caller's vectors (src/orb.cpp:100-102);
// keypoints are level-0 coordinates after int(x * scale) (src/orb.cpp:94-98); levels are concatenated 0..L-1.
// Semantics fixed where the reference is broken or nondeterministic are listed in DESIGN.md (D1-D10).
#ifndef ORB_H
#define ORB_H

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#else
#include "orb_cv_min.hpp"
#endif
#else
#include "orb_cv_min.hpp"
#endif

#include "orb_b200.h"

struct Keypoint { int x, y; };

struct ORBDescriptor {
    uint8_t data[32];
};

static_assert(sizeof(Keypoint) == sizeof(orb_keypoint) && sizeof(ORBDescriptor) == sizeof(orb_descriptor),
              "record layouts must match the C ABI");

namespace orb_b200_detail {

inline void check_image(const cv::Mat& image) {
    if (image.empty() || image.type() != CV_8UC1)   // CV_Assert(image.type() == CV_8UC1), reference src/orb_cpu.cpp:26
        throw std::runtime_error("orb_b200: image must be a non-empty CV_8UC1 cv::Mat");
}

// One C-ABI context, (re)created on demand when a larger image shows up.
class Handle {
public:
    explicit Handle(const orb_params& p) : params_(p) {}
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    ~Handle() { orb_destroy(ctx_); }

    orb_ctx* get(int w, int h) {
        if (!ctx_ || w > params_.max_width || h > params_.max_height) {
            orb_destroy(ctx_);
            ctx_ = nullptr;
            if (w > params_.max_width) params_.max_width = w;
            if (h > params_.max_height) params_.max_height = h;
            if (const char* d = std::getenv("ORB_B200_DEVICE")) params_.device = std::atoi(d);
            int rc = orb_create(&params_, &ctx_);
            if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: orb_create failed: ") + orb_last_error(nullptr));
        }
        return ctx_;
    }
    void check(int rc) const {
        if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error(ctx_));
    }
    const orb_params& params() const { return params_; }

private:
    orb_params params_;
    orb_ctx* ctx_ = nullptr;
};

inline orb_params make_params(int nfeatures, float scaleFactor, int nlevels, int threshold, int n, int nms_window,
                              int patch_size, int policy) {
    orb_params p;
    orb_default_params(&p);
    p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels;
    p.fast_threshold = threshold; p.fast_n = n; p.nms_window = nms_window; p.orient_patch = patch_size;
    p.select_policy = policy;
    return p;
}

}  // namespace orb_b200_detail

class OrientedFAST {
public:
    OrientedFAST(int threshold=20, int n=9, int nms_window=3, int patch_size=31)
        : threshold(threshold), n(n), nms_window(nms_window), patch_size(patch_size),
          h_(new orb_b200_detail::Handle(orb_b200_detail::make_params(ORB_MAX_STAGE_FEATURES, 1.2f, 1, threshold, n, nms_window,
                                                                     patch_size, ORB_SELECT_RASTER_FIRST_N))) {}
    OrientedFAST(const OrientedFAST& o) : OrientedFAST(o.threshold, o.n, o.nms_window, o.patch_size) {}
    OrientedFAST& operator=(const OrientedFAST&) = delete;
    ~OrientedFAST() { delete h_; }

    // FAST-n + SAD score + 3x3 NMS; the first `nfeatures` survivors in raster order
    // (reference src/orb.cpp:22-27 -> Fast(), src/cuda/Fast.cu:211-270; deterministic order as src/orb_cpu.cpp:110)
    std::vector<Keypoint> detect(const cv::Mat& image, int nfeatures) {
        orb_b200_detail::check_image(image);
        this->nfeatures = nfeatures;
        if (nfeatures <= 0) return std::vector<Keypoint>();      // detect(pyr, 2 * quota) with quota 0: the reference returns an empty vector
        std::vector<Keypoint> kps(nfeatures);
        int count = 0;
        h_->check(orb_fast_detect(h_->get(image.cols, image.rows), image.data, image.cols, image.rows, image.step, nfeatures,
                                  reinterpret_cast<orb_keypoint*>(kps.data()), &count));
        kps.resize(count);
        return kps;
    }
    // intensity-centroid angle per keypoint (reference src/orb.cpp:29-33 -> Orientations())
    std::vector<float> compute_orientations(const cv::Mat& image, const std::vector<Keypoint>& keypoints) {
        orb_b200_detail::check_image(image);
        std::vector<float> out(keypoints.size());
        h_->check(orb_orientations(h_->get(image.cols, image.rows), image.data, image.cols, image.rows, image.step,
                                   reinterpret_cast<const orb_keypoint*>(keypoints.data()), (int)keypoints.size(), out.data()));
        return out;
    }

private:
    enum { ORB_MAX_STAGE_FEATURES = 8192 };
    int nfeatures = 0;
    int threshold;
    int n;
    int nms_window;
    int patch_size;
    orb_b200_detail::Handle* h_;
};

class RotatedBRIEF {
public:
    RotatedBRIEF() : h_(new orb_b200_detail::Handle(orb_b200_detail::make_params(8192, 1.2f, 1, 20, 9, 3, 31, ORB_SELECT_RASTER_FIRST_N))) {}
    RotatedBRIEF(const RotatedBRIEF&) : RotatedBRIEF() {}
    RotatedBRIEF& operator=(const RotatedBRIEF&) = delete;
    ~RotatedBRIEF() { delete h_; }

    // 256-bit rotated BRIEF (reference src/orb.cpp:40-44 -> Brief(), src/cuda/Brief.cu:97-137)
    std::vector<ORBDescriptor> compute(const cv::Mat& image, const std::vector<Keypoint>& keypoints, const std::vector<float>& orientations) {
        orb_b200_detail::check_image(image);
        if (keypoints.size() != orientations.size()) throw std::runtime_error("orb_b200: keypoints / orientations size mismatch");
        std::vector<ORBDescriptor> out(keypoints.size());
        h_->check(orb_brief(h_->get(image.cols, image.rows), image.data, image.cols, image.rows, image.step,
                            reinterpret_cast<const orb_keypoint*>(keypoints.data()), orientations.data(), (int)keypoints.size(),
                            reinterpret_cast<orb_descriptor*>(out.data())));
        return out;
    }

private:
    int n_bits = 256;
    int patch_size = 31;
    orb_b200_detail::Handle* h_;
};

class ORB {
public:
    ORB(int nfeatures=500, float scaleFactor=1.2f, int nlevels=8)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels),
          h_(new orb_b200_detail::Handle(orb_b200_detail::make_params(nfeatures, scaleFactor, nlevels, 20, 9, 3, 31, ORB_SELECT_HARRIS_TOP_N))) {}
    // extra knobs keep the reference signature above intact
    ORB(int nfeatures, float scaleFactor, int nlevels, int fast_threshold, int patch_size, int select_policy = ORB_SELECT_HARRIS_TOP_N,
        bool blur_levels = true)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), h_(nullptr) {
        orb_params p = orb_b200_detail::make_params(nfeatures, scaleFactor, nlevels, fast_threshold, 9, 3, patch_size, select_policy);
        p.blur_levels = blur_levels ? 1 : 0;
        h_ = new orb_b200_detail::Handle(p);
    }
    ORB(const ORB&) = delete;
    ORB& operator=(const ORB&) = delete;
    ~ORB() { delete h_; }

    // reference src/orb.cpp:58-109: pyramid -> per level {FAST, Harris, top quota_l, orientation, BRIEF} -> append
    void detectAndCompute(const cv::Mat& image, std::vector<Keypoint>& keypoints, std::vector<float>& orientations, std::vector<ORBDescriptor>& descriptors) {
        orb_b200_detail::check_image(image);
        orb_ctx* c = h_->get(image.cols, image.rows);
        int cap = 0;
        for (int l = 0; l < nlevels; l++) cap += orb_level_quota(c, l);
        if (cap < 1) cap = 1;
        std::vector<Keypoint> k(cap);
        std::vector<float> a(cap);
        std::vector<ORBDescriptor> d(cap);
        int count = 0;
        h_->check(orb_detect_and_compute(c, image.data, image.cols, image.rows, image.step, cap, reinterpret_cast<orb_keypoint*>(k.data()),
                                         a.data(), reinterpret_cast<orb_descriptor*>(d.data()), &count, nullptr));
        keypoints.insert(keypoints.end(), k.begin(), k.begin() + count);
        orientations.insert(orientations.end(), a.begin(), a.begin() + count);
        descriptors.insert(descriptors.end(), d.begin(), d.begin() + count);
    }

private:
    int nfeatures;
    float scaleFactor;
    int nlevels;
    orb_b200_detail::Handle* h_;
};

#endif // ORB_H
