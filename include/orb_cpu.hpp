// orb_cpu.hpp -- name-compatible twins of the reference's include/orb_cpu.hpp (OrientedFASTCPU :4-16,
// RotatedBRIEFCPU :18-26, ORBCPU :28-43) so that code written against the reference's CPU classes links too.
// They are NOT a CPU implementation: they run the same B200 kernels, configured the way the reference's CPU path
// behaves -- in particular ORBCPU::detectAndCompute is single-level, FAST threshold 50, first 3000 survivors in
// raster order, orientation patch 9, whatever its constructor arguments are (reference src/orb_cpu.cpp:260-276:
// the arguments are stored and never used) -- and it ASSIGNS its outputs (:272-275).
#ifndef ORB_CPU_H
#define ORB_CPU_H
#include "orb.hpp"

class OrientedFASTCPU {
public:
    OrientedFASTCPU(int nfeatures=3000, int threshold=50, int n=9, int nms_window=3, int patch_size=9)
        : nfeatures(nfeatures), fast(threshold, n, nms_window, patch_size) {}
    std::vector<Keypoint> detect(const cv::Mat& image) { return fast.detect(image, nfeatures); }
    std::vector<float> compute_orientations(const cv::Mat& image, const std::vector<Keypoint>& keypoints) {
        return fast.compute_orientations(image, keypoints);
    }
private:
    int nfeatures;
    OrientedFAST fast;
};

class RotatedBRIEFCPU {
public:
    RotatedBRIEFCPU() {}
    // 5x5 box sum around (x, y) from an integral image (reference include/orb_cpu.hpp:21, src/orb_cpu.cpp:190-201): four taps
    // of the caller's CV_32S matrix, addressed through its own row step like cv::Mat::at<int> (`width` is unused there too).
    // Host arithmetic on caller data -- the descriptors themselves come from compute(), i.e. from the GPU.
    int sum5x5(const cv::Mat& integral, int x, int y, int /*width*/) {
        const int x0 = x - 2, y0 = y - 2, x1 = x + 3, y1 = y + 3;
        const unsigned char* base = integral.data;
        auto at = [&](int r, int c) { return reinterpret_cast<const int*>(base + (size_t)r * integral.step)[c]; };
        return at(y1, x1) + at(y0, x0) - at(y0, x1) - at(y1, x0);
    }
    std::vector<ORBDescriptor> compute(const cv::Mat& image, const std::vector<Keypoint>& keypoints, const std::vector<float>& orientations) {
        return brief.compute(image, keypoints, orientations);
    }
private:
    RotatedBRIEF brief;
};

class ORBCPU {
public:
    ORBCPU(int nfeatures=500, float scaleFactor=1.2f, int nlevels=8)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), orb(3000, scaleFactor, 1, 50, 9, ORB_SELECT_RASTER_FIRST_N) {}
    void detectAndCompute(const cv::Mat& image, std::vector<Keypoint>& keypoints, std::vector<float>& orientations, std::vector<ORBDescriptor>& descriptors) {
        keypoints.clear(); orientations.clear(); descriptors.clear();
        orb.detectAndCompute(image, keypoints, orientations, descriptors);
    }
private:
    int nfeatures;
    float scaleFactor;
    int nlevels;
    ORB orb;
};

#endif // ORB_CPU_H
