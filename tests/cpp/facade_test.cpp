// Reads a raw 8-bit image, runs the reference-style C++ facade (include/orb.hpp, include/orb_cpu.hpp) and dumps the
// results as raw arrays so the pytest harness can compare them with the oracle.  Written the way the reference's
// only driver uses the classes (src/compare.cpp:39-65).
//   facade_test <raw> <w> <h> <outprefix>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "orb.hpp"
#include "orb_cpu.hpp"
#include "orb_pattern.hpp"
#include "orb_stages.hpp"

template <class T> static void dump(const std::string& path, const std::vector<T>& v) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 5) return 1;
    int w = atoi(argv[2]), h = atoi(argv[3]);
    std::string out = argv[4];
    std::vector<unsigned char> pix((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(pix.data(), 1, pix.size(), f) != pix.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    cv::Mat image(h, w, CV_8UC1, pix.data(), (size_t)w);
    try {
        ORB orb_gpu;                                   // src/compare.cpp:40
        std::vector<Keypoint> keypoints_gpu; std::vector<float> orientations_gpu; std::vector<ORBDescriptor> descriptors_gpu;
        orb_gpu.detectAndCompute(image, keypoints_gpu, orientations_gpu, descriptors_gpu);   // :65
        size_t first = keypoints_gpu.size();
        orb_gpu.detectAndCompute(image, keypoints_gpu, orientations_gpu, descriptors_gpu);   // appends (src/orb.cpp:100-102)
        if (keypoints_gpu.size() != 2 * first) { fprintf(stderr, "append semantics broken\n"); return 3; }
        keypoints_gpu.resize(first); orientations_gpu.resize(first); descriptors_gpu.resize(first);
        dump(out + ".orb.kps", keypoints_gpu); dump(out + ".orb.ang", orientations_gpu); dump(out + ".orb.desc", descriptors_gpu);

        ORBCPU orb_cpu;                                // src/compare.cpp:39 (commented there)
        std::vector<Keypoint> kc(7); std::vector<float> ac(7); std::vector<ORBDescriptor> dc(7);
        orb_cpu.detectAndCompute(image, kc, ac, dc);   // assigns (src/orb_cpu.cpp:272-275)
        dump(out + ".cpu.kps", kc); dump(out + ".cpu.ang", ac); dump(out + ".cpu.desc", dc);

        OrientedFAST fast;                             // include/orb.hpp:12 defaults
        RotatedBRIEF brief;
        std::vector<Keypoint> ks = fast.detect(image, 700);
        std::vector<float> as = fast.compute_orientations(image, ks);
        std::vector<ORBDescriptor> ds = brief.compute(image, ks, as);
        dump(out + ".stage.kps", ks); dump(out + ".stage.ang", as); dump(out + ".stage.desc", ds);
        // the free functions of the reference's .cuh seam (src/orb.cpp:24,31,42,65)
        std::vector<Keypoint> kf; std::vector<float> af, hf; std::vector<ORBDescriptor> df;
        int kp_count = Fast(image, kf, 20, 9, 3, 700);
        Orientations(image, kf, af, 31);
        Brief(image, kf, af, df, 256, 31);
        HarrisScore(image, kf, hf, 7, 0.04f);
        // the reference's call site passes a double literal (src/orb.cpp:65): resolves to the floating-point overload, k = 0.04
        std::vector<float> hd; HarrisScore(image, kf, hd, 7, 0.04);
        if (hd != hf) { fprintf(stderr, "HarrisScore(double) differs from HarrisScore(float)\n"); return 6; }
        // the reference's exact signature, through a pointer of that type (include/HarrisScore.cuh:5): integer k
        void (*harris_ref)(const cv::Mat&, std::vector<Keypoint>&, std::vector<float>&, int, int) = &HarrisScore;
        std::vector<float> h0; harris_ref(image, kf, h0, 7, 0);
        dump(out + ".free.harris_k0", h0);
        // NMS() over a caller's score map (include/NMS.cuh:5): a synthetic float map with plateaus
        std::vector<float> smap((size_t)w * h);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) smap[(size_t)y * w + x] = (float)(((x * 7 + y * 13) % 31) * ((x ^ y) & 1));
        cv::Mat score(h, w, CV_32FC1, smap.data());
        std::vector<Keypoint> kn; NMS(score, kn, 3, 5000, 10.0f);
        dump(out + ".free.nms", kn);
        // RotatedBRIEFCPU::sum5x5 (include/orb_cpu.hpp:21): four taps of a caller's integral image
        std::vector<int> integ((size_t)(w + 1) * (h + 1), 0);
        for (int y = 0; y < h; y++) { int run = 0; for (int x = 0; x < w; x++) { run += pix[(size_t)y * w + x]; integ[(size_t)(y + 1) * (w + 1) + x + 1] = integ[(size_t)y * (w + 1) + x + 1] + run; } }
        cv::Mat imat(h + 1, w + 1, CV_32SC1, integ.data());
        RotatedBRIEFCPU bcpu;
        long box = 0; for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) box += pix[(size_t)(100 + dy) * w + 200 + dx];
        if (bcpu.sum5x5(imat, 200, 100, w + 1) != box) { fprintf(stderr, "sum5x5\n"); return 7; }
        // the reference's stand-alone filters (include/Convolution.cuh, GaussianBlur.cuh, GaussianBlur.hpp, Sobel.hpp)
        {
            cv::Mat g5, g1d, g7, sbx, sby, cv3;
            GaussianBlur(image, g5); GaussianBlur1D(image, g1d); GaussianBlurCUDA(image, g7, 7); SobelCUDA(image, sbx, 0); SobelCUDA(image, sby, 1);
            float box3[9] = {0.111f, 0.111f, 0.111f, 0.111f, 0.112f, 0.111f, 0.111f, 0.111f, 0.111f};
            if (conv2d(image, cv3, box3, 3) != 0 || cv3.rows != h - 2 || cv3.cols != w - 2) { fprintf(stderr, "conv2d shape\n"); return 9; }
            auto dump_mat = [&](const std::string& tag, const cv::Mat& m) {
                std::vector<unsigned char> v((size_t)m.rows * m.cols);
                for (int y = 0; y < m.rows; y++) for (int x = 0; x < m.cols; x++) v[(size_t)y * m.cols + x] = m.ptr<unsigned char>(y)[x];
                dump(out + ".filt." + tag, v);
            };
            dump_mat("g5", g5); dump_mat("g1d", g1d); dump_mat("g7", g7); dump_mat("sx", sbx); dump_mat("sy", sby); dump_mat("c3", cv3);
        }
        // detect(image, 0): the reference returns an empty vector (quota 0 levels)
        if (!fast.detect(image, 0).empty()) { fprintf(stderr, "detect(0)\n"); return 8; }
        if (kp_count != (int)kf.size() || kf.size() != ks.size()) { fprintf(stderr, "Fast() count\n"); return 5; }
        dump(out + ".free.kps", kf); dump(out + ".free.ang", af); dump(out + ".free.desc", df); dump(out + ".free.harris", hf);
        long s = 0; for (int i = 0; i < 1024; i++) s += bit_pattern_31_[i];
        printf("FACADE_OK %zu %zu %zu pattern_sum %ld\n", first, kc.size(), ks.size(), s);
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 4;
    }
    return 0;
}
