// ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" bridge over the reference's OWN CPU classes (ORBCPU, OrientedFASTCPU, RotatedBRIEFCPU:
// include/orb_cpu.hpp:4-43, src/orb_cpu.cpp), compiled unmodified from /root/reference against
// oracle/cv_shim.  Output: oracle/_ref/liborbcpu_ref.so (git-ignored; built by oracle/Makefile).
#include <opencv2/opencv.hpp>
#include <sstream>
#include <vector>
#include "orb.hpp"       // reference include/orb.hpp (Keypoint, ORBDescriptor)
#include "orb_cpu.hpp"   // reference include/orb_cpu.hpp

namespace {
struct Quiet {   // the reference prints its parameters / keypoint counts on every call
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
cv::Mat view(const uint8_t* img, int w, int h, size_t pitch) { return cv::Mat(h, w, CV_8UC1, (void*)img, pitch); }
}  // namespace

extern "C" {

// ORBCPU orb; orb.detectAndCompute(...)  -- exactly what src/compare.cpp:39,48 (commented) would run.
int ref_orbcpu_detect_and_compute(const uint8_t* img, int w, int h, size_t pitch, int cap, Keypoint* kps, float* angles,
                                  ORBDescriptor* desc) {
  Quiet q;
  ORBCPU orb;
  std::vector<Keypoint> k; std::vector<float> a; std::vector<ORBDescriptor> d;
  orb.detectAndCompute(view(img, w, h, pitch), k, a, d);
  int n = (int)std::min<size_t>(k.size(), (size_t)cap);
  for (int i = 0; i < n; i++) { kps[i] = k[i]; angles[i] = a[i]; desc[i] = d[i]; }
  return (int)k.size();
}

int ref_fast_detect(const uint8_t* img, int w, int h, size_t pitch, int nfeatures, int thr, int n, int nms, int cap,
                    Keypoint* kps) {
  OrientedFASTCPU fast(nfeatures, thr, n, nms, 9);
  std::vector<Keypoint> k = fast.detect(view(img, w, h, pitch));
  int m = (int)std::min<size_t>(k.size(), (size_t)cap);
  for (int i = 0; i < m; i++) kps[i] = k[i];
  return (int)k.size();
}

void ref_orientations(const uint8_t* img, int w, int h, size_t pitch, const Keypoint* kps, int n, int patch, float* out) {
  OrientedFASTCPU fast(3000, 50, 9, 3, patch);
  std::vector<Keypoint> k(kps, kps + n);
  std::vector<float> a = fast.compute_orientations(view(img, w, h, pitch), k);
  for (int i = 0; i < n; i++) out[i] = a[i];
}

void ref_brief(const uint8_t* img, int w, int h, size_t pitch, const Keypoint* kps, const float* angles, int n,
               ORBDescriptor* out) {
  RotatedBRIEFCPU brief;
  std::vector<Keypoint> k(kps, kps + n);
  std::vector<float> a(angles, angles + n);
  std::vector<ORBDescriptor> d = brief.compute(view(img, w, h, pitch), k, a);
  for (int i = 0; i < n; i++) out[i] = d[i];
}

extern int bit_pattern_31_[256 * 4];
const int* ref_pattern(void) { return bit_pattern_31_; }

}  // extern "C"
