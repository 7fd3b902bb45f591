// Drives the cv::Feature2D adapter the way the reference's VO loops drive their detector
// (src/feature_tracking.cpp:59,201-202; src/feature_matching.cpp:56,164), against tests/cpp/mock_opencv.
//   feature2d_test <raw> <w> <h> <outprefix>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "orb_feature2d.hpp"

int main(int argc, char** argv) {
  if (argc < 5) return 1;
  int w = atoi(argv[2]), h = atoi(argv[3]);
  std::string out = argv[4];
  std::vector<unsigned char> pix((size_t)w * h);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(pix.data(), 1, pix.size(), f) != pix.size()) return 2;
  fclose(f);
  cv::Mat img(h, w, CV_8UC1, pix.data(), (size_t)w);
  try {
    cv::Ptr<cv::Feature2D> orb = orb_b200::ORBFeature2D::create(3000);   // reference: cv::ORB::create(3000)
    std::vector<cv::KeyPoint> kp1, kp2;
    cv::Mat des1;
    orb->detect(img, kp1);                                               // src/feature_tracking.cpp:59
    orb->detectAndCompute(img, cv::noArray(), kp2, des1);                // src/feature_tracking.cpp:201
    if (kp1.size() != kp2.size() || des1.rows != (int)kp2.size() || des1.cols != 32 || des1.type() != CV_8U) return 3;
    if (orb->descriptorSize() != 32 || orb->descriptorType() != CV_8U || orb->defaultNorm() != cv::NORM_HAMMING) return 4;
    FILE* o = fopen((out + ".kp").c_str(), "wb");
    for (auto& k : kp2) { float rec[6] = {k.pt.x, k.pt.y, k.size, k.angle, k.response, (float)k.octave}; fwrite(rec, 4, 6, o); }
    fclose(o);
    o = fopen((out + ".desc").c_str(), "wb");
    for (int i = 0; i < des1.rows; i++) fwrite(des1.ptr<unsigned char>(i), 1, 32, o);
    fclose(o);
    printf("FEATURE2D_OK %zu\n", kp2.size());
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 5;
  }
  return 0;
}
