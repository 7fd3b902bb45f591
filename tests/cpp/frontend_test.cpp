// Drives include/orb_vo_frontend.hpp the way the reference's feature_tracking loop drives OpenCV
// (src/feature_tracking.cpp:56,174-180,196), against tests/cpp/mock_opencv.
//   frontend_test <png0> <png1> <points.f32> <outprefix>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "orb_vo_frontend.hpp"

int main(int argc, char** argv) {
  if (argc < 5) return 1;
  try {
    cv::Mat img_1 = orb_b200::imread(argv[1], cv::IMREAD_GRAYSCALE);
    cv::Mat img_2 = orb_b200::imread(argv[2], cv::IMREAD_GRAYSCALE);
    if (img_1.empty() || img_2.empty()) return 2;
    if (!orb_b200::imread("/nonexistent/file.png").empty()) return 3;
    std::vector<cv::Point2f> points1, points2;
    FILE* f = fopen(argv[3], "rb");
    float xy[2];
    while (f && fread(xy, 4, 2, f) == 2) points1.emplace_back(xy[0], xy[1]);
    if (f) fclose(f);
    std::vector<unsigned char> status;
    std::vector<float> err;
    cv::Size winSize = cv::Size(21, 21);                                                       // src/feature_tracking.cpp:176
    cv::TermCriteria termcrit = cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01);
    orb_b200::calcOpticalFlowPyrLK(img_1, img_2, points1, points2, status, err, winSize, 3, termcrit, 0, 0.001);
    std::string out = argv[4];
    FILE* o = fopen((out + ".img").c_str(), "wb");
    for (int y = 0; y < img_1.rows; y++) fwrite(img_1.ptr<unsigned char>(y), 1, img_1.cols, o);
    fclose(o);
    o = fopen((out + ".pts").c_str(), "wb"); fwrite(points2.data(), 8, points2.size(), o); fclose(o);
    o = fopen((out + ".st").c_str(), "wb"); fwrite(status.data(), 1, status.size(), o); fclose(o);
    o = fopen((out + ".err").c_str(), "wb"); fwrite(err.data(), 4, err.size(), o); fclose(o);
    printf("FRONTEND_OK %d x %d, %zu points\n", img_1.cols, img_1.rows, points2.size());
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 5;
  }
  return 0;
}
