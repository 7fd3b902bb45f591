// orb_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see orb_oracle.h for who may load it).
//
// Dependency-free C++17 restatement of the reference's CPU ORB (src/orb_cpu.cpp) plus the level
// loop of its GPU facade (src/orb.cpp:58-109) under decisions D1-D10 of SURVEY.md 8(c).
// Build with -O2 -ffp-contract=off, no -ffast-math, no -march=native: the reference's float
// expressions (c*x - s*y, Harris) must not be contracted into FMAs.
// All file:line citations are relative to the reference tree.
#include "orb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/orb_brief_pattern.h"

namespace {

// FAST Bresenham ring, radius 3, in the reference's order (src/orb_cpu.cpp:8-13): (dx, dy)
const int kRing[16][2] = {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},   {2, 2},   {1, 3},
                          {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};

inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

struct Img {
  const uint8_t* d;
  int w, h;
  size_t p;
  int at(int y, int x) const { return d[(size_t)y * p + x]; }
};

// cv::resize(..., INTER_LINEAR) tap table for one axis: OpenCV's float recipe, 11-bit coefficients.
void resize_table(int src, int dst, std::vector<int>& ofs, std::vector<int16_t>& a0, std::vector<int16_t>& a1) {
  ofs.resize(dst); a0.resize(dst); a1.resize(dst);
  double inv_scale = (double)dst / src;
  double scale = 1.0 / inv_scale;
  for (int d = 0; d < dst; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
    ofs[d] = s;
    a0[d] = (int16_t)std::lrintf((1.f - f) * 2048.f);   // saturate_cast<short>(float) = round-half-even
    a1[d] = (int16_t)std::lrintf(f * 2048.f);
  }
}

}  // namespace

extern "C" {

const int8_t* orc_pattern(void) { return ORB_BRIEF_PATTERN_31; }
long orc_lround_f(float v) { return std::lround(v); }
float orc_atan2f(float y, float x) { return std::atan2(y, x); }
float orc_cosf(float a) { return std::cos(a); }
float orc_sinf(float a) { return std::sin(a); }

float orc_level_scale(float f, int level) {
  // float scale = pow(scaleFactor, i);   src/orb_cpu.cpp:284 (double pow, narrowed to float)
  return (float)std::pow((double)f, (double)level);
}

void orc_level_size(int W, int H, float f, int level, int* w, int* h) {
  if (level == 0) { *w = W; *h = H; return; }
  float scale = orc_level_scale(f, level);
  // cv::Size newSize(round(W / scale), round(H / scale));   src/orb_cpu.cpp:285
  *w = (int)std::round(W / scale);
  *h = (int)std::round(H / scale);
}

int orc_level_quota(int nfeatures, float f, int nlevels, int level) {
  // int nfeatures_l = nfeatures * ((1 - 1/scaleFactor) / (1 - std::pow(1/scaleFactor, nlevels)))
  //                   * std::pow(1/scaleFactor, l);                                  src/orb.cpp:62
  float inv = 1 / f;
  double a = (1 - inv) / (1 - std::pow((double)inv, (double)nlevels));
  double q = nfeatures * a * std::pow((double)inv, (double)level);
  return (int)q;
}

void orc_resize_table(int src, int dst, int32_t* ofs, int16_t* a0, int16_t* a1) {
  std::vector<int> o; std::vector<int16_t> x0, x1;
  resize_table(src, dst, o, x0, x1);
  for (int i = 0; i < dst; i++) { ofs[i] = o[i]; a0[i] = x0[i]; a1[i] = x1[i]; }
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sp, uint8_t* dst, int dw, int dh, size_t dp) {
  std::vector<int> xo, yo; std::vector<int16_t> xa0, xa1, yb0, yb1;
  resize_table(sw, dw, xo, xa0, xa1);
  resize_table(sh, dh, yo, yb0, yb1);
  std::vector<int> r0(dw), r1(dw);
  for (int y = 0; y < dh; y++) {
    int sy0 = yo[y], sy1 = std::min(sy0 + 1, sh - 1);
    const uint8_t* S0 = src + (size_t)sy0 * sp;
    const uint8_t* S1 = src + (size_t)sy1 * sp;
    for (int x = 0; x < dw; x++) {
      int sx0 = xo[x], sx1 = std::min(sx0 + 1, sw - 1);
      r0[x] = S0[sx0] * xa0[x] + S0[sx1] * xa1[x];
      r1[x] = S1[sx0] * xa0[x] + S1[sx1] * xa1[x];
    }
    int b0 = yb0[y], b1 = yb1[y];
    for (int x = 0; x < dw; x++) {
      int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
      dst[(size_t)y * dp + x] = (uint8_t)std::min(255, std::max(0, v));
    }
  }
}

void orc_gauss5x5_u8(const uint8_t* src, int w, int h, size_t sp, uint8_t* dst, size_t dp) {
  static const int k[5] = {1, 4, 6, 4, 1};
  std::vector<int> hs((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int s = 0;
      for (int i = -2; i <= 2; i++) s += k[i + 2] * src[(size_t)y * sp + reflect101(x + i, w)];
      hs[(size_t)y * w + x] = s;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int s = 0;
      for (int i = -2; i <= 2; i++) s += k[i + 2] * hs[(size_t)reflect101(y + i, h) * w + x];
      dst[(size_t)y * dp + x] = (uint8_t)((s + 128) >> 8);
    }
}

void orc_integral_flat(const uint8_t* src, int w, int h, size_t sp, int32_t* out) {
  const int W1 = w + 1;
  std::memset(out, 0, sizeof(int32_t) * (size_t)W1 * (h + 4));
  for (int y = 0; y < h; y++) {
    int row = 0;
    for (int x = 0; x < w; x++) {
      row += src[(size_t)y * sp + x];
      out[(size_t)(y + 1) * W1 + x + 1] = out[(size_t)y * W1 + x + 1] + row;
    }
  }
}

void orc_build_level(const uint8_t* img, int W, int H, size_t pitch, const orc_params* p, int level, uint8_t* dst) {
  int w, h;
  orc_level_size(W, H, p->scale_factor, level, &w, &h);
  if (level == 0) {   // pyramid[0] = image;   src/orb_cpu.cpp:279
    for (int y = 0; y < H; y++) std::memcpy(dst + (size_t)y * W, img + (size_t)y * pitch, W);
    return;
  }
  // every level is resampled from level 0 (src/orb_cpu.cpp:288, src/orb.cpp:119), then blurred (:289)
  if (p->blur_levels) {
    std::vector<uint8_t> tmp((size_t)w * h);
    orc_resize_linear_u8(img, W, H, pitch, tmp.data(), w, h, w);
    orc_gauss5x5_u8(tmp.data(), w, h, w, dst, w);
  } else {
    orc_resize_linear_u8(img, W, H, pitch, dst, w, h, w);
  }
}

// ---- FAST: src/orb_cpu.cpp:34-103 ----------------------------------------------------------
void orc_fast_scores(const uint8_t* data, int w, int h, size_t pitch, int thr, int n, float* scores) {
  Img im{data, w, h, pitch};
  std::fill(scores, scores + (size_t)w * h, 0.0f);
  for (int y = 3; y < h - 3; y++) {
    for (int x = 3; x < w - 3; x++) {
      int Ip = im.at(y, x);
      static const int check_idx[4] = {0, 4, 8, 12};          // :40
      int brighter = 0, darker = 0;
      for (int k = 0; k < 4; k++) {
        int cp = im.at(y + kRing[check_idx[k]][1], x + kRing[check_idx[k]][0]);
        if (cp >= Ip + thr) brighter++;
        else if (cp <= Ip - thr) darker++;
      }
      if (std::max(brighter, darker) < 3) continue;            // :57
      int cv[32];
      for (int i = 0; i < 16; i++) {
        int v = im.at(y + kRing[i][1], x + kRing[i][0]);
        cv[i] = v; cv[i + 16] = v;
      }
      for (int i = 0; i < 16; i++) {                           // :73
        bool all_b = true, all_d = true;
        for (int j = 0; j < n; j++) {
          int v = cv[i + j];
          if (v < Ip + thr) all_b = false;
          if (v > Ip - thr) all_d = false;
        }
        if (all_b || all_d) {
          float score = 0.0f;                                  // :90-96
          for (int r = 0; r < 16; r++) score += std::abs(Ip - cv[r]);
          scores[(size_t)y * w + x] = score;
          break;
        }
      }
    }
  }
}

// ---- NMS + raster cap: src/orb_cpu.cpp:105-134 ---------------------------------------------
int orc_nms(const float* scores, int w, int h, int nms_window, int cap, orc_keypoint* kps) {
  int r = nms_window / 2;
  int n = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      float s = scores[(size_t)y * w + x];
      if (s <= 0.0f || n >= cap) continue;                      // :110
      if (r != 0) {
        double mx = -1e300;                                     // cv::minMaxLoc on the ROI, :121-124
        for (int yy = std::max(0, y - r); yy <= std::min(h - 1, y + r); yy++)
          for (int xx = std::max(0, x - r); xx <= std::min(w - 1, x + r); xx++)
            mx = std::max(mx, (double)scores[(size_t)yy * w + xx]);
        if (std::abs(s - mx) < 1e-6f) kps[n++] = {x, y};        // :126 (ties keep both)
      } else {
        kps[n++] = {x, y};
      }
    }
  return n;
}

// ---- Harris (decision D5): intent of src/cuda/HarrisScore.cu:23-89 + src/Sobel.cpp + src/GaussianBlur.cpp
void orc_harris_weights(float* kernel) {
  // createGaussianKernel(7) with the sigma heuristic, src/GaussianBlur.cpp:7-37, float arithmetic
  const int kernelSize = 7;
  float sigma = 0.3f * ((kernelSize - 1) * 0.5f) + 0.8f;
  int half = kernelSize / 2;
  float sum = 0.0f;
  for (int y = -half; y <= half; ++y)
    for (int x = -half; x <= half; ++x) {
      float value = std::exp(-(x * x + y * y) / (2 * sigma * sigma));
      kernel[(y + half) * kernelSize + (x + half)] = value;
      sum += value;
    }
  for (int i = 0; i < kernelSize * kernelSize; ++i) kernel[i] /= sum;
}

void orc_harris(const uint8_t* data, int w, int h, size_t pitch, const orc_keypoint* kps, int n, float k, float* out) {
  Img im{data, w, h, pitch};
  float wt[49];
  orc_harris_weights(wt);
  for (int i = 0; i < n; i++) {
    int x = kps[i].x, y = kps[i].y;
    float A = 0.f, B = 0.f, C = 0.f;   // G7(Ix^2), G7(IxIy), G7(Iy^2) at (x,y); conv2d order src/cuda/Convolution.cu:45-49
    for (int dy = -3; dy <= 3; dy++)
      for (int dx = -3; dx <= 3; dx++) {
        int yy = y + dy, xx = x + dx;   // window centre; Sobel taps use BORDER_REFLECT_101 (src/Sobel.cpp:29)
        int ym = reflect101(yy - 1, h), y0 = reflect101(yy, h), yp = reflect101(yy + 1, h);
        int xm = reflect101(xx - 1, w), x0 = reflect101(xx, w), xp = reflect101(xx + 1, w);
        int ix = (im.at(ym, xp) + 2 * im.at(y0, xp) + im.at(yp, xp)) - (im.at(ym, xm) + 2 * im.at(y0, xm) + im.at(yp, xm));
        int iy = (im.at(yp, xm) + 2 * im.at(yp, x0) + im.at(yp, xp)) - (im.at(ym, xm) + 2 * im.at(ym, x0) + im.at(ym, xp));
        float g = wt[(dy + 3) * 7 + (dx + 3)];
        A += (float)(ix * ix) * g;
        B += (float)(ix * iy) * g;
        C += (float)(iy * iy) * g;
      }
    float det = A * C - B * B;          // src/cuda/HarrisScore.cu:35
    float trace = A + C;                // :36
    out[i] = det - k * trace * trace;   // :38 with k = 0.04f (the reference's `int k` truncates it to 0)
  }
}

// ---- the reference's filter wrappers ----------------------------------------------------------------------------
// createGaussianKernel(k), src/GaussianBlur.cpp:7-37 (float arithmetic, sigma heuristic)
void orc_gaussian_kernel(int kernelSize, float* kernel) {
  float sigma = 0.3f * ((kernelSize - 1) * 0.5f) + 0.8f;
  int halfSize = kernelSize / 2;
  float sum = 0.0f;
  for (int y = -halfSize; y <= halfSize; ++y)
    for (int x = -halfSize; x <= halfSize; ++x) {
      float value = std::exp(-(x * x + y * y) / (2 * sigma * sigma));
      kernel[(y + halfSize) * kernelSize + (x + halfSize)] = value;
      sum += value;
    }
  for (int i = 0; i < kernelSize * kernelSize; ++i) kernel[i] /= sum;
}

// cv::Mat::convertTo(CV_8U) of a float: saturate_cast<uchar>(cvRound(v)), round half to even
static inline uint8_t to_u8(float v) {
  long r = std::lrint(v);
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

// conv2d (src/cuda/Convolution.cu:20-103): valid-mode K x K correlation of the u8 image promoted to float, accumulated row
// by row as `sum += tile * kernel` -- an FMA in the reference's default nvcc build --, result convertTo(CV_8U).
// reflect != 0: the image is first extended by K/2 with BORDER_REFLECT_101 (GaussianBlurCUDA src/GaussianBlur.cpp:39-49,
// SobelCUDA src/Sobel.cpp:18-31, GaussianBlur src/cuda/GaussianBlur.cu:73-77), so the output has the input's size.
// divisor != 0: the sum is divided by it before the conversion (d_GaussianBlur: sum / 273.0f, src/cuda/GaussianBlur.cu:67).
void orc_conv2d_u8(const uint8_t* data, int w, int h, size_t pitch, const float* kernel, int K, int reflect, float divisor, uint8_t* out) {
  Img im{data, w, h, pitch};
  const int r = K / 2, ow = reflect ? w : w - K + 1, oh = reflect ? h : h - K + 1;
  for (int y = 0; y < oh; y++)
    for (int x = 0; x < ow; x++) {
      float sum = 0;
      for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) {
          int yy = reflect ? reflect101(y + i - r, h) : y + i, xx = reflect ? reflect101(x + j - r, w) : x + j;
          sum = std::fmaf((float)im.at(yy, xx), kernel[i * K + j], sum);
        }
      if (divisor != 0.0f) sum = sum / divisor;
      out[(size_t)y * ow + x] = to_u8(sum);
    }
}

// GaussianBlur1D (src/cuda/GaussianBlur1D.cu:34-166): [1 4 6 4 1] / 16 along x, then along y, float, reflect-101 (the
// intent of its halo code), convertTo(CV_8U)
void orc_gaussian_blur_1d(const uint8_t* data, int w, int h, size_t pitch, uint8_t* out) {
  Img im{data, w, h, pitch};
  const float k[5] = {1, 4, 6, 4, 1};
  std::vector<float> tmp((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float sum = 0.0f;
      for (int t = 0; t < 5; t++) sum += k[t] * (float)im.at(y, reflect101(x - 2 + t, w));
      tmp[(size_t)y * w + x] = sum / 16.0f;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float sum = 0.0f;
      for (int t = 0; t < 5; t++) sum += k[t] * tmp[(size_t)reflect101(y - 2 + t, h) * w + x];
      out[(size_t)y * w + x] = to_u8(sum / 16.0f);
    }
}

// ---- orientation: src/orb_cpu.cpp:139-183 --------------------------------------------------
void orc_orientations(const uint8_t* data, int w, int h, size_t pitch, const orc_keypoint* kps, int n, int patch, float* out) {
  Img im{data, w, h, pitch};
  int pr = patch / 2;
  for (int i = 0; i < n; i++) {
    int x = kps[i].x, y = kps[i].y;
    if (x - pr < 0 || x + pr >= w || y - pr < 0 || y + pr >= h) { out[i] = 0.0f; continue; }   // :152-156
    float m10 = 0.0f, m01 = 0.0f;
    for (int r = -pr; r <= pr; ++r)
      for (int c = -pr; c <= pr; ++c) {
        float intensity = (float)im.at(y + r, x + c);
        m10 += c * intensity;
        m01 += r * intensity;
      }
    out[i] = std::atan2(m01, m10);   // :178
  }
}

// ---- rotated BRIEF: src/orb_cpu.cpp:190-258, integral addressed flat + zero tail (D7) --------
static inline int sum5x5_flat(const int32_t* ii, int W1, int x, int y) {
  int x0 = x - 2, y0 = y - 2, x1 = x + 3, y1 = y + 3;    // :192-195
  return ii[(size_t)y1 * W1 + x1] + ii[(size_t)y0 * W1 + x0] - ii[(size_t)y0 * W1 + x1] - ii[(size_t)y1 * W1 + x0];
}

static void brief_core(const int32_t* ii, int w, int h, const orc_keypoint* kps, const float* angles, int n,
                       orc_descriptor* out, uint8_t* flags) {
  const int width = w + 1, height = h + 1;   // integral.cols / integral.rows, :210-211
  for (int idx = 0; idx < n; idx++) {
    float angle = angles[idx];
    float c = std::cos(angle), s = std::sin(angle);   // :217-218
    orc_descriptor desc{};
    uint8_t fl = 0;
    for (int i = 0; i < 256; i++) {
      int x1 = ORB_BRIEF_PATTERN_31[i * 4], y1 = ORB_BRIEF_PATTERN_31[i * 4 + 1];
      int x2 = ORB_BRIEF_PATTERN_31[i * 4 + 2], y2 = ORB_BRIEF_PATTERN_31[i * 4 + 3];
      int dx1 = (int)std::lround(c * x1 - s * y1);    // :228-232
      int dy1 = (int)std::lround(s * x1 + c * y1);
      int dx2 = (int)std::lround(c * x2 - s * y2);
      int dy2 = (int)std::lround(s * x2 + c * y2);
      int cx1 = kps[idx].x + dx1, cy1 = kps[idx].y + dy1;
      int cx2 = kps[idx].x + dx2, cy2 = kps[idx].y + dy2;
      const int r = 5 / 2;                              // :240-245
      if (cx1 < r || cy1 < r || cx1 > width - r || cy1 > height - r ||
          cx2 < r || cy2 < r || cx2 > width - r || cy2 > height - r) { fl |= 1; continue; }
      if (cx1 + 3 > w || cy1 + 3 > h || cx2 + 3 > w || cy2 + 3 > h) fl |= 2;
      if (ii) {
        int s1 = sum5x5_flat(ii, width, cx1, cy1);
        int s2 = sum5x5_flat(ii, width, cx2, cy2);
        if (s1 < s2) desc.data[i >> 3] |= (uint8_t)(1u << (i & 7));   // :250-252
      }
    }
    if (out) out[idx] = desc;
    if (flags) flags[idx] = fl;
  }
}

void orc_brief(const uint8_t* data, int w, int h, size_t pitch, const orc_keypoint* kps, const float* angles, int n,
               orc_descriptor* out) {
  std::vector<int32_t> ii((size_t)(w + 1) * (h + 4));
  orc_integral_flat(data, w, h, pitch, ii.data());
  brief_core(ii.data(), w, h, kps, angles, n, out, nullptr);
}

void orc_brief_flags(int w, int h, const orc_keypoint* kps, const float* angles, int n, uint8_t* flags) {
  brief_core(nullptr, w, h, kps, angles, n, nullptr, flags);
}

// ---- whole path (normative pseudo-code of SURVEY.md 8(c)) -----------------------------------
int orc_detect_and_compute(const uint8_t* img, int W, int H, size_t pitch, const orc_params* p, int cap,
                           orc_keypoint* kps, float* angles, orc_descriptor* desc, int* n_per_level,
                           orc_keypoint* level_xy, int32_t* level_id, float* response) {
  int total = 0;
  std::vector<uint8_t> lvl;
  std::vector<float> scores;
  for (int l = 0; l < p->nlevels; l++) {                                   // src/orb.cpp:61
    int w, h;
    orc_level_size(W, H, p->scale_factor, l, &w, &h);
    lvl.resize((size_t)w * h);
    orc_build_level(img, W, H, pitch, p, l, lvl.data());
    scores.resize((size_t)w * h);
    orc_fast_scores(lvl.data(), w, h, w, p->fast_threshold, p->fast_n, scores.data());
    std::vector<orc_keypoint> cand((size_t)w * h);
    int nc = orc_nms(scores.data(), w, h, p->nms_window, w * h, cand.data());   // D4: all survivors
    cand.resize(nc);
    std::vector<orc_keypoint> keep;
    std::vector<float> keepR;
    if (p->select_policy == 0) {                                           // src/orb_cpu.cpp:110
      int m = std::min(p->nfeatures, nc);
      keep.assign(cand.begin(), cand.begin() + m);
      keepR.assign(m, 0.0f);
    } else {                                                               // src/orb.cpp:62-86 under D4-D6
      int quota = orc_level_quota(p->nfeatures, p->scale_factor, p->nlevels, l);
      std::vector<float> R(nc);
      orc_harris(lvl.data(), w, h, w, cand.data(), nc, p->harris_k, R.data());
      std::vector<int> order(nc);
      for (int i = 0; i < nc; i++) order[i] = i;
      std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (R[a] != R[b]) return R[a] > R[b];
        if (cand[a].y != cand[b].y) return cand[a].y < cand[b].y;
        return cand[a].x < cand[b].x;
      });
      int m = std::max(0, std::min(quota, nc));
      order.resize(m);
      std::sort(order.begin(), order.end());                               // candidates are in raster order
      for (int i : order) { keep.push_back(cand[i]); keepR.push_back(R[i]); }
    }
    int m = (int)keep.size();
    std::vector<float> th(m);
    std::vector<orc_descriptor> de(m);
    orc_orientations(lvl.data(), w, h, w, keep.data(), m, p->orient_patch, th.data());
    orc_brief(lvl.data(), w, h, w, keep.data(), th.data(), m, de.data());   // on pyramid[l] (D8)
    float sc = orc_level_scale(p->scale_factor, l);                        // src/orb.cpp:95
    int wrote = 0;
    for (int i = 0; i < m && total < cap; i++, total++, wrote++) {
      orc_keypoint k = keep[i];
      if (level_xy) level_xy[total] = k;
      if (level_id) level_id[total] = l;
      if (response) response[total] = keepR[i];
      int gx = k.x, gy = k.y;
      gx = (int)(gx * sc);   // kp.x *= scale;   src/orb.cpp:96-97 (int * float, truncated)
      gy = (int)(gy * sc);
      if (kps) kps[total] = {gx, gy};
      if (angles) angles[total] = th[i];
      if (desc) desc[total] = de[i];
    }
    if (n_per_level) n_per_level[l] = wrote;
  }
  return total;
}

void orc_match_knn2(const orc_descriptor* query, int nq, const orc_descriptor* train, int nt, int32_t* out4, float ratio,
                    uint8_t* keep) {
  for (int i = 0; i < nq; i++) {
    int d1 = 0x7fffffff, d2 = 0x7fffffff, i1 = -1, i2 = -1;
    for (int j = 0; j < nt; j++) {
      int d = 0;
      for (int k = 0; k < 32; k++) d += __builtin_popcount((unsigned)(query[i].data[k] ^ train[j].data[k]));   // as src/compare.cpp:95-97
      if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
      else if (d < d2) { d2 = d; i2 = j; }
    }
    out4[4 * i] = i1; out4[4 * i + 1] = d1; out4[4 * i + 2] = i2; out4[4 * i + 3] = d2;
    // if (m.distance < 0.8 * n.distance): float distances, double product (src/feature_matching.cpp:178)
    if (keep) keep[i] = i2 >= 0 && (double)(float)d1 < (double)ratio * (double)(float)d2;
  }
}

int orc_detect_and_compute_batch(const uint8_t* frames, int n_frames, size_t frame_stride, int W, int H, size_t pitch,
                                 const orc_params* p, int cap, orc_keypoint* kps, float* angles, orc_descriptor* desc,
                                 int* n_out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  n_threads = std::min(n_threads, std::max(1, n_frames));
  auto work = [&](int t) {
    int per = (n_frames + n_threads - 1) / n_threads;
    int lo = t * per, hi = std::min(n_frames, lo + per);
    std::vector<orc_keypoint> k(cap); std::vector<float> a(cap); std::vector<orc_descriptor> d(cap);
    for (int f = lo; f < hi; f++) {
      orc_keypoint* ko = kps ? kps + (size_t)f * cap : k.data();
      float* ao = angles ? angles + (size_t)f * cap : a.data();
      orc_descriptor* dd = desc ? desc + (size_t)f * cap : d.data();
      int n = orc_detect_and_compute(frames + (size_t)f * frame_stride, W, H, pitch, p, cap, ko, ao, dd,
                                     nullptr, nullptr, nullptr, nullptr);
      if (n_out) n_out[f] = n;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  return 0;
}

}  // extern "C"

// =============================================================================================
// Pyramidal Lucas-Kanade.  The reference calls
//   cv::calcOpticalFlowPyrLK(img_1, img_2, points1, points2, status, err, Size(21,21), 3,
//                            TermCriteria(COUNT+EPS, 30, 0.01), 0, 0.001)      (src/feature_tracking.cpp:174-180)
// OpenCV is a third-party dependency that is absent from /root/reference, so this restates the published algorithm of
// OpenCV 4.x modules/video/src/lkpyramid.cpp (scalar path): pyrDown pyramid (5x5 [1 4 6 4 1], REFLECT_101), Scharr
// derivatives as int16 (zero outside the image: derivBorder = BORDER_CONSTANT), 14-bit fixed-point bilinear window samples
// (intensity with 5 fractional bits), float normal equations, the iteration / termination rules, the L1 error.  Float sums
// over the window are taken in a fixed order -- 32 interleaved partial sums (element i goes to partial i % 32) combined by
// an xor-butterfly 16, 8, 4, 2, 1 -- which is the order the CUDA kernel uses, so that product and checker agree bit for bit;
// OpenCV's own SIMD order differs in the last float bits, hence the tolerance when this checker is pinned against cv2.
namespace {
inline int lk_reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}
struct LkLevel { int w, h; std::vector<uint8_t> img; };
inline int lk_px(const LkLevel& L, int x, int y) { return L.img[(size_t)lk_reflect101(y, L.h) * L.w + lk_reflect101(x, L.w)]; }
// Scharr derivative of the level at (x, y): calcSharrDeriv with its REFLECT_101 rows / columns; 0 outside the image
inline void lk_deriv(const LkLevel& L, int x, int y, int* dx, int* dy) {
  if (x < 0 || x >= L.w || y < 0 || y >= L.h) { *dx = 0; *dy = 0; return; }
  int t0[3], t1[3];
  for (int k = -1; k <= 1; k++) {
    const int a = lk_px(L, x + k, y - 1), b = lk_px(L, x + k, y), c = lk_px(L, x + k, y + 1);
    t0[k + 1] = (a + c) * 3 + b * 10;
    t1[k + 1] = c - a;
  }
  *dx = t0[2] - t0[0];
  *dy = (t1[2] + t1[0]) * 3 + t1[1] * 10;
}
inline int lk_descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }
inline float lk_tree_sum(float* part) {
  for (int d = 16; d; d >>= 1)
    for (int i = 0; i < 32; i++) if (!(i & d)) { const float s = part[i] + part[i ^ d]; part[i] = s; part[i ^ d] = s; }
  return part[0];
}
void lk_pyr_down(const LkLevel& S, LkLevel* D) {
  D->w = (S.w + 1) / 2; D->h = (S.h + 1) / 2;
  D->img.resize((size_t)D->w * D->h);
  static const int k[5] = {1, 4, 6, 4, 1};
  for (int y = 0; y < D->h; y++)
    for (int x = 0; x < D->w; x++) {
      int s = 0;
      for (int j = 0; j < 5; j++) {
        int r = 0;
        for (int i = 0; i < 5; i++) r += k[i] * lk_px(S, 2 * x + i - 2, 2 * y + j - 2);
        s += k[j] * r;
      }
      D->img[(size_t)y * D->w + x] = (uint8_t)((s + 128) >> 8);
    }
}
}  // namespace

int orc_lk_levels(int w, int h, int win, int max_level) {
  int L = 0;
  while (L < max_level) {
    w = (w + 1) / 2; h = (h + 1) / 2;
    if (w <= win || h <= win) break;
    L++;
  }
  return L;   // highest level index actually used
}

void orc_lk_pyr_down(const uint8_t* src, int w, int h, size_t pitch, uint8_t* dst) {
  LkLevel S{w, h, {}}, D;
  S.img.resize((size_t)w * h);
  for (int y = 0; y < h; y++) memcpy(&S.img[(size_t)y * w], src + (size_t)y * pitch, w);
  lk_pyr_down(S, &D);
  memcpy(dst, D.img.data(), D.img.size());
}

void orc_lk_track(const uint8_t* prev, const uint8_t* next, int w, int h, size_t pitch, const float* prev_pts, int n, int win,
                  int max_level, int max_iter, double eps, float min_eig, float* next_pts, uint8_t* status, float* err) {
  // TermCriteria handling of calcOpticalFlowPyrLK
  max_iter = std::min(std::max(max_iter, 0), 100);
  eps = std::min(std::max(eps, 0.), 10.);
  eps *= eps;
  const int top = orc_lk_levels(w, h, win, max_level);
  std::vector<LkLevel> P(top + 1), N(top + 1);
  P[0].w = N[0].w = w; P[0].h = N[0].h = h;
  P[0].img.resize((size_t)w * h); N[0].img.resize((size_t)w * h);
  for (int y = 0; y < h; y++) {
    memcpy(&P[0].img[(size_t)y * w], prev + (size_t)y * pitch, w);
    memcpy(&N[0].img[(size_t)y * w], next + (size_t)y * pitch, w);
  }
  for (int l = 1; l <= top; l++) { lk_pyr_down(P[l - 1], &P[l]); lk_pyr_down(N[l - 1], &N[l]); }
  const int area = win * win;
  std::vector<short> Iw(area), dIw(2 * area);
  const float halfWin = (win - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  for (int i = 0; i < n; i++) { status[i] = 1; if (err) err[i] = 0; }
  for (int pt = 0; pt < n; pt++) {
    float nx = 0, ny = 0;
    for (int level = top; level >= 0; level--) {
      const LkLevel& I = P[level];
      const LkLevel& J = N[level];
      const float sc = (float)(1. / (1 << level));
      float px = prev_pts[2 * pt] * sc, py = prev_pts[2 * pt + 1] * sc;
      if (level == top) { nx = px; ny = py; } else { nx = next_pts[2 * pt] * 2.f; ny = next_pts[2 * pt + 1] * 2.f; }
      next_pts[2 * pt] = nx; next_pts[2 * pt + 1] = ny;
      px -= halfWin; py -= halfWin;
      const int ipx = (int)std::floor(px), ipy = (int)std::floor(py);
      if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
        if (level == 0) { status[pt] = 0; if (err) err[pt] = 0; }
        continue;
      }
      float a = px - ipx, b = py - ipy;
      int iw00 = (int)std::nearbyint((1.f - a) * (1.f - b) * (1 << 14));
      int iw01 = (int)std::nearbyint(a * (1.f - b) * (1 << 14));
      int iw10 = (int)std::nearbyint((1.f - a) * b * (1 << 14));
      int iw11 = (1 << 14) - iw00 - iw01 - iw10;
      float p11[32] = {0}, p12[32] = {0}, p22[32] = {0};
      for (int y = 0; y < win; y++)
        for (int x = 0; x < win; x++) {
          const int X = ipx + x, Y = ipy + y;
          const int ival = lk_descale(lk_px(I, X, Y) * iw00 + lk_px(I, X + 1, Y) * iw01 + lk_px(I, X, Y + 1) * iw10 +
                                      lk_px(I, X + 1, Y + 1) * iw11, 14 - 5);
          int d00x, d00y, d01x, d01y, d10x, d10y, d11x, d11y;
          lk_deriv(I, X, Y, &d00x, &d00y); lk_deriv(I, X + 1, Y, &d01x, &d01y);
          lk_deriv(I, X, Y + 1, &d10x, &d10y); lk_deriv(I, X + 1, Y + 1, &d11x, &d11y);
          const int ixval = lk_descale(d00x * iw00 + d01x * iw01 + d10x * iw10 + d11x * iw11, 14);
          const int iyval = lk_descale(d00y * iw00 + d01y * iw01 + d10y * iw10 + d11y * iw11, 14);
          const int e = y * win + x;
          Iw[e] = (short)ival; dIw[2 * e] = (short)ixval; dIw[2 * e + 1] = (short)iyval;
          p11[e & 31] += (float)(ixval * ixval); p12[e & 31] += (float)(ixval * iyval); p22[e & 31] += (float)(iyval * iyval);
        }
      const float A11 = lk_tree_sum(p11) * FLT_SCALE, A12 = lk_tree_sum(p12) * FLT_SCALE, A22 = lk_tree_sum(p22) * FLT_SCALE;
      float D = A11 * A22 - A12 * A12;
      const float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
      if (minEig < min_eig || D < 1.1920928955078125e-07f) {   // FLT_EPSILON
        if (level == 0) status[pt] = 0;
        continue;
      }
      D = 1.f / D;
      nx -= halfWin; ny -= halfWin;
      float pdx = 0, pdy = 0;
      for (int j = 0; j < max_iter; j++) {
        const int inx = (int)std::floor(nx), iny = (int)std::floor(ny);
        if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
          if (level == 0) status[pt] = 0;
          break;
        }
        a = nx - inx; b = ny - iny;
        iw00 = (int)std::nearbyint((1.f - a) * (1.f - b) * (1 << 14));
        iw01 = (int)std::nearbyint(a * (1.f - b) * (1 << 14));
        iw10 = (int)std::nearbyint((1.f - a) * b * (1 << 14));
        iw11 = (1 << 14) - iw00 - iw01 - iw10;
        float q1[32] = {0}, q2[32] = {0};
        for (int y = 0; y < win; y++)
          for (int x = 0; x < win; x++) {
            const int X = inx + x, Y = iny + y, e = y * win + x;
            const int diff = lk_descale(lk_px(J, X, Y) * iw00 + lk_px(J, X + 1, Y) * iw01 + lk_px(J, X, Y + 1) * iw10 +
                                        lk_px(J, X + 1, Y + 1) * iw11, 14 - 5) - Iw[e];
            q1[e & 31] += (float)(diff * dIw[2 * e]); q2[e & 31] += (float)(diff * dIw[2 * e + 1]);
          }
        const float b1 = lk_tree_sum(q1) * FLT_SCALE, b2 = lk_tree_sum(q2) * FLT_SCALE;
        const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
        nx += dx; ny += dy;
        next_pts[2 * pt] = nx + halfWin; next_pts[2 * pt + 1] = ny + halfWin;
        if ((double)dx * dx + (double)dy * dy <= eps) break;
        if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) {
          next_pts[2 * pt] -= dx * 0.5f; next_pts[2 * pt + 1] -= dy * 0.5f;
          break;
        }
        pdx = dx; pdy = dy;
      }
      if (status[pt] && err && level == 0) {
        const float ex = next_pts[2 * pt] - halfWin, ey = next_pts[2 * pt + 1] - halfWin;
        const int inx = (int)std::floor(ex), iny = (int)std::floor(ey);
        if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) { status[pt] = 0; continue; }
        a = ex - inx; b = ey - iny;
        iw00 = (int)std::nearbyint((1.f - a) * (1.f - b) * (1 << 14));
        iw01 = (int)std::nearbyint(a * (1.f - b) * (1 << 14));
        iw10 = (int)std::nearbyint((1.f - a) * b * (1 << 14));
        iw11 = (1 << 14) - iw00 - iw01 - iw10;
        float ev[32] = {0};
        for (int y = 0; y < win; y++)
          for (int x = 0; x < win; x++) {
            const int X = inx + x, Y = iny + y, e = y * win + x;
            const int diff = lk_descale(lk_px(J, X, Y) * iw00 + lk_px(J, X + 1, Y) * iw01 + lk_px(J, X, Y + 1) * iw10 +
                                        lk_px(J, X + 1, Y + 1) * iw11, 14 - 5) - Iw[e];
            ev[e & 31] += std::abs((float)diff);
          }
        err[pt] = lk_tree_sum(ev) * 1.f / (32 * win * win);
      }
    }
  }
}
