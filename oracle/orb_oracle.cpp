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
