// orb_kernels.cuh -- sm_100a kernels of the ORB hot path (u8 end to end, no tensor cores: nothing
// here is a dense contraction).  Three launches per chunk of frames:
//   k_pyramid_fast : per (frame, level, tile): bilinear resize from level 0 + 5x5 Gaussian in shared
//                    memory, level pixels written once, FAST-n segment test + SAD score + 3x3 NMS on the
//                    tile still in shared memory, Harris response of the survivors, 5x5 box-sum image
//                    for BRIEF.  Replaces ref ORB::buildPyramid (src/orb.cpp:111-120 / src/orb_cpu.cpp:
//                    278-290), d_Fast (src/cuda/Fast.cu:30-209), d_NMS (src/cuda/NMS.cu:21-128),
//                    HarrisScore (src/cuda/HarrisScore.cu:23-89) and cv::integral (src/cuda/Brief.cu:101-105).
//   k_select       : per (frame, level): exact top-quota selection under the total order
//                    (response desc, y asc, x asc) by 64-bit radix select, then raster sort
//                    (ref std::nth_element at src/orb.cpp:73-86; raster cap at src/orb_cpu.cpp:110).
//   k_describe     : one warp per kept keypoint: intensity-centroid orientation (ref d_Orientations,
//                    src/cuda/Orientations.cu:22-63 == src/orb_cpu.cpp:139-183) and rotated BRIEF with
//                    ballot-packed words (ref d_Brief, src/cuda/Brief.cu:40-95 == src/orb_cpu.cpp:203-258).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_math.cuh"
#include "orb_plan.h"

namespace orbk {

struct Bufs {
  const uint8_t* frames;         // level 0 of the chunk's first frame
  unsigned long long frame_stride;
  int pitch0;
  uint8_t* pyr;                  // [chunk][pyr_frame_bytes]
  uint16_t* box;                 // [chunk][box_frame_elems]
  unsigned long long* cand;      // [chunk][cand_frame_elems]
  int* cand_count;               // [chunk][ORB_MAX_LEVELS]
  uint32_t* kept_xy;             // [chunk][kept_per_frame]   (y << 16 | x), level space
  float* kept_r;                 // [chunk][kept_per_frame]
  int* kept_count;               // [chunk][ORB_MAX_LEVELS]
  const OrbTap* xtab;
  const OrbTap* ytab;
  const float* harris_w;         // 49 window weights
  const char4* pattern;          // 256 BRIEF tests (x1,y1,x2,y2)
  int* flags;                    // bit 0: candidate overflow
  orb_keypoint* out_kps;         // [chunk][out_cap]
  float* out_angles;
  orb_descriptor* out_desc;
  int* out_n;                    // [chunk]
  int out_cap;
  orb_keypoint* side_xy;         // nullable side arrays, [chunk][out_cap]
  int* side_level;
  float* side_resp;
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// order-preserving map float -> u32 (ascending)
__device__ __forceinline__ uint32_t f2ord(float r) {
  uint32_t u = (uint32_t)__float_as_int(__fadd_rn(r, 0.0f));
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __int_as_float((int)u);
}


// ---------------------------------------------------------------------------------------------
// Harris response at a pixel, decision D5: integer 3x3 Sobel, float 7x7 window in
// row-major order, separate multiply and add (no FMA), det - k * trace * trace.
// PIX(y, x) must return the pixel of the reflect-101 extended level.
template <typename PIX>
__device__ __forceinline__ float harris_at(PIX pix, int y, int x, const float* __restrict__ wt, float k) {
  float A = 0.f, B = 0.f, C = 0.f;
#pragma unroll 1
  for (int dy = -3; dy <= 3; dy++) {
    // sliding 3-column window of vertical sums along the row
    int yy = y + dy;
    int a0 = pix(yy - 1, x - 4), a1 = pix(yy, x - 4), a2 = pix(yy + 1, x - 4);
    int b0 = pix(yy - 1, x - 3), b1 = pix(yy, x - 3), b2 = pix(yy + 1, x - 3);
#pragma unroll
    for (int dx = -3; dx <= 3; dx++) {
      int xx = x + dx;
      int c0 = pix(yy - 1, xx + 1), c1 = pix(yy, xx + 1), c2 = pix(yy + 1, xx + 1);
      int ix = (c0 + 2 * c1 + c2) - (a0 + 2 * a1 + a2);
      int iy = (a2 + 2 * b2 + c2) - (a0 + 2 * b0 + c0);
      float g = __ldg(wt + (dy + 3) * 7 + (dx + 3));
      A = orbm::fadd(A, orbm::fmul((float)(ix * ix), g));
      B = orbm::fadd(B, orbm::fmul((float)(ix * iy), g));
      C = orbm::fadd(C, orbm::fmul((float)(iy * iy), g));
      a0 = b0; a1 = b1; a2 = b2;
      b0 = c0; b1 = c1; b2 = c2;
    }
  }
  float det = orbm::fsub(orbm::fmul(A, C), orbm::fmul(B, B));
  float tr = orbm::fadd(A, C);
  return orbm::fsub(det, orbm::fmul(orbm::fmul(k, tr), tr));
}

// FAST-n corner score of one pixel (ref src/orb_cpu.cpp:37-100): 0 if not a corner.
// PIX(dy, dx) returns the pixel at an offset from the centre.
template <typename PIX>
__device__ __forceinline__ int fast_score(PIX pix, int thr, int n) {
  int Ip = pix(0, 0);
  int hi = Ip + thr, lo = Ip - thr;
  int v0 = pix(-3, 0), v4 = pix(0, 3), v8 = pix(3, 0), v12 = pix(0, -3);
  int br = (v0 >= hi) + (v4 >= hi) + (v8 >= hi) + (v12 >= hi);
  int dk = (v0 < hi && v0 <= lo) + (v4 < hi && v4 <= lo) + (v8 < hi && v8 <= lo) + (v12 < hi && v12 <= lo);
  if (max(br, dk) < 3) return 0;
  // FAST ring in the reference's order (src/orb_cpu.cpp:8-13)
  constexpr int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  constexpr int ring_dy[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};
  uint32_t mb = 0, md = 0;
  int sad = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int v = pix(ring_dy[i], ring_dx[i]);
    mb |= (uint32_t)(v >= hi) << i;
    md |= (uint32_t)(v <= lo) << i;
    sad += abs(Ip - v);
  }
  mb |= mb << 16;
  md |= md << 16;
  uint32_t rb = mb, rd = md;
  for (int j = 1; j < n; j++) { rb &= mb >> j; rd &= md >> j; }
  return ((rb | rd) & 0xffffu) ? sad : 0;
}

// ---------------------------------------------------------------------------------------------
template <int TW_, int TH_>
struct Tile {
  static constexpr int TW = TW_, TH = TH_;
  static constexpr int RW = TW + 12, RH = TH + 12;   // resized pixels, halo 6 (4 + blur 2)
  static constexpr int PW = TW + 8, PH = TH + 8;     // level pixels, halo 4 (FAST 3 + NMS 1)
  static constexpr int SW = TW + 2, SH = TH + 2;     // scores, halo 1
  static constexpr int RWP = (RW + 3) & ~3, PWP = (PW + 3) & ~3;
  static constexpr int LIST_CAP = 2048;
  static constexpr int RES_BYTES = RH * RWP;                 // u8
  static constexpr int TMP_BYTES = RH * PW * 2;              // u16: h-blur, then scores, then box rows
  static constexpr int PIX_BYTES = PH * PWP;                 // u8
  static constexpr int SMEM = RES_BYTES + TMP_BYTES + PIX_BYTES + LIST_CAP * 2 + 16;
  static_assert(SH * SW * 2 <= TMP_BYTES && (TH + 4) * TW * 2 <= TMP_BYTES, "tmp aliasing");
  static_assert(TW % 4 == 0, "tile width");
};

constexpr int K1_THREADS = 256;

template <class T>
__global__ void __launch_bounds__(K1_THREADS) k_pyramid_fast(const OrbPlan P, const Bufs B) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* s_res = smem;
  uint16_t* s_tmp = (uint16_t*)(smem + T::RES_BYTES);
  uint8_t* s_pix = smem + T::RES_BYTES + T::TMP_BYTES;
  uint16_t* s_list = (uint16_t*)(s_pix + T::PIX_BYTES);
  int* s_ctr = (int*)(s_list + T::LIST_CAP);   // [0] list count, [1] global base

  const int tid = threadIdx.x;
  const int f = blockIdx.y;
  int t = blockIdx.x, l = 0;
  while (l + 1 < P.nlevels && t >= P.lv[l + 1].tile_ofs) l++;
  const OrbLevel& G = P.lv[l];
  t -= G.tile_ofs;
  const int x0 = (t % G.tiles_x) * T::TW, y0 = (t / G.tiles_x) * T::TH;
  const int w = G.w, h = G.h;
  const uint8_t* __restrict__ src = B.frames + (size_t)f * B.frame_stride;
  const int sp = B.pitch0;
  if (tid == 0) s_ctr[0] = 0;

  // ---- phase 1: level pixels with halo into shared memory ---------------------------------
  if (l == 0) {
    for (int i = tid; i < T::PH * T::PW; i += K1_THREADS) {
      int py = i / T::PW, px = i - py * T::PW;
      int ly = y0 - 4 + py, lx = x0 - 4 + px;
      if (ly >= h + 4 || lx >= w + 4) continue;
      s_pix[py * T::PWP + px] = __ldg(src + (size_t)reflect101(ly, h) * sp + reflect101(lx, w));
    }
  } else {
    const OrbTap* __restrict__ xt = B.xtab + G.xtab_ofs;
    const OrbTap* __restrict__ yt = B.ytab + G.ytab_ofs;
    const bool blur = P.blur_levels != 0;
    // resized pixels (cv::resize INTER_LINEAR restated: 11-bit taps, >>4, >>16, +2 >>2)
    const int halo = blur ? 6 : 4;
    const int rw = T::TW + 2 * halo, rh = T::TH + 2 * halo;
    for (int i = tid; i < rw * rh; i += K1_THREADS) {
      int ry = i / rw, rx = i - ry * rw;
      int ly = y0 - halo + ry, lx = x0 - halo + rx;
      if (ly >= h + halo || lx >= w + halo) continue;
      OrbTap ty = yt[reflect101(ly, h)], tx = xt[reflect101(lx, w)];
      const uint8_t* r0 = src + (size_t)ty.s0 * sp;
      const uint8_t* r1 = src + (size_t)ty.s1 * sp;
      int h0 = __ldg(r0 + tx.s0) * tx.a0 + __ldg(r0 + tx.s1) * tx.a1;
      int h1 = __ldg(r1 + tx.s0) * tx.a0 + __ldg(r1 + tx.s1) * tx.a1;
      int v = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
      v = min(v, 255);
      if (blur) s_res[ry * T::RWP + rx] = (uint8_t)v;
      else s_pix[ry * T::PWP + rx] = (uint8_t)v;
    }
    if (blur) {
      __syncthreads();
      // cv::GaussianBlur 5x5 sigma 0 on u8: [1 4 6 4 1] both ways, one rounding: (sum + 128) >> 8
      for (int i = tid; i < T::RH * T::PW; i += K1_THREADS) {
        int ry = i / T::PW, px = i - ry * T::PW;
        const uint8_t* r = s_res + ry * T::RWP + px;
        s_tmp[i] = (uint16_t)(r[0] + 4 * r[1] + 6 * r[2] + 4 * r[3] + r[4]);
      }
      __syncthreads();
      for (int i = tid; i < T::PH * T::PW; i += K1_THREADS) {
        int py = i / T::PW, px = i - py * T::PW;
        const uint16_t* c = s_tmp + py * T::PW + px;
        int s = c[0] + 4 * c[T::PW] + 6 * c[2 * T::PW] + 4 * c[3 * T::PW] + c[4 * T::PW];
        s_pix[py * T::PWP + px] = (uint8_t)((s + 128) >> 8);
      }
    }
  }
  __syncthreads();

  // ---- phase 2: write the level once (levels >= 1) ----------------------------------------
  if (l > 0) {
    uint8_t* dst = B.pyr + (size_t)f * P.pyr_frame_bytes + G.lvl_ofs;
    const int nrow = min(T::TH, h - y0), nwords = (min(T::TW, w - x0) + 3) >> 2;
    for (int i = tid; i < nrow * (T::TW / 4); i += K1_THREADS) {
      int iy = i / (T::TW / 4), wx = i - iy * (T::TW / 4);
      if (wx >= nwords) continue;
      uint32_t v = *(const uint32_t*)(s_pix + (iy + 4) * T::PWP + 4 + wx * 4);
      *(uint32_t*)(dst + (size_t)(y0 + iy) * G.pitch + x0 + wx * 4) = v;
    }
  }

  // ---- phase 3: FAST-n score on the tile + 1 halo ------------------------------------------
  const int thr = P.fast_threshold, fn = P.fast_n;
  for (int i = tid; i < T::SH * T::SW; i += K1_THREADS) {
    int sy = i / T::SW, sx = i - sy * T::SW;
    int cy = y0 - 1 + sy, cx = x0 - 1 + sx;
    int sc = 0;
    if (cx >= 3 && cx < w - 3 && cy >= 3 && cy < h - 3) {
      const uint8_t* c = s_pix + (sy + 3) * T::PWP + (sx + 3);
      sc = fast_score([&](int dy, int dx) { return (int)c[dy * T::PWP + dx]; }, thr, fn);
    }
    s_tmp[i] = (uint16_t)sc;
  }
  __syncthreads();

  // ---- phase 4: NMS (ties keep both, ref src/orb_cpu.cpp:126) + candidate emission ---------
  unsigned long long* cand = B.cand + (size_t)f * P.cand_frame_elems + G.cand_ofs;
  int* gcount = B.cand_count + f * ORB_MAX_LEVELS + l;
  const int nmsr = P.nms_radius;
  auto pixat = [&](int yy, int xx) { return (int)s_pix[yy * T::PWP + xx]; };
  auto emit = [&](int iy, int ix, int slot) {
    int lx = x0 + ix, ly = y0 + iy;
    uint32_t hi = 0;
    if (P.select_policy == ORB_SELECT_HARRIS_TOP_N) {
      float r = harris_at(pixat, iy + 4, ix + 4, B.harris_w, P.harris_k);
      hi = ~f2ord(r);
    }
    if (slot < G.cand_cap)
      cand[slot] = ((unsigned long long)hi << 32) | (unsigned)((ly << 16) | lx);
  };
  for (int i = tid; i < T::TH * T::TW; i += K1_THREADS) {
    int iy = i / T::TW, ix = i - iy * T::TW;
    if (y0 + iy >= h || x0 + ix >= w) continue;
    const uint16_t* s = s_tmp + (iy + 1) * T::SW + (ix + 1);
    int v = s[0];
    if (v == 0) continue;
    bool keep = true;
    if (nmsr) {
      int m = max(max(max(s[-T::SW - 1], s[-T::SW]), max(s[-T::SW + 1], s[-1])),
                  max(max(s[1], s[T::SW - 1]), max(s[T::SW], s[T::SW + 1])));
      keep = v >= m;
    }
    if (!keep) continue;
    int slot = atomicAdd(&s_ctr[0], 1);
    if (slot < T::LIST_CAP) s_list[slot] = (uint16_t)i;
    else emit(iy, ix, atomicAdd(gcount, 1));   // list full: finish this survivor inline
  }
  __syncthreads();
  const int nlist = min(s_ctr[0], T::LIST_CAP);
  if (nlist > 0) {
    if (tid == 0) s_ctr[1] = atomicAdd(gcount, nlist);
    __syncthreads();
    const int base = s_ctr[1];
    for (int j = tid; j < nlist; j += K1_THREADS) {
      int i = s_list[j];
      int iy = i / T::TW, ix = i - iy * T::TW;
      emit(iy, ix, base + j);
    }
  }
  __syncthreads();

  // ---- phase 5: 5x5 box sums of the level for BRIEF (replaces the int32 integral image) ----
  for (int i = tid; i < (T::TH + 4) * T::TW; i += K1_THREADS) {
    int r = i / T::TW, ix = i - r * T::TW;
    const uint8_t* p = s_pix + (r + 2) * T::PWP + ix + 2;
    s_tmp[i] = (uint16_t)(p[0] + p[1] + p[2] + p[3] + p[4]);
  }
  __syncthreads();
  uint16_t* box = B.box + (size_t)f * P.box_frame_elems + G.box_ofs;
  {
    const int nrow = min(T::TH, h - y0), npair = (min(T::TW, w - x0) + 1) >> 1;
    for (int i = tid; i < nrow * (T::TW / 2); i += K1_THREADS) {
      int iy = i / (T::TW / 2), px = i - iy * (T::TW / 2);
      if (px >= npair) continue;
      const uint16_t* c = s_tmp + iy * T::TW + px * 2;
      uint32_t lo = c[0] + c[T::TW] + c[2 * T::TW] + c[3 * T::TW] + c[4 * T::TW];
      uint32_t hi = c[1] + c[T::TW + 1] + c[2 * T::TW + 1] + c[3 * T::TW + 1] + c[4 * T::TW + 1];
      *(uint32_t*)(box + (size_t)(y0 + iy) * G.bpitch + x0 + px * 2) = lo | (hi << 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Selection: CTA per (frame, level).
constexpr int K2_THREADS = 512;

__global__ void __launch_bounds__(K2_THREADS) k_select(const OrbPlan P, const Bufs B) {
  extern __shared__ __align__(16) unsigned long long s_sort[];   // [npow2]
  __shared__ int s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_k, s_n;
  const int tid = threadIdx.x, l = blockIdx.x, f = blockIdx.y;
  const OrbLevel& G = P.lv[l];
  int n = B.cand_count[f * ORB_MAX_LEVELS + l];
  if (n > G.cand_cap) {
    if (tid == 0) atomicOr(B.flags, 1);
    n = G.cand_cap;
  }
  const int m = min(G.quota, n);
  const unsigned long long* __restrict__ keys = B.cand + (size_t)f * P.cand_frame_elems + G.cand_ofs;
  unsigned long long T = ~0ull;
  if (n > m && m > 0) {
    // m-th smallest key by MSB-first radix select, 8 bits per pass
    if (tid == 0) { s_prefix = 0; s_k = m; }
    for (int pass = 7; pass >= 0; pass--) {
      for (int i = tid; i < 256; i += K2_THREADS) s_hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const int shift = pass * 8;
      for (int i = tid; i < n; i += K2_THREADS) {
        unsigned long long k = keys[i];
        if (pass == 7 || (k >> (shift + 8)) == prefix) atomicAdd(&s_hist[(int)(k >> shift) & 255], 1);
      }
      __syncthreads();
      if (tid < 32) {
        int local[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { local[j] = s_hist[tid * 8 + j]; sum += local[j]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (tid >= d) incl += v;
        }
        const int k = s_k;
        unsigned hit = __ballot_sync(0xffffffffu, incl >= k);
        int lane = __ffs(hit) - 1;
        if (tid == lane) {
          int before = incl - sum, b = 0;
          for (; b < 8; b++) { if (before + local[b] >= k) break; before += local[b]; }
          s_k = k - before;
          s_prefix = (prefix << 8) | (unsigned)(tid * 8 + b);
        }
      }
      __syncthreads();
    }
    T = s_prefix;
  }
  // gather the kept keys as (raster << 32 | response order) and sort by raster
  if (tid == 0) s_n = 0;
  int npow2 = 1;
  while (npow2 < m) npow2 <<= 1;
  for (int i = tid; i < npow2; i += K2_THREADS) s_sort[i] = ~0ull;
  __syncthreads();
  if (m > 0)
    for (int i = tid; i < n; i += K2_THREADS) {
      unsigned long long k = keys[i];
      if (k <= T) {
        int pos = atomicAdd(&s_n, 1);
        if (pos < npow2) s_sort[pos] = (k << 32) | (k >> 32);
      }
    }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (npow2 >> 1); i += K2_THREADS) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool up = (lo & size) == 0;
        unsigned long long a = s_sort[lo], b = s_sort[hi];
        if ((a > b) == up) { s_sort[lo] = b; s_sort[hi] = a; }
      }
      __syncthreads();
    }
  uint32_t* kxy = B.kept_xy + (size_t)f * P.kept_per_frame + G.kept_ofs;
  float* kr = B.kept_r + (size_t)f * P.kept_per_frame + G.kept_ofs;
  for (int i = tid; i < m; i += K2_THREADS) {
    unsigned long long v = s_sort[i];
    kxy[i] = (uint32_t)(v >> 32);
    kr[i] = P.select_policy == ORB_SELECT_HARRIS_TOP_N ? ord2f(~(uint32_t)v) : 0.0f;
  }
  if (tid == 0) B.kept_count[f * ORB_MAX_LEVELS + l] = m;
}

// ---------------------------------------------------------------------------------------------
// Orientation + rotated BRIEF: one warp per keypoint.
constexpr int K3_WARPS = 8;

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// sum of the u8 image over rows [ya,yb) x cols [xa,xb), whole warp cooperates
__device__ __forceinline__ int warp_rect_sum(const uint8_t* __restrict__ img, int pitch, int xa, int xb, int ya, int yb, int lane) {
  int s = 0;
  for (int y = ya; y < yb; y++)
    for (int x = xa + lane; x < xb; x += 32) s += img[(size_t)y * pitch + x];
  return warp_sum(s);
}

// 5x5 box value of the reference's sum5x5 (src/orb_cpu.cpp:190-201) for a centre whose box leaves the
// image on the right / bottom: the reference indexes its (H+1)x(W+1) integral image flat, so column
// overruns wrap into the next row and row overruns fall off the end (decision D7: those read 0).
__device__ __forceinline__ int box_edge(const uint8_t* __restrict__ img, int pitch, int W, int H, int cx, int cy, int lane) {
  const int x0 = cx - 2, y0 = cy - 2, x1 = cx + 3, y1 = cy + 3;
  const bool xo = x1 > W, yo = y1 > H;
  if (!xo) return -warp_rect_sum(img, pitch, x0, x1, 0, y0, lane);   // yo only
  const int col = x1 - W - 1;                                        // wrapped column: 0 or 1
  if (!yo) {
    int s = -warp_rect_sum(img, pitch, 0, x0, y0, y1, lane);
    if (col == 1)
      s += (y1 + 1 <= H) ? warp_rect_sum(img, pitch, 0, 1, y0 + 1, y1 + 1, lane)
                         : -warp_rect_sum(img, pitch, 0, 1, 0, y0 + 1, lane);
    return s;
  }
  int s = warp_rect_sum(img, pitch, 0, x0, 0, y0, lane);
  if (col == 1) s -= warp_rect_sum(img, pitch, 0, 1, 0, y0 + 1, lane);
  return s;
}

struct DescribeJob {             // where the keypoints of this launch come from
  int mode;                      // 0: kept lists of the pipeline; 1: explicit list, orientation only;
                                 // 2: explicit list + given angles, descriptors only
  const orb_keypoint* list_kps;  // modes 1, 2
  const float* list_angles;      // mode 2
  int list_n;
};

__device__ __forceinline__ float orientation_of(const uint8_t* __restrict__ img, int pitch, int w, int h, int x, int y,
                                                int pr, int lane) {
  // ref src/orb_cpu.cpp:152-178; moments are exact integers (|m| < 2^24), so integer accumulation in any
  // order equals the reference's float accumulation
  if (x - pr < 0 || x + pr >= w || y - pr < 0 || y + pr >= h) return 0.0f;
  int m10 = 0, m01 = 0;
  for (int r = -pr; r <= pr; r++) {
    const uint8_t* row = img + (size_t)(y + r) * pitch + x;
    for (int c = -pr + lane; c <= pr; c += 32) {
      int I = row[c];
      m10 += c * I;
      m01 += r * I;
    }
  }
  m10 = warp_sum(m10);
  m01 = warp_sum(m01);
  return orbm::atan2f_glibc((float)m01, (float)m10);
}

__device__ __forceinline__ void brief_of(const uint8_t* __restrict__ img, int pitch, const uint16_t* __restrict__ box,
                                         int bpitch, int W, int H, int kx, int ky, float angle,
                                         const char4* __restrict__ pattern, int lane, uint32_t* out_words) {
  const float c = orbm::cosf_glibc(angle), s = orbm::sinf_glibc(angle);   // ref src/orb_cpu.cpp:217-218
  uint32_t mine = 0;
  for (int wd = 0; wd < 8; wd++) {
    char4 t = __ldg(pattern + wd * 32 + lane);
    float x1 = (float)t.x, y1 = (float)t.y, x2 = (float)t.z, y2 = (float)t.w;
    int cx1 = kx + orbm::lround_f(orbm::fsub(orbm::fmul(c, x1), orbm::fmul(s, y1)));   // :228-237
    int cy1 = ky + orbm::lround_f(orbm::fadd(orbm::fmul(s, x1), orbm::fmul(c, y1)));
    int cx2 = kx + orbm::lround_f(orbm::fsub(orbm::fmul(c, x2), orbm::fmul(s, y2)));
    int cy2 = ky + orbm::lround_f(orbm::fadd(orbm::fmul(s, x2), orbm::fmul(c, y2)));
    // bound rule of :240-245 against the integral image dims (W+1, H+1)
    bool skip = cx1 < 2 || cy1 < 2 || cx1 > W - 1 || cy1 > H - 1 || cx2 < 2 || cy2 < 2 || cx2 > W - 1 || cy2 > H - 1;
    bool e1 = !skip && (cx1 > W - 3 || cy1 > H - 3), e2 = !skip && (cx2 > W - 3 || cy2 > H - 3);
    int s1 = 0, s2 = 0;
    if (!skip) {
      if (!e1) s1 = box[(size_t)cy1 * bpitch + cx1];
      if (!e2) s2 = box[(size_t)cy2 * bpitch + cx2];
    }
    unsigned pend = __ballot_sync(0xffffffffu, e1);
    while (pend) {
      int src = __ffs(pend) - 1;
      pend &= pend - 1;
      int v = box_edge(img, pitch, W, H, __shfl_sync(0xffffffffu, cx1, src), __shfl_sync(0xffffffffu, cy1, src), lane);
      if (lane == src) s1 = v;
    }
    pend = __ballot_sync(0xffffffffu, e2);
    while (pend) {
      int src = __ffs(pend) - 1;
      pend &= pend - 1;
      int v = box_edge(img, pitch, W, H, __shfl_sync(0xffffffffu, cx2, src), __shfl_sync(0xffffffffu, cy2, src), lane);
      if (lane == src) s2 = v;
    }
    uint32_t word = __ballot_sync(0xffffffffu, !skip && s1 < s2);   // bit i of word wd == test 32*wd + i
    if (lane == wd) mine = word;
  }
  *out_words = mine;
}

__global__ void __launch_bounds__(K3_WARPS * 32) k_describe(const OrbPlan P, const Bufs B, const DescribeJob J) {
  const int lane = threadIdx.x & 31;
  const int widx = blockIdx.x * K3_WARPS + (threadIdx.x >> 5);
  const int f = blockIdx.y;
  int l = 0, i = widx;
  int x, y;
  float resp = 0.f;
  if (J.mode == 0) {
    const int* kc = B.kept_count + f * ORB_MAX_LEVELS;
    int total = 0;
    for (int q = 0; q < P.nlevels; q++) total += kc[q];
    if (widx == 0 && lane == 0) B.out_n[f] = min(total, B.out_cap);
    if (widx >= total || widx >= B.out_cap) return;
    while (i >= kc[l]) { i -= kc[l]; l++; }
    uint32_t xy = B.kept_xy[(size_t)f * P.kept_per_frame + P.lv[l].kept_ofs + i];
    resp = B.kept_r[(size_t)f * P.kept_per_frame + P.lv[l].kept_ofs + i];
    x = xy & 0xffff; y = xy >> 16;
  } else {
    if (widx >= J.list_n) return;
    x = J.list_kps[widx].x; y = J.list_kps[widx].y;
  }
  const OrbLevel& G = P.lv[l];
  const uint8_t* img;
  int pitch;
  if (l == 0) { img = B.frames + (size_t)f * B.frame_stride; pitch = B.pitch0; }
  else { img = B.pyr + (size_t)f * P.pyr_frame_bytes + G.lvl_ofs; pitch = G.pitch; }
  const uint16_t* box = B.box + (size_t)f * P.box_frame_elems + G.box_ofs;
  const size_t o = (size_t)f * B.out_cap + widx;

  float angle;
  if (J.mode == 2) angle = J.list_angles[widx];
  else angle = orientation_of(img, pitch, G.w, G.h, x, y, P.patch_radius, lane);
  if (J.mode != 2 && lane == 0) B.out_angles[o] = angle;
  if (J.mode != 1) {
    uint32_t word;
    brief_of(img, pitch, box, G.bpitch, G.w, G.h, x, y, angle, B.pattern, lane, &word);
    if (lane < 8) ((uint32_t*)B.out_desc)[o * 8 + lane] = word;
  }
  if (J.mode == 0 && lane == 0) {
    // kp.x *= scale (int * float, truncated): ref src/orb.cpp:94-98
    orb_keypoint kp;
    kp.x = __float2int_rz(orbm::fmul((float)x, G.scale));
    kp.y = __float2int_rz(orbm::fmul((float)y, G.scale));
    B.out_kps[o] = kp;
    if (B.side_xy) { B.side_xy[o] = orb_keypoint{x, y}; B.side_level[o] = l; B.side_resp[o] = resp; }
  }
}

// Harris response for an explicit keypoint list on a level-0 image (stage entry point orb_harris)
__global__ void k_harris_list(const uint8_t* __restrict__ img, int pitch, int w, int h, const orb_keypoint* kps, int n,
                              const float* wt, float k, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto pix = [&](int yy, int xx) { return (int)img[(size_t)reflect101(yy, h) * pitch + reflect101(xx, w)]; };
  out[i] = harris_at(pix, kps[i].y, kps[i].x, wt, k);
}

// libm twins evaluated on arrays (tests/test_gpu_math.py): op 0 atan2f(a,b), 1 cosf(a), 2 sinf(a), 3 lround(a)
__global__ void k_eval_math(int op, const float* a, const float* b, int n, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r;
  if (op == 0) r = orbm::atan2f_glibc(a[i], b[i]);
  else if (op == 1) r = orbm::cosf_glibc(a[i]);
  else if (op == 2) r = orbm::sinf_glibc(a[i]);
  else r = (float)orbm::lround_f(a[i]);
  out[i] = r;
}

}  // namespace orbk
