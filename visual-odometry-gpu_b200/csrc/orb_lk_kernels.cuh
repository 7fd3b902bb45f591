// orb_lk_kernels.cuh -- pyramidal Lucas-Kanade tracker: the device counterpart of
//   cv::calcOpticalFlowPyrLK(img_1, img_2, points1, points2, status, err, Size(21,21), 3,
//                            TermCriteria(COUNT+EPS, 30, 0.01), 0, 0.001)          (reference src/feature_tracking.cpp:174-180)
// following OpenCV 4.x modules/video/src/lkpyramid.cpp (the third-party code behind that call) step for step:
//   k_lk_pyrdown : cv::pyrDown -- 5x5 [1 4 6 4 1]^2, REFLECT_101, (s + 128) >> 8, size (w+1)/2 x (h+1)/2
//   k_lk_track   : one warp per point, all pyramid levels in one launch.  Per level: the 14-bit fixed-point bilinear window
//                  of the previous frame (intensity with 5 fractional bits) and of its Scharr derivatives (computed on the fly
//                  from a 4x4 neighbourhood, zero outside the image like OpenCV's BORDER_CONSTANT derivative border) goes to
//                  shared memory, the 2x2 normal matrix is summed in float, then up to max_iter Newton steps sample the next
//                  frame.  Float sums run as 32 interleaved partial sums (lane = element % 32) + xor butterfly: the order the
//                  CPU checker uses, so results are bit-identical to it; every float step is a single IEEE operation (no FMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "orb_internal.h"

namespace orbk {

// LK_MAX_LEVELS (8) is defined in orb_internal.h: the context keeps the level layout of the last call
constexpr int LK_MAX_WIN = 33;
constexpr int LK_WARPS = 4;

struct LkPyr {
  const uint8_t* img[LK_MAX_LEVELS];   // levels >= 1 are packed (pitch = width); level 0 may be the caller's frame with its own pitch
  int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
  int pitch0;                          // row pitch of level 0
};

__device__ __forceinline__ int lk_reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

// blockIdx.z = frame of a batch (src_stride / dst_stride bytes between the levels of consecutive frames; one frame: 0)
__global__ void k_lk_pyrdown(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw, int dh,
                             size_t src_stride, size_t dst_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  src += (size_t)blockIdx.z * src_stride;
  dst += (size_t)blockIdx.z * dst_stride;
  int s = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const uint8_t* row = src + (size_t)lk_reflect101(2 * y + j - 2, sh) * spitch;
    const int r = row[lk_reflect101(2 * x - 2, sw)] + 4 * row[lk_reflect101(2 * x - 1, sw)] + 6 * row[lk_reflect101(2 * x, sw)] +
                  4 * row[lk_reflect101(2 * x + 1, sw)] + row[lk_reflect101(2 * x + 2, sw)];
    s += (j == 0 || j == 4 ? 1 : (j == 2 ? 6 : 4)) * r;
  }
  dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
}

__device__ __forceinline__ int lk_descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

__device__ __forceinline__ float lk_warp_sum(float v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

struct LkWeights { int w00, w01, w10, w11; };
__device__ __forceinline__ LkWeights lk_weights(float a, float b) {
  // cvRound((1.f - a)*(1.f - b)*(1 << W_BITS)) ...: float products, round half to even
  LkWeights W;
  const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
  W.w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), 16384.f));
  W.w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), 16384.f));
  W.w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), 16384.f));
  W.w11 = 16384 - W.w00 - W.w01 - W.w10;
  return W;
}

// bilinear sample of the image (5 fractional bits) at integer corner (X, Y)
__device__ __forceinline__ int lk_sample(const uint8_t* __restrict__ img, int w, int h, int pitch, int X, int Y, const LkWeights& W, bool inside) {
  int p00, p01, p10, p11;
  if (inside) {
    const uint8_t* r = img + (size_t)Y * pitch + X;
    p00 = r[0]; p01 = r[1]; p10 = r[pitch]; p11 = r[pitch + 1];
  } else {
    const int x0 = lk_reflect101(X, w), x1 = lk_reflect101(X + 1, w);
    const uint8_t* r0 = img + (size_t)lk_reflect101(Y, h) * pitch;
    const uint8_t* r1 = img + (size_t)lk_reflect101(Y + 1, h) * pitch;
    p00 = r0[x0]; p01 = r0[x1]; p10 = r1[x0]; p11 = r1[x1];
  }
  return lk_descale(p00 * W.w00 + p01 * W.w01 + p10 * W.w10 + p11 * W.w11, 14 - 5);
}

// blockIdx.y = frame pair of a batch: its pyramids lie pair * pair_stride bytes (level 0: pair * pair_stride0) behind P / N, its points pair * pts_stride
// entries behind prev_pts / next_pts / status / err, and it tracks n_arr[pair] (clipped to n) points; one pair: strides 0,
// n_arr NULL.
__global__ void __launch_bounds__(LK_WARPS * 32) k_lk_track(const LkPyr P, const LkPyr N, int top, const float* __restrict__ prev_pts, int n,
                                                            int win, int max_iter, double eps2, float min_eig,
                                                            float* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                            float* __restrict__ err, size_t pair_stride0, size_t pair_stride, size_t pts_stride,
                                                            const int* __restrict__ n_arr) {
  __shared__ short s_I[LK_WARPS][LK_MAX_WIN * LK_MAX_WIN];
  __shared__ short s_dI[LK_WARPS][2 * LK_MAX_WIN * LK_MAX_WIN];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pt = blockIdx.x * LK_WARPS + wid;
  if (n_arr) n = max(0, min(n, n_arr[blockIdx.y]));
  if (pt >= n) return;
  prev_pts += 2 * (size_t)blockIdx.y * pts_stride; next_pts += 2 * (size_t)blockIdx.y * pts_stride;
  status += (size_t)blockIdx.y * pts_stride;
  if (err) err += (size_t)blockIdx.y * pts_stride;
  short* Iw = s_I[wid];
  short* dIw = s_dI[wid];
  const int area = win * win;
  const float halfWin = __fmul_rn((float)(win - 1), 0.5f);
  const float FLT_SCALE = 1.f / (1 << 20);
  const float p0x = prev_pts[2 * pt], p0y = prev_pts[2 * pt + 1];
  float outx = 0.f, outy = 0.f, errv = 0.f;     // nextPts[ptidx], err[ptidx]
  int st = 1;
  // lane's first element and its (x, y) inside the window
  const int ex0 = lane % win, ey0 = lane / win;

  for (int level = top; level >= 0; level--) {
    const size_t img_ofs = (size_t)blockIdx.y * (level == 0 ? pair_stride0 : pair_stride);
    const uint8_t* I = P.img[level] + img_ofs;
    const uint8_t* J = N.img[level] + img_ofs;
    const int w = P.w[level], h = P.h[level], pitch = level == 0 ? P.pitch0 : w;
    const float sc = (float)(1. / (double)(1 << level));
    float px = __fmul_rn(p0x, sc), py = __fmul_rn(p0y, sc);
    float nx, ny;
    if (level == top) { nx = px; ny = py; } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
    outx = nx; outy = ny;
    px = __fsub_rn(px, halfWin); py = __fsub_rn(py, halfWin);
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
      if (level == 0) { st = 0; errv = 0.f; }
      continue;
    }
    LkWeights W = lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy));
    float a11 = 0.f, a12 = 0.f, a22 = 0.f;
    __syncwarp();
    {
      // the 4x4 neighbourhood of every window element covers the four 3x3 Scharr stencils of its bilinear corners
      const bool inside = ipx >= 1 && ipy >= 1 && ipx + win + 2 <= w - 1 && ipy + win + 2 <= h - 1;
      int x = ex0, y = ey0;
      for (int e = lane; e < area; e += 32) {
        const int X = ipx + x, Y = ipy + y;
        int p[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const uint8_t* row = I + (size_t)(inside ? Y + r - 1 : lk_reflect101(Y + r - 1, h)) * pitch;
#pragma unroll
          for (int c = 0; c < 4; c++) p[r][c] = row[inside ? X + c - 1 : lk_reflect101(X + c - 1, w)];
        }
        const int ival = lk_descale(p[1][1] * W.w00 + p[1][2] * W.w01 + p[2][1] * W.w10 + p[2][2] * W.w11, 14 - 5);
        // Scharr at the four corners (cx, cy) in {0,1}^2: smoothing (3, 10, 3) across, difference along
        int dx[2][2], dy[2][2];
#pragma unroll
        for (int cy = 0; cy < 2; cy++) {
          int t0[4], t1[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            t0[c] = (p[cy][c] + p[cy + 2][c]) * 3 + p[cy + 1][c] * 10;
            t1[c] = p[cy + 2][c] - p[cy][c];
          }
#pragma unroll
          for (int cx = 0; cx < 2; cx++) {
            const bool in_img = (unsigned)(X + cx) < (unsigned)w && (unsigned)(Y + cy) < (unsigned)h;
            dx[cy][cx] = in_img ? t0[cx + 2] - t0[cx] : 0;
            dy[cy][cx] = in_img ? (t1[cx + 2] + t1[cx]) * 3 + t1[cx + 1] * 10 : 0;
          }
        }
        const int ixval = lk_descale(dx[0][0] * W.w00 + dx[0][1] * W.w01 + dx[1][0] * W.w10 + dx[1][1] * W.w11, 14);
        const int iyval = lk_descale(dy[0][0] * W.w00 + dy[0][1] * W.w01 + dy[1][0] * W.w10 + dy[1][1] * W.w11, 14);
        Iw[e] = (short)ival; dIw[2 * e] = (short)ixval; dIw[2 * e + 1] = (short)iyval;
        a11 = __fadd_rn(a11, (float)(ixval * ixval));
        a12 = __fadd_rn(a12, (float)(ixval * iyval));
        a22 = __fadd_rn(a22, (float)(iyval * iyval));
        x += 32;
        while (x >= win) { x -= win; y++; }
      }
    }
    __syncwarp();
    const float A11 = __fmul_rn(lk_warp_sum(a11), FLT_SCALE), A12 = __fmul_rn(lk_warp_sum(a12), FLT_SCALE),
                A22 = __fmul_rn(lk_warp_sum(a22), FLT_SCALE);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                                   (float)(2 * win * win));
    if (minEig < min_eig || D < 1.1920928955078125e-07f) {
      if (level == 0) st = 0;
      continue;
    }
    D = __fdiv_rn(1.f, D);
    nx = __fsub_rn(nx, halfWin); ny = __fsub_rn(ny, halfWin);
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_iter; j++) {
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -win || inx >= w || iny < -win || iny >= h) {
        if (level == 0) st = 0;
        break;
      }
      W = lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny));
      const bool inside = inx >= 0 && iny >= 0 && inx + win + 1 <= w - 1 && iny + win + 1 <= h - 1;
      float b1 = 0.f, b2 = 0.f;
      int x = ex0, y = ey0;
      for (int e = lane; e < area; e += 32) {
        const int diff = lk_sample(J, w, h, pitch, inx + x, iny + y, W, inside) - Iw[e];
        b1 = __fadd_rn(b1, (float)(diff * dIw[2 * e]));
        b2 = __fadd_rn(b2, (float)(diff * dIw[2 * e + 1]));
        x += 32;
        while (x >= win) { x -= win; y++; }
      }
      b1 = __fmul_rn(lk_warp_sum(b1), FLT_SCALE);
      b2 = __fmul_rn(lk_warp_sum(b2), FLT_SCALE);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
      outx = __fadd_rn(nx, halfWin); outy = __fadd_rn(ny, halfWin);
      if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2) break;
      if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
        outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
        outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
        break;
      }
      pdx = dx; pdy = dy;
    }
    if (st && level == 0) {
      const float ex = __fsub_rn(outx, halfWin), ey = __fsub_rn(outy, halfWin);
      const int inx = (int)floorf(ex), iny = (int)floorf(ey);
      if (inx < -win || inx >= w || iny < -win || iny >= h) { st = 0; continue; }
      W = lk_weights(__fsub_rn(ex, (float)inx), __fsub_rn(ey, (float)iny));
      const bool inside = inx >= 0 && iny >= 0 && inx + win + 1 <= w - 1 && iny + win + 1 <= h - 1;
      float ev = 0.f;
      int x = ex0, y = ey0;
      for (int e = lane; e < area; e += 32) {
        const int diff = lk_sample(J, w, h, pitch, inx + x, iny + y, W, inside) - Iw[e];
        ev = __fadd_rn(ev, fabsf((float)diff));
        x += 32;
        while (x >= win) { x -= win; y++; }
      }
      errv = __fdiv_rn(__fmul_rn(lk_warp_sum(ev), 1.f), (float)(32 * win * win));
    }
  }
  if (lane == 0) {
    next_pts[2 * pt] = outx; next_pts[2 * pt + 1] = outy;
    status[pt] = (uint8_t)st;
    if (err) err[pt] = errv;
  }
}

}  // namespace orbk
