// orb_lk.cu -- pyramidal Lucas-Kanade tracker: host side of orb_lk_track; kernels in orb_lk_kernels.cuh.
#include <algorithm>

#include "orb_internal.h"
#include "orb_lk_kernels.cuh"

extern "C" {

int orb_lk_levels(int w, int h, int win, int max_level) {
  // buildOpticalFlowPyramid stops before a level that is not larger than the window
  int L = 0;
  while (L < max_level && L + 1 < orbk::LK_MAX_LEVELS) {
    w = (w + 1) / 2; h = (h + 1) / 2;
    if (w <= win || h <= win) break;
    L++;
  }
  return L;
}

int orb_lk_track(orb_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h, size_t pitch, const float* prev_pts, int n,
                 int win, int max_level, int max_iter, double eps, float min_eig, float* next_pts, uint8_t* status, float* err) {
  if (!ctx) return ORB_E_INVALID;
  if (!prev || !next || !prev_pts || !next_pts || !status || n < 0 || w < 1 || h < 1 || pitch < (size_t)w)
    return fail(ctx, ORB_E_INVALID, "bad tracker arguments");
  if (win < 3 || win > orbk::LK_MAX_WIN) return fail(ctx, ORB_E_INVALID, "window %d outside [3, %d]", win, orbk::LK_MAX_WIN);
  if (max_level < 0) return fail(ctx, ORB_E_INVALID, "negative max_level");
  if (n == 0) return ORB_OK;
  CK(cudaSetDevice(ctx->p.device));
  // TermCriteria handling of calcOpticalFlowPyrLK
  max_iter = std::min(std::max(max_iter, 0), 100);
  eps = std::min(std::max(eps, 0.), 10.);
  const double eps2 = eps * eps;
  const int top = orb_lk_levels(w, h, win, max_level);
  size_t lvl_ofs[orbk::LK_MAX_LEVELS + 1];
  int lw[orbk::LK_MAX_LEVELS], lh[orbk::LK_MAX_LEVELS];
  lvl_ofs[0] = 0;
  for (int l = 0, cw = w, ch = h; l <= top; l++, cw = (cw + 1) / 2, ch = (ch + 1) / 2) {
    lw[l] = cw; lh[l] = ch;
    lvl_ofs[l + 1] = lvl_ofs[l] + (((size_t)cw * ch + 255) & ~(size_t)255);
  }
  const size_t pyr = lvl_ofs[top + 1];
  const size_t pts_ofs = 2 * pyr, need = pts_ofs + (size_t)n * (8 + 8 + 4 + 4) + 1024;
  if (need > ctx->d_lk_bytes) {
    if (ctx->d_lk) CK(cudaFree(ctx->d_lk));
    ctx->d_lk = nullptr; ctx->d_lk_bytes = 0;
    CK(cudaMalloc((void**)&ctx->d_lk, need));
    ctx->d_lk_bytes = need;
  }
  uint8_t* dP = ctx->d_lk;
  uint8_t* dN = ctx->d_lk + pyr;
  float* d_prev = reinterpret_cast<float*>(ctx->d_lk + pts_ofs);
  float* d_next = d_prev + 2 * (size_t)n;
  float* d_err = d_next + 2 * (size_t)n;
  uint8_t* d_status = reinterpret_cast<uint8_t*>(d_err + n);
  cudaStream_t q = ctx->stream;
  CK(cudaMemcpy2DAsync(dP, w, prev, pitch, w, h, cudaMemcpyHostToDevice, q));
  CK(cudaMemcpy2DAsync(dN, w, next, pitch, w, h, cudaMemcpyHostToDevice, q));
  CK(cudaMemcpyAsync(d_prev, prev_pts, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, q));
  orbk::LkPyr P, N;
  for (int l = 0; l <= top; l++) {
    P.img[l] = dP + lvl_ofs[l]; N.img[l] = dN + lvl_ofs[l];
    P.w[l] = N.w[l] = lw[l]; P.h[l] = N.h[l] = lh[l];
  }
  P.pitch0 = N.pitch0 = w;
  ctx->lk_top = top; ctx->lk_pyr = pyr;
  for (int l = 0; l <= top; l++) { ctx->lk_w[l] = lw[l]; ctx->lk_h[l] = lh[l]; ctx->lk_ofs[l] = lvl_ofs[l]; }
  ctx->launches = 0;
  for (int l = 1; l <= top; l++) {
    const dim3 blk(32, 8), grd((lw[l] + 31) / 32, (lh[l] + 7) / 8);
    orbk::k_lk_pyrdown<<<grd, blk, 0, q>>>(P.img[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], dP + lvl_ofs[l], lw[l], lh[l], 0, 0);
    orbk::k_lk_pyrdown<<<grd, blk, 0, q>>>(N.img[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], dN + lvl_ofs[l], lw[l], lh[l], 0, 0);
    ctx->launches += 2;
  }
  orbk::k_lk_track<<<(n + orbk::LK_WARPS - 1) / orbk::LK_WARPS, orbk::LK_WARPS * 32, 0, q>>>(
      P, N, top, d_prev, n, win, max_iter, eps2, min_eig, d_next, d_status, err ? d_err : nullptr, 0, 0, 0, nullptr);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(next_pts, d_next, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, q));
  CK(cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, q));
  if (err) CK(cudaMemcpyAsync(err, d_err, sizeof(float) * n, cudaMemcpyDeviceToHost, q));
  CK(cudaStreamSynchronize(q));
  return ORB_OK;
}


// Batch form for a VO loop that keeps its frames on the device: frame t is tracked into frame t + 1 for t = 0 .. n_frames - 2.
// Every frame's pyrDown pyramid is built once (it is the "next" image of one pair and the "previous" image of the following
// one); one launch tracks all points of all pairs (grid.y = pair).
int orb_lk_track_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h, size_t pitch,
                       size_t frame_stride, const float* prev_pts, const int* n_pts, int cap, int pts_on_device, int win, int max_level,
                       int max_iter, double eps, float min_eig, float* next_pts, uint8_t* status, float* err) {
  if (!ctx) return ORB_E_INVALID;
  if (!frames || !prev_pts || !next_pts || !status || n_frames < 0 || cap < 0 || w < 1 || h < 1 || pitch < (size_t)w ||
      frame_stride < pitch * (size_t)h)
    return fail(ctx, ORB_E_INVALID, "bad tracker arguments");
  if (win < 3 || win > orbk::LK_MAX_WIN) return fail(ctx, ORB_E_INVALID, "window %d outside [3, %d]", win, orbk::LK_MAX_WIN);
  if (max_level < 0) return fail(ctx, ORB_E_INVALID, "negative max_level");
  const int npairs = n_frames - 1;
  if (npairs <= 0 || cap == 0) return ORB_OK;
  if (npairs > 65535) return fail(ctx, ORB_E_CAPACITY, "tracker: at most 65536 frames per call");
  CK(cudaSetDevice(ctx->p.device));
  max_iter = std::min(std::max(max_iter, 0), 100);
  eps = std::min(std::max(eps, 0.), 10.);
  const double eps2 = eps * eps;
  const int top = orb_lk_levels(w, h, win, max_level);
  size_t lvl_ofs[orbk::LK_MAX_LEVELS + 1];
  int lw[orbk::LK_MAX_LEVELS], lh[orbk::LK_MAX_LEVELS];
  lvl_ofs[0] = 0;
  for (int l = 0, cw = w, ch = h; l <= top; l++, cw = (cw + 1) / 2, ch = (ch + 1) / 2) {
    lw[l] = cw; lh[l] = ch;
    lvl_ofs[l + 1] = lvl_ofs[l] + (((size_t)cw * ch + 255) & ~(size_t)255);
  }
  // device memory: [levels >= 1 of every frame][level 0 of every frame, packed -- host frames only][staged points -- host only]
  const size_t up = lvl_ofs[top + 1] - lvl_ofs[1], l0 = lvl_ofs[1];
  const size_t per_pair = (size_t)cap * (8 + 8 + 4 + 4);
  const size_t l0_ofs = (size_t)n_frames * up, pts_ofs = l0_ofs + (frames_on_device ? 0 : (size_t)n_frames * l0);
  const size_t need = pts_ofs + (pts_on_device ? 0 : (size_t)npairs * per_pair + (size_t)npairs * 4) + 1024;
  if (need > ctx->d_lk_bytes) {
    if (ctx->d_lk) CK(cudaFree(ctx->d_lk));
    ctx->d_lk = nullptr; ctx->d_lk_bytes = 0;
    CK(cudaMalloc((void**)&ctx->d_lk, need));
    ctx->d_lk_bytes = need;
  }
  cudaStream_t q = ctx->stream;
  uint8_t* dU = ctx->d_lk;
  // level 0: the caller's frames where they are (device), or a packed copy (host)
  const uint8_t* L0 = frames;
  size_t stride0 = frame_stride;
  int pitch0 = (int)pitch;
  if (!frames_on_device) {
    uint8_t* d0 = ctx->d_lk + l0_ofs;
    if (frame_stride == pitch * (size_t)h && l0 == (size_t)w * h) {
      CK(cudaMemcpy2DAsync(d0, w, frames, pitch, w, (size_t)h * n_frames, cudaMemcpyHostToDevice, q));
    } else {
      for (int f = 0; f < n_frames; f++)
        CK(cudaMemcpy2DAsync(d0 + (size_t)f * l0, w, frames + (size_t)f * frame_stride, pitch, w, h, cudaMemcpyHostToDevice, q));
    }
    L0 = d0; stride0 = l0; pitch0 = w;
  }
  const float* d_prev = prev_pts; float* d_next = next_pts; uint8_t* d_status = status; float* d_err = err; const int* d_n = n_pts;
  if (!pts_on_device) {
    float* a = reinterpret_cast<float*>(ctx->d_lk + pts_ofs);
    float* b = a + 2 * (size_t)npairs * cap;
    float* c = b + 2 * (size_t)npairs * cap;
    int* dn = reinterpret_cast<int*>(c + (size_t)npairs * cap);
    uint8_t* e = reinterpret_cast<uint8_t*>(dn + npairs);
    CK(cudaMemsetAsync(b, 0, (size_t)npairs * per_pair - sizeof(float) * 2 * (size_t)npairs * cap, q));   // entries beyond a pair's count read as 0
    CK(cudaMemcpyAsync(a, prev_pts, sizeof(float) * 2 * (size_t)npairs * cap, cudaMemcpyHostToDevice, q));
    if (n_pts) CK(cudaMemcpyAsync(dn, n_pts, sizeof(int) * npairs, cudaMemcpyHostToDevice, q));
    d_prev = a; d_next = b; d_err = err ? c : nullptr; d_status = e; d_n = n_pts ? dn : nullptr;
  }
  orbk::LkPyr P, N;
  P.img[0] = L0; N.img[0] = L0 + stride0;
  for (int l = 1; l <= top; l++) { P.img[l] = dU + (lvl_ofs[l] - lvl_ofs[1]); N.img[l] = P.img[l] + up; }
  for (int l = 0; l <= top; l++) { P.w[l] = N.w[l] = lw[l]; P.h[l] = N.h[l] = lh[l]; }
  P.pitch0 = N.pitch0 = pitch0;
  ctx->lk_top = -1;                                            // orb_lk_get_level serves the single-pair call only
  ctx->launches = 0;
  for (int l = 1; l <= top; l++) {
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {            // grid.z limit
      const int nf = std::min(65535, n_frames - f0);
      const dim3 blk(32, 8), grd((lw[l] + 31) / 32, (lh[l] + 7) / 8, nf);
      const uint8_t* src = l == 1 ? L0 + (size_t)f0 * stride0 : P.img[l - 1] + (size_t)f0 * up;
      orbk::k_lk_pyrdown<<<grd, blk, 0, q>>>(src, lw[l - 1], lh[l - 1], l == 1 ? pitch0 : lw[l - 1], dU + (size_t)f0 * up + (lvl_ofs[l] - lvl_ofs[1]),
                                             lw[l], lh[l], l == 1 ? stride0 : up, up);
      ctx->launches += 1;
    }
  }
  orbk::k_lk_track<<<dim3((cap + orbk::LK_WARPS - 1) / orbk::LK_WARPS, npairs), orbk::LK_WARPS * 32, 0, q>>>(
      P, N, top, d_prev, cap, win, max_iter, eps2, min_eig, d_next, d_status, d_err, stride0, up, (size_t)cap, d_n);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (!pts_on_device) {
    CK(cudaMemcpyAsync(next_pts, d_next, sizeof(float) * 2 * (size_t)npairs * cap, cudaMemcpyDeviceToHost, q));
    CK(cudaMemcpyAsync(status, d_status, (size_t)npairs * cap, cudaMemcpyDeviceToHost, q));
    if (err) CK(cudaMemcpyAsync(err, d_err, sizeof(float) * (size_t)npairs * cap, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
  }
  return ORB_OK;
}

int orb_lk_get_level(orb_ctx* ctx, int which, int level, uint8_t* dst, int* w, int* h) {
  if (!ctx || !dst || !ctx->d_lk || ctx->lk_top < 0) return ORB_E_INVALID;
  if (level < 0 || level > ctx->lk_top || which < 0 || which > 1) return fail(ctx, ORB_E_INVALID, "no such pyramid level");
  CK(cudaSetDevice(ctx->p.device));
  CK(cudaMemcpy(dst, ctx->d_lk + (which ? ctx->lk_pyr : 0) + ctx->lk_ofs[level], (size_t)ctx->lk_w[level] * ctx->lk_h[level],
                cudaMemcpyDeviceToHost));
  if (w) *w = ctx->lk_w[level];
  if (h) *h = ctx->lk_h[level];
  return ORB_OK;
}


}  // extern "C"
