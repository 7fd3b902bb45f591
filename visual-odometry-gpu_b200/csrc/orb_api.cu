// orb_api.cu -- host side of the C ABI declared in include/orb_b200.h.
// Owns the device arena, the per-shape plan (level geometry, quotas, resize tap tables) and the
// launch sequence.  No CPU compute path exists here: every result comes from the kernels in
// orb_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/orb_b200.h"
#include "../../include/orb_brief_pattern.h"
#include "orb_internal.h"
#include "orb_kernels.cuh"
#include "orb_match_tc.cuh"

#ifndef ORB_E2E_RAMP_DOWN
#define ORB_E2E_RAMP_DOWN 1
#endif

using orbk::Bufs;
using orbk::DescribeJob;


// == extern int bit_pattern_31_[256*4] of the reference (include/orb_pattern.hpp:2)
extern "C" { int bit_pattern_31_[256 * 4]; }
namespace {
struct PatternInit {
  PatternInit() { for (int i = 0; i < 1024; i++) bit_pattern_31_[i] = ORB_BRIEF_PATTERN_31[i]; }
} g_pattern_init;
}  // namespace
thread_local char g_orb_create_error[512] = "";   // errors without a context (orb_create, PNG helpers): per thread
#define g_create_error g_orb_create_error

namespace {

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// float scale = pow(scaleFactor, i);   ref src/orb_cpu.cpp:284 / src/orb.cpp:95
float level_scale(float f, int l) { return (float)std::pow((double)f, (double)l); }
void level_size(int W, int H, float f, int l, int* w, int* h) {
  if (l == 0) { *w = W; *h = H; return; }
  float s = level_scale(f, l);
  *w = (int)std::round(W / s);   // cv::Size(round(W / scale), round(H / scale)), ref src/orb_cpu.cpp:285
  *h = (int)std::round(H / s);
}
// int nfeatures_l = nfeatures * ((1 - 1/f) / (1 - pow(1/f, L))) * pow(1/f, l);   ref src/orb.cpp:62
int level_quota(int nfeatures, float f, int L, int l) {
  float inv = 1 / f;
  double a = (1 - inv) / (1 - std::pow((double)inv, (double)L));
  return (int)(nfeatures * a * std::pow((double)inv, (double)l));
}

// cv::resize(INTER_LINEAR) tap table for one axis (OpenCV's float recipe, 11-bit coefficients)
void make_taps(int src, int dst, OrbTap* t) {
  double scale = 1.0 / ((double)dst / src);
  for (int d = 0; d < dst; d++) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(fx);
    fx -= s;
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= src - 1) { s = src - 1; fx = 0.f; }
    t[d].s0 = (uint16_t)s;
    t[d].s1 = (uint16_t)std::min(s + 1, src - 1);
    t[d].a0 = (int16_t)std::lrintf((1.f - fx) * 2048.f);
    t[d].a1 = (int16_t)std::lrintf(fx * 2048.f);
  }
}

// createGaussianKernel(7), ref src/GaussianBlur.cpp:7-37 (float arithmetic, sigma heuristic)
void harris_weights(float* k) {
  const int ks = 7, half = 3;
  float sigma = 0.3f * ((ks - 1) * 0.5f) + 0.8f, sum = 0.0f;
  for (int y = -half; y <= half; ++y)
    for (int x = -half; x <= half; ++x) {
      float v = std::exp(-(x * x + y * y) / (2 * sigma * sigma));
      k[(y + half) * ks + (x + half)] = v;
      sum += v;
    }
  for (int i = 0; i < ks * ks; ++i) k[i] /= sum;
}

// geometry of one frame shape; nlevels/policy/quota may be overridden for the single-image stages
void build_plan(const orb_params& p, int W, int H, int nlevels, int policy, int quota_override, OrbPlan* P) {
  memset(P, 0, sizeof(*P));
  P->nlevels = nlevels; P->W = W; P->H = H;
  P->fast_threshold = p.fast_threshold; P->fast_n = p.fast_n;
  P->nms_radius = p.nms_window / 2; P->patch_radius = p.orient_patch / 2;
  P->select_policy = policy; P->blur_levels = p.blur_levels; P->harris_k = p.harris_k;
  int tile = 0, atile = 0, kept = 0, xo = 0, yo = 0, eo = 0, e2 = 0;
  unsigned long long lv = 0, bx = 0, cd = 0;
  for (int l = 0; l < nlevels; l++) {
    OrbLevel& G = P->lv[l];
    level_size(W, H, p.scale_factor, l, &G.w, &G.h);
    G.w = std::max(G.w, 1); G.h = std::max(G.h, 1);
    G.pitch = align_up(G.w, 16);
    G.bpitch = align_up(G.w + 1, 8);
    G.tiles_x = (G.w + orbk::B_TW - 1) / orbk::B_TW;
    G.tiles_y = (G.h + orbk::B_TH - 1) / orbk::B_TH;
    G.tile_ofs = tile; tile += G.tiles_x * G.tiles_y;
    G.a_tiles_x = (G.w + orbk::A_TW - 1) / orbk::A_TW;
    G.a_tiles_y = (G.h + orbk::A_TH - 1) / orbk::A_TH;
    G.a_tile_ofs = atile; if (l > 0) atile += G.a_tiles_x * G.a_tiles_y;
    int q = quota_override >= 0 ? quota_override
            : (policy == ORB_SELECT_HARRIS_TOP_N ? level_quota(p.nfeatures, p.scale_factor, p.nlevels, l) : p.nfeatures);
    G.quota = std::max(0, std::min(q, ORB_SORT_CAP));
    G.cand_cap = std::max(2048, G.w * G.h / 8);
    G.kept_ofs = kept; kept += align_up(std::max(G.quota, 1), 4);
    G.xtab_ofs = xo; G.ytab_ofs = yo; xo += G.w; yo += G.h;
    G.edge_ofs = eo; G.edge_w = align_up(G.w, 4); eo += G.edge_w + align_up(G.h, 4);
    G.edge2_ofs = e2; e2 += orbk::edge2_level_elems(G.edge_w, G.h);
    G.scale = level_scale(p.scale_factor, l);
    G.lvl_ofs = lv; if (l > 0) lv += (unsigned long long)align_up(G.h * G.pitch, 256);
    G.box_ofs = bx; bx += (unsigned long long)align_up((G.h + 1) * G.bpitch, 128);
    G.cand_ofs = cd; cd += (unsigned long long)align_up(G.cand_cap, 32);
  }
  P->tiles_per_frame = tile; P->a_tiles_per_frame = atile; P->kept_per_frame = kept; P->edge_frame_elems = eo; P->edge2_frame_elems = e2;
  P->pyr_frame_bytes = std::max<unsigned long long>(lv, 256); P->box_frame_elems = bx; P->cand_frame_elems = cd;
}

// flattened tile lists of the two tiled kernels: level | tile_x << 4 | tile_y << 18
void make_tile_tables(const OrbPlan& P, std::vector<uint32_t>* ta, std::vector<uint32_t>* tb) {
  ta->clear(); tb->clear();
  for (int l = 0; l < P.nlevels; l++) {
    const OrbLevel& G = P.lv[l];
    if (l > 0)
      for (int ty = 0; ty < G.a_tiles_y; ty++)
        for (int tx = 0; tx < G.a_tiles_x; tx++) ta->push_back((uint32_t)l | ((uint32_t)tx << 4) | ((uint32_t)ty << 18));
    for (int ty = 0; ty < G.tiles_y; ty++)
      for (int tx = 0; tx < G.tiles_x; tx++) tb->push_back((uint32_t)l | ((uint32_t)tx << 4) | ((uint32_t)ty << 18));
  }
}

int upload_tables(orb_ctx* ctx, const OrbPlan& P) {
  std::vector<OrbTap> xt, yt;
  for (int l = 0; l < P.nlevels; l++) {
    const OrbLevel& G = P.lv[l];
    xt.resize(G.xtab_ofs + G.w); yt.resize(G.ytab_ofs + G.h);
    make_taps(P.W, G.w, xt.data() + G.xtab_ofs);
    make_taps(P.H, G.h, yt.data() + G.ytab_ofs);
  }
  if ((int)xt.size() > ctx->xtab_cap || (int)yt.size() > ctx->ytab_cap) return fail(ctx, ORB_E_CAPACITY, "tap tables exceed arena");
  std::vector<uint32_t> ta, tb;
  make_tile_tables(P, &ta, &tb);
  if ((int)ta.size() > ctx->tile_a_cap || (int)tb.size() > ctx->tile_b_cap) return fail(ctx, ORB_E_CAPACITY, "tile tables exceed arena");
  if (!ta.empty()) CK(cudaMemcpyAsync(ctx->d_tile_a, ta.data(), ta.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_tile_b, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_xtab, xt.data(), xt.size() * sizeof(OrbTap), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_ytab, yt.data(), yt.size() * sizeof(OrbTap), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // host vectors die here
  return ORB_OK;
}

int get_plan(orb_ctx* ctx, int w, int h) {
  if (w < 1 || h < 1) return fail(ctx, ORB_E_INVALID, "bad image size %dx%d", w, h);
  if (w > ctx->p.max_width || h > ctx->p.max_height)
    return fail(ctx, ORB_E_CAPACITY, "image %dx%d exceeds context capacity %dx%d", w, h, ctx->p.max_width, ctx->p.max_height);
  if (ctx->plan_valid && ctx->plan.W == w && ctx->plan.H == h) return ORB_OK;
  build_plan(ctx->p, w, h, ctx->p.nlevels, ctx->p.select_policy, -1, &ctx->plan);
  ctx->plan_valid = false;
  int rc = upload_tables(ctx, ctx->plan);
  if (rc) return rc;
  ctx->plan_valid = true;
  return ORB_OK;
}

void fill_bufs(orb_ctx* ctx, Bufs* B) {
  memset(B, 0, sizeof(*B));
  B->pyr = ctx->d_pyr; B->box = ctx->d_box; B->cand = ctx->d_cand; B->cand_count = ctx->d_cand_count;
  B->zero_stride = (int)(ctx->zero_bytes_per_frame / sizeof(int)); B->kept_xy = ctx->d_kept_xy; B->kept_r = ctx->d_kept_r; B->kept_count = ctx->d_kept_count;
  B->xtab = ctx->d_xtab; B->ytab = ctx->d_ytab; B->tile_a = ctx->d_tile_a; B->tile_b = ctx->d_tile_b; B->pattern = ctx->d_pattern;
  B->flags = ctx->d_flags;
  B->tmaps = ctx->d_tmaps; B->frame0 = 0; B->edge2 = ctx->d_edge2;
}

// ---- TMA tensor maps -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D map (x, y, frame) over `n` images of w x h elements; out-of-range elements read as zero
int encode_map(orb_ctx* ctx, CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, int w, int h, int n,
               size_t pitch_bytes, size_t frame_bytes, int box_w, int box_h, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE) {
  if ((uintptr_t)base % 16 || pitch_bytes % 16 || frame_bytes % 16 || (size_t)box_w * esize % 16 || box_w > 256 || box_h > 256)
    return fail(ctx, ORB_E_INVALID, "tensor map geometry not TMA-compatible (base %p pitch %zu stride %zu box %dx%d)", base, pitch_bytes,
                frame_bytes, box_w, box_h);
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)std::max(n, 1)};
  cuuint64_t strides[2] = {(cuuint64_t)pitch_bytes, (cuuint64_t)frame_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)ctx->tmap_encode)(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, ORB_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return ORB_OK;
}

// (re-)encode the maps of plan P for level-0 frames at `base` (n frames, row pitch / frame stride in bytes); the arena
// levels and box-sum images are fixed per plan.  Stream-ordered upload: earlier launches keep the maps they were given.
int update_tmaps(orb_ctx* ctx, const OrbPlan& P, const uint8_t* base, size_t pitch, size_t stride, int n) {
  const orb_ctx::TmapKey key{base, pitch, stride, n, P.W, P.H, P.nlevels, P.patch_radius};
  if (!memcmp(&key, &ctx->tmap_key, sizeof(key))) return ORB_OK;
  memset(&ctx->tmap_key, 0, sizeof(ctx->tmap_key));
  int rc;
  for (int l = 0; l < P.nlevels; l++) {
    const OrbLevel& G = P.lv[l];
    const uint8_t* img = l == 0 ? base : ctx->d_pyr + G.lvl_ofs;
    const size_t ip = l == 0 ? pitch : (size_t)G.pitch, is = l == 0 ? stride : (size_t)P.pyr_frame_bytes;
    const int in = l == 0 ? n : ctx->chunk;
    if ((rc = encode_map(ctx, &ctx->h_tmaps[orbk::TM_PIX + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, img, G.w, G.h, in, ip, is, orbk::B_SP,
                         orbk::B_PH)))
      return rc;
  }
  CK(cudaMemcpyAsync(ctx->d_tmaps, ctx->h_tmaps, sizeof(ctx->h_tmaps), cudaMemcpyHostToDevice, ctx->stream));
  ctx->tmap_key = key;
  return ORB_OK;
}

// event bracket around one kernel launch (only when profiling is on)
struct StageTimer {
  orb_ctx* c; orb_ctx::Span* sp = nullptr;
  StageTimer(orb_ctx* ctx, int stage) : c(ctx) {
    if (!c->profiling) return;
    if (c->spans_used == c->spans.size()) {
      orb_ctx::Span n{stage, nullptr, nullptr};
      if (cudaEventCreate(&n.a) != cudaSuccess || cudaEventCreate(&n.b) != cudaSuccess) return;
      c->spans.push_back(n);
    }
    sp = &c->spans[c->spans_used++];
    sp->stage = stage;
    cudaEventRecord(sp->a, c->stream);
  }
  ~StageTimer() { if (sp) cudaEventRecord(sp->b, c->stream); }
};

int launch_pyramid_fast(orb_ctx* ctx, const OrbPlan& P, const Bufs& B, int nframes) {
  if (ctx->edges_pending) {      // a k_edges of an earlier call (one that did not describe) may still read the tables
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    ctx->edges_pending = false;
  }
  // candidate counters and BRIEF border tables of the wave's frames (one small memset)
  CK(cudaMemsetAsync(ctx->d_cand_count, 0, ctx->zero_bytes_per_frame * nframes, ctx->stream));
  if (P.a_tiles_per_frame > 0) {
    StageTimer t(ctx, 0);
    orbk::k_pyramid<<<dim3(P.a_tiles_per_frame, nframes), orbk::A_THREADS, 0, ctx->stream>>>(P, B);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  {
    StageTimer t(ctx, 1);
    orbk::k_fast<<<dim3(P.tiles_per_frame, nframes), orbk::B_THREADS, orbk::B_SMEM, ctx->stream>>>(P, B);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  {
    // border-box tables for BRIEF from the strip tables k_fast has just finished (~2 (W + H) entries per level: a small,
    // latency-bound kernel).  Nothing before k_describe needs them, so they are made on a side stream next to k_harris /
    // k_select; launch_describe() joins.
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_side, ctx->ev_fork, 0));
    orbk::k_edges<<<dim3(P.nlevels, nframes), orbk::E_THREADS, 0, ctx->s_side>>>(P, B);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_join, ctx->s_side));
    ctx->edges_pending = true;
    ctx->launches += 1;
  }
  if (P.select_policy == ORB_SELECT_HARRIS_TOP_N) {
    // grid sized for ~1.5 % of the pyramid pixels surviving NMS; denser frames take more grid-stride rounds
    int total_px = 0;
    for (int l = 0; l < P.nlevels; l++) total_px += P.lv[l].w * P.lv[l].h;
    const int blocks = std::max(1, std::min(total_px / 128 / orbk::C_THREADS + 1, 4096));
    StageTimer t(ctx, 2);
    orbk::k_harris<<<dim3(blocks, nframes), orbk::C_THREADS, 0, ctx->stream>>>(P, B);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return ORB_OK;
}
int launch_select(orb_ctx* ctx, const OrbPlan& P, const Bufs& B, int nframes) {
  int mq = 1;
  for (int l = 0; l < P.nlevels; l++) mq = std::max(mq, P.lv[l].quota);
  int npow2 = 1;
  while (npow2 < mq) npow2 <<= 1;
  dim3 grid(P.nlevels, nframes);
  {
    StageTimer t(ctx, 3);
    const size_t rows = (size_t)std::min(P.H, orbk::K2_MAX_ROWS) + 1;        // counting sort over rows: [npow2] row groups + row counters
    orbk::k_select<<<grid, orbk::K2_THREADS, (size_t)(2 * npow2 + orbk::K2_SMEM_KEYS) * 8 + rows * 4, ctx->stream>>>(P, B, npow2);
  }
  CK(cudaGetLastError());
  ctx->launches += 1;
  return ORB_OK;
}
int launch_describe(orb_ctx* ctx, const OrbPlan& P, const Bufs& B, const DescribeJob& J, int nwarps, int nframes) {
  if (ctx->edges_pending) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    ctx->edges_pending = false;
  }
  if (nwarps <= 0) return ORB_OK;
  dim3 grid((nwarps + orbk::K3_KPS - 1) / orbk::K3_KPS, nframes);
  {
    StageTimer t(ctx, 4);
    orbk::k_describe<<<grid, orbk::K3_WARPS * 32, 0, ctx->stream>>>(P, B, J);
  }
  CK(cudaGetLastError());
  ctx->launches += 1;
  return ORB_OK;
}

// Wave boundaries of a batch of n_frames: begins[i] .. begins[i + 1] is wave i (begins ends with n_frames).
// ramp_up (frames staged from the host): the first waves are wave/8, wave/4, wave/2 frames, so that the kernels start while
// most of the batch is still in flight over PCIe.  ramp_down (results go back to the host): the last waves shrink again
// (wave/2, wave/4, the rest), so that what is left after the last frame has arrived is the computation and the result
// copy of a short wave only; batches shorter than four waves keep full waves to the end.
void wave_schedule(int n_frames, int wave, bool ramp_up, bool ramp_down, std::vector<int>& begins) {
  begins.clear();
  wave = std::max(1, wave);
  int c0 = 0, ramp = (ramp_up && n_frames > wave) ? std::max(1, wave / 8) : wave;
  int tail = 0;
  if (ramp_down && n_frames >= 4 * wave) tail = wave / 2 + wave / 4 + wave / 8;
  while (c0 < n_frames - tail) {
    begins.push_back(c0);
    c0 += std::min(std::min(ramp, wave), n_frames - tail - c0);
    if (ramp < wave) ramp *= 2;
  }
  for (int d = wave / 2; tail > 0 && c0 < n_frames; d = std::max(1, d / 2)) {
    begins.push_back(c0);
    c0 += std::min(d <= wave / 8 ? n_frames - c0 : d, n_frames - c0);
  }
  begins.push_back(std::max(n_frames, 0));
}

int check_flags(orb_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (*ctx->h_flags & 1) {
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    return fail(ctx, ORB_E_OVERFLOW, "corner candidates exceeded the per-level arena (w*h/8 slots)");
  }
  return ORB_OK;
}

// copy one host image into frame slot 0 of the staging area (single-image stage entry points)
int stage_image(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch) {
  if (!img || w < 1 || h < 1 || pitch < (size_t)w) return fail(ctx, ORB_E_INVALID, "bad image argument");
  if (w > ctx->p.max_width || h > ctx->p.max_height)
    return fail(ctx, ORB_E_CAPACITY, "image %dx%d exceeds context capacity %dx%d", w, h, ctx->p.max_width, ctx->p.max_height);
  CK(cudaMemcpy2DAsync(ctx->d_frames, ctx->frames_pitch, img, pitch, w, h, cudaMemcpyHostToDevice, ctx->stream));
  ctx->last_n = 0;   // stage calls reuse slot 0 of the arena
  return ORB_OK;
}

int stage_plan(orb_ctx* ctx, int w, int h, int policy, int quota, OrbPlan* P, Bufs* B) {
  build_plan(ctx->p, w, h, 1, policy, quota, P);
  P->lv[0].cand_cap = ctx->max_plan.lv[0].cand_cap;
  fill_bufs(ctx, B);
  {
    std::vector<uint32_t> ta, tb;
    make_tile_tables(*P, &ta, &tb);
    cudaMemcpyAsync(ctx->d_tile_b1, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaStreamSynchronize(ctx->stream);   // tb dies here
    B->tile_b = ctx->d_tile_b1;
  }
  B->frames = ctx->d_frames; B->frame_stride = ctx->frames_slot_bytes; B->pitch0 = ctx->frames_pitch;
  B->frames_bytes = ctx->frames_slot_bytes * (size_t)ctx->p.max_batch + 16;
  {
    const int rc = update_tmaps(ctx, *P, ctx->d_frames, (size_t)ctx->frames_pitch, ctx->frames_slot_bytes, ctx->p.max_batch);
    if (rc) return rc;
  }
  B->out_kps = ctx->d_kps; B->out_angles = ctx->d_angles; B->out_desc = ctx->d_desc; B->out_n = ctx->d_nout;
  B->out_cap = ctx->list_cap;
  return ORB_OK;
}

}  // namespace

int orb_internal_get_plan(orb_ctx* ctx, int w, int h) { return get_plan(ctx, w, h); }

extern "C" {

int orb_abi_version(void) { return ORB_B200_ABI_VERSION; }

void orb_default_params(orb_params* p) {
  memset(p, 0, sizeof(*p));
  p->nfeatures = 500; p->scale_factor = 1.2f; p->nlevels = 8;            // ref include/orb.hpp:36
  p->fast_threshold = 20; p->fast_n = 9; p->nms_window = 3; p->orient_patch = 31;   // ref include/orb.hpp:12
  p->select_policy = ORB_SELECT_HARRIS_TOP_N; p->blur_levels = 1; p->harris_k = 0.04f;
  p->device = 0; p->max_width = 1241; p->max_height = 376; p->max_batch = 1; p->chunk_frames = 0;
  p->max_keypoints = 0; p->keep_side_arrays = 0;
}

const char* orb_last_error(const orb_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

void orb_destroy(orb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->p.device);
  void* ptrs[] = {ctx->d_frames, ctx->d_pyr, ctx->d_box, ctx->d_cand, ctx->d_cand_count, ctx->d_kept_xy, ctx->d_kept_r,
                  ctx->d_kept_count, ctx->d_xtab, ctx->d_ytab, ctx->d_tile_a, ctx->d_tile_b, ctx->d_tile_b1, ctx->d_harris_w, ctx->d_pattern, ctx->d_flags, ctx->d_kps,
                  ctx->d_angles, ctx->d_desc, ctx->d_nout, ctx->d_side_xy, ctx->d_side_level, ctx->d_side_resp,
                  ctx->d_list_kps, ctx->d_list_angles, ctx->d_list_out, ctx->d_tmaps, ctx->d_edge2};
  for (void* q : ptrs) if (q) cudaFree(q);
  for (auto& sp : ctx->spans) { if (sp.a) cudaEventDestroy(sp.a); if (sp.b) cudaEventDestroy(sp.b); }
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->h_ingest) cudaFreeHost(ctx->h_ingest);
  if (ctx->d_lk) cudaFree(ctx->d_lk);
  if (ctx->d_scores) cudaFree(ctx->d_scores);
  for (void* q : ctx->d_scratch) if (q) cudaFree(q);
  if (ctx->d_match_exp) cudaFree(ctx->d_match_exp);
  if (ctx->h_comp) cudaFreeHost(ctx->h_comp);
  if (ctx->h_descs) cudaFreeHost(ctx->h_descs);
  if (ctx->h_inf_status) cudaFreeHost(ctx->h_inf_status);
  if (ctx->d_comp) cudaFree(ctx->d_comp);
  if (ctx->d_raw) cudaFree(ctx->d_raw);
  if (ctx->d_descs) cudaFree(ctx->d_descs);
  if (ctx->d_inf_status) cudaFree(ctx->d_inf_status);
  if (ctx->d_adler) cudaFree(ctx->d_adler);
  if (ctx->h_crc) cudaFreeHost(ctx->h_crc);
  if (ctx->h_crc_n) cudaFreeHost(ctx->h_crc_n);
  if (ctx->d_crc) cudaFree(ctx->d_crc);
  if (ctx->d_crc_n) cudaFree(ctx->d_crc_n);
  for (cudaStream_t q : ctx->s_ingest) if (q) cudaStreamDestroy(q);
  for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_done) cudaEventDestroy(e);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int orb_create(const orb_params* p, orb_ctx** out) {
  orb_ctx* ctx = nullptr;   // errors before allocation go to the global create-error string
  if (!p || !out) return fail(ctx, ORB_E_INVALID, "null argument");
  *out = nullptr;
  if (p->nlevels < 1 || p->nlevels > ORB_MAX_LEVELS) return fail(ctx, ORB_E_INVALID, "nlevels must be 1..%d", ORB_MAX_LEVELS);
  if (!(p->scale_factor > 1.0f) && p->nlevels > 1) return fail(ctx, ORB_E_INVALID, "scale_factor must be > 1");
  if (p->fast_n < 1 || p->fast_n > 16) return fail(ctx, ORB_E_INVALID, "fast_n must be 1..16");
  if (p->fast_threshold < 0 || p->fast_threshold > 255) return fail(ctx, ORB_E_INVALID, "fast_threshold must be 0..255");
  if (p->nms_window != 1 && p->nms_window != 3 && p->nms_window != 0 && p->nms_window != 2)
    return fail(ctx, ORB_E_INVALID, "nms_window must be 1 or 3 (radius 0 or 1)");
  if (p->orient_patch < 1 || p->orient_patch > 63) return fail(ctx, ORB_E_INVALID, "orient_patch must be 1..63");
  if (p->select_policy != ORB_SELECT_RASTER_FIRST_N && p->select_policy != ORB_SELECT_HARRIS_TOP_N)
    return fail(ctx, ORB_E_INVALID, "unknown select_policy");
  if (p->nfeatures < 1) return fail(ctx, ORB_E_INVALID, "nfeatures must be >= 1");
  if (p->max_width < 1 || p->max_height < 1 || p->max_width > 65535 || p->max_height > 65535 || p->max_batch < 1)
    return fail(ctx, ORB_E_INVALID, "bad capacity fields");
  // per-level pixel counts and offsets are 32-bit in the plan: 2^28 pixels per frame keep every product below 2^31
  if ((long long)p->max_width * p->max_height > (1ll << 28))
    return fail(ctx, ORB_E_CAPACITY, "frame capacity %dx%d exceeds 2^28 pixels", p->max_width, p->max_height);
  if ((p->select_policy == ORB_SELECT_RASTER_FIRST_N ? p->nfeatures : level_quota(p->nfeatures, p->scale_factor, p->nlevels, 0)) > ORB_SORT_CAP)
    return fail(ctx, ORB_E_INVALID, "per-level keypoint budget exceeds %d", ORB_SORT_CAP);

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(ctx, ORB_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (p->device < 0 || p->device >= ndev) return fail(ctx, ORB_E_INVALID, "device %d out of range (%d devices)", p->device, ndev);

  ctx = new orb_ctx();
  ctx->p = *p;
  ctx->err[0] = 0;
  int rc = ORB_OK;
  auto body = [&]() -> int {
    CK(cudaSetDevice(p->device));
    CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    build_plan(ctx->p, p->max_width, p->max_height, p->nlevels, p->select_policy, -1, &ctx->max_plan);
    const OrbPlan& M = ctx->max_plan;
    // single-image stages reuse slot 0 with a 1-level plan whose kept list may hold ORB_SORT_CAP entries
    OrbPlan S;
    build_plan(ctx->p, p->max_width, p->max_height, 1, ORB_SELECT_RASTER_FIRST_N, ORB_SORT_CAP, &S);
    int kept_per_frame = std::max(M.kept_per_frame, S.kept_per_frame);
    // outputs per frame
    int total_quota = 0;
    for (int l = 0; l < M.nlevels; l++) total_quota += M.lv[l].quota;
    ctx->max_kp = p->max_keypoints > 0 ? p->max_keypoints : total_quota;
    // chunking: keep one chunk's scratch (levels + box sums) well inside the 126 MB L2
    size_t per_frame = (size_t)M.pyr_frame_bytes + (size_t)M.box_frame_elems * 2 + (size_t)p->max_width * p->max_height;
    // measured on B200: larger waves win (launch gaps and kernel tails outweigh L2 residency of the scratch; 1000 KITTI
    // frames: 162.7 k frames/s in waves of 512, 164.5 k in one wave): up to 1024 frames per wave when frames and results
    // stay on the device, bounded by 6 GB of scratch (1394 KITTI frames, ~300 at 1080p, ~75 at 4K)
    int chunk = p->chunk_frames > 0 ? p->chunk_frames : (int)std::max<size_t>(1, std::min<size_t>(1024, ((size_t)6 << 30) / per_frame));
    // host-staged waves: ~30 MB of frames each (64 KITTI frames, 14 at 1080p, 4 at 4K) -- small enough that the first wave's
    // upload and the last wave's computation, the two parts the link cannot hide, stay short; measured on KITTI frames:
    // 32-64 frames 107-108 k frames/s end to end, 96-128 frames 103 k, 24 frames 101 k (launch overhead)
    ctx->chunk_staged = p->chunk_frames > 0 ? p->chunk_frames
                                            : (int)std::max<size_t>(2, std::min<size_t>(128, (30000000 + (size_t)p->max_width * p->max_height / 2) /
                                                                                                 ((size_t)p->max_width * p->max_height)));
    if (const char* e = getenv("ORB_B200_STAGED_WAVE")) { const int v = atoi(e); if (v > 0) ctx->chunk_staged = v; }   // tuning knob
    ctx->chunk = std::max(1, std::min(chunk, p->max_batch));
    ctx->chunk_staged = std::max(1, std::min(ctx->chunk_staged, ctx->chunk));
    const int C = ctx->chunk, Bn = p->max_batch;
    // staged rows are 16-byte multiples with at least one spare byte after the last pixel (k_pyramid's second tap)
    ctx->frames_pitch = align_up(p->max_width + 1, 16);
    ctx->frames_slot_bytes = (size_t)ctx->frames_pitch * p->max_height;
    CK(cudaMalloc(&ctx->d_frames, ctx->frames_slot_bytes * Bn + 16));   // k_pyramid may read one byte past the last row
    CK(cudaMalloc(&ctx->d_pyr, (size_t)M.pyr_frame_bytes * C));
    CK(cudaMalloc(&ctx->d_box, (size_t)M.box_frame_elems * 2 * C));
    CK(cudaMalloc(&ctx->d_cand, (size_t)M.cand_frame_elems * 8 * C));
    // per-chunk accumulators, zeroed by one memset per chunk: candidate counters | BRIEF border tables
    {
      const int edge_max = std::max(M.edge_frame_elems, S.edge_frame_elems);
      ctx->zero_bytes_per_frame = sizeof(int) * ORB_MAX_LEVELS + sizeof(int) * (size_t)edge_max;
      CK(cudaMalloc(&ctx->d_cand_count, ctx->zero_bytes_per_frame * C));
    }
    CK(cudaMalloc(&ctx->d_kept_count, sizeof(int) * ORB_MAX_LEVELS * C));
    CK(cudaMalloc(&ctx->d_edge2, sizeof(int) * (size_t)std::max(M.edge2_frame_elems, S.edge2_frame_elems) * C));
    CK(cudaMalloc(&ctx->d_kept_xy, sizeof(uint32_t) * (size_t)kept_per_frame * C));
    CK(cudaMalloc(&ctx->d_kept_r, sizeof(float) * (size_t)kept_per_frame * C));
    ctx->xtab_cap = 0; ctx->ytab_cap = 0;
    for (int l = 0; l < M.nlevels; l++) { ctx->xtab_cap += M.lv[l].w; ctx->ytab_cap += M.lv[l].h; }
    ctx->tile_a_cap = std::max(M.a_tiles_per_frame, 1); ctx->tile_b_cap = M.tiles_per_frame;
    CK(cudaMalloc(&ctx->d_tile_a, 4 * (size_t)ctx->tile_a_cap));
    CK(cudaMalloc(&ctx->d_tile_b, 4 * (size_t)ctx->tile_b_cap));
    CK(cudaMalloc(&ctx->d_tile_b1, 4 * (size_t)S.tiles_per_frame));   // single-level stage plans
    CK(cudaMalloc(&ctx->d_xtab, sizeof(OrbTap) * ctx->xtab_cap));
    CK(cudaMalloc(&ctx->d_ytab, sizeof(OrbTap) * ctx->ytab_cap));
    CK(cudaMalloc(&ctx->d_harris_w, sizeof(float) * 49));
    CK(cudaMalloc(&ctx->d_pattern, sizeof(float4) * 256));
    CK(cudaMalloc(&ctx->d_flags, sizeof(int)));
    CK(cudaMalloc(&ctx->d_tmaps, sizeof(ctx->h_tmaps)));
    {
      cudaDriverEntryPointQueryResult q;
      CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->tmap_encode, cudaEnableDefault, &q));
      if (q != cudaDriverEntryPointSuccess || !ctx->tmap_encode) return fail(ctx, ORB_E_CUDA, "driver has no cuTensorMapEncodeTiled (TMA)");
    }
    CK(cudaMallocHost(&ctx->h_flags, sizeof(int)));
    ctx->list_cap = std::max(ORB_SORT_CAP, Bn * ctx->max_kp);
    size_t nrec = (size_t)ctx->list_cap;
    CK(cudaMalloc(&ctx->d_kps, sizeof(orb_keypoint) * nrec));
    CK(cudaMalloc(&ctx->d_angles, sizeof(float) * nrec));
    CK(cudaMalloc(&ctx->d_desc, sizeof(orb_descriptor) * nrec));
    CK(cudaMalloc(&ctx->d_nout, sizeof(int) * Bn));
    if (p->keep_side_arrays) {
      CK(cudaMalloc(&ctx->d_side_xy, sizeof(orb_keypoint) * nrec));
      CK(cudaMalloc(&ctx->d_side_level, sizeof(int) * nrec));
      CK(cudaMalloc(&ctx->d_side_resp, sizeof(float) * nrec));
    }
    CK(cudaMalloc(&ctx->d_list_kps, sizeof(orb_keypoint) * nrec));
    CK(cudaMalloc(&ctx->d_list_angles, sizeof(float) * nrec));
    CK(cudaMalloc(&ctx->d_list_out, sizeof(float) * nrec));
    harris_weights(ctx->harris_w);
    CK(cudaMemcpy(ctx->d_harris_w, ctx->harris_w, sizeof(float) * 49, cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(orbk::c_harris_w, ctx->harris_w, sizeof(float) * 49));   // identical for every context
    {
      float pat[1024];   // the 256 tests as floats (x1,y1,x2,y2): the reference converts them per use (src/orb_cpu.cpp:228)
      for (int i = 0; i < 1024; i++) pat[i] = (float)ORB_BRIEF_PATTERN_31[i];
      CK(cudaMemcpy(ctx->d_pattern, pat, sizeof(pat), cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(ctx->d_flags, 0, sizeof(int)));
    CK(cudaFuncSetAttribute(orbk::k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, orbk::B_SMEM));
    CK(cudaFuncSetAttribute(orbk::k_match_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, orbk::MT_SMEM));
    { int dev = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, dev)); }
    CK(cudaFuncSetAttribute(orbk::k_select, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (2 * ORB_SORT_CAP + orbk::K2_SMEM_KEYS) * 8 + (orbk::K2_MAX_ROWS + 1) * 4));
    return ORB_OK;
  };
  rc = body();
  if (rc != ORB_OK) {
    snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
    orb_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return ORB_OK;
}

int orb_set_stream(orb_ctx* ctx, void* s) {
  if (!ctx) return ORB_E_INVALID;
  ctx->stream = (cudaStream_t)s;
  return ORB_OK;
}

int orb_use_own_stream(orb_ctx* ctx) {
  if (!ctx) return ORB_E_INVALID;
  ctx->stream = ctx->own_stream;
  return ORB_OK;
}

int orb_synchronize(orb_ctx* ctx) {
  if (!ctx) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  return check_flags(ctx);
}

int orb_level_size(const orb_ctx* ctx, int w, int h, int level, int* lw, int* lh) {
  if (!ctx || level < 0 || level >= ctx->p.nlevels || !lw || !lh) return ORB_E_INVALID;
  level_size(w, h, ctx->p.scale_factor, level, lw, lh);
  return ORB_OK;
}
int orb_level_quota(const orb_ctx* ctx, int level) {
  if (!ctx || level < 0 || level >= ctx->p.nlevels) return ORB_E_INVALID;
  return ctx->p.select_policy == ORB_SELECT_HARRIS_TOP_N ? level_quota(ctx->p.nfeatures, ctx->p.scale_factor, ctx->p.nlevels, level)
                                                         : ctx->p.nfeatures;
}
int orb_get_harris_weights(const orb_ctx* ctx, float* w49) {
  if (!ctx || !w49) return ORB_E_INVALID;
  memcpy(w49, ctx->harris_w, sizeof(float) * 49);
  return ORB_OK;
}
int orb_last_launch_count(const orb_ctx* ctx) { return ctx ? ctx->launches : ORB_E_INVALID; }

int orb_set_profiling(orb_ctx* ctx, int enable) {
  if (!ctx) return ORB_E_INVALID;
  ctx->profiling = enable != 0;
  ctx->spans_used = 0;
  return ORB_OK;
}

int orb_get_stage_ms(orb_ctx* ctx, float ms[5], int launches[5]) {
  if (!ctx || !ms || !launches) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 5; i++) { ms[i] = 0.f; launches[i] = 0; }
  for (size_t i = 0; i < ctx->spans_used; i++) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, ctx->spans[i].a, ctx->spans[i].b));
    ms[ctx->spans[i].stage] += t;
    launches[ctx->spans[i].stage]++;
  }
  ctx->spans_used = 0;
  return ORB_OK;
}

}  // extern "C"

static int run_batch_body(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h,
                     size_t pitch, size_t frame_stride, int cap, orb_keypoint* kps, float* angles,
                     orb_descriptor* desc, int* n_out, int outputs_on_device, WaveSource* source);

int orb_internal_run_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h,
                     size_t pitch, size_t frame_stride, int cap, orb_keypoint* kps, float* angles,
                     orb_descriptor* desc, int* n_out, int outputs_on_device, WaveSource* source) {
  const int rc = run_batch_body(ctx, frames, frames_on_device, n_frames, w, h, pitch, frame_stride, cap, kps, angles, desc, n_out,
                                outputs_on_device, source);
  if (rc != ORB_OK && ctx && ctx->s_h2d) {
    // a failure in the middle of the wave pipeline: copies into / out of the caller's host buffers may still be in flight on
    // the copy streams -- drain them before the buffers go back to the caller (the error text is already in ctx->err)
    cudaStreamSynchronize(ctx->s_h2d);
    cudaStreamSynchronize(ctx->s_d2h);
    cudaStreamSynchronize(ctx->s_side);
    cudaStreamSynchronize(ctx->stream);
  }
  return rc;
}

static int run_batch_body(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h,
                     size_t pitch, size_t frame_stride, int cap, orb_keypoint* kps, float* angles,
                     orb_descriptor* desc, int* n_out, int outputs_on_device, WaveSource* source) {
  if (!ctx) return ORB_E_INVALID;
  if ((!frames && !source) || !kps || !angles || !desc || !n_out) return fail(ctx, ORB_E_INVALID, "null buffer");
  if (n_frames < 1 || cap < 1 || (!source && (pitch < (size_t)w || frame_stride < pitch * (size_t)(h - 1) + w)))
    return fail(ctx, ORB_E_INVALID, "bad batch geometry");
  if (n_frames > ctx->p.max_batch) return fail(ctx, ORB_E_CAPACITY, "batch %d exceeds max_batch %d", n_frames, ctx->p.max_batch);
  if (!outputs_on_device && cap > ctx->max_kp)
    return fail(ctx, ORB_E_CAPACITY, "cap %d exceeds max_keypoints %d of the context", cap, ctx->max_kp);
  if (ctx->p.keep_side_arrays && cap > ctx->max_kp) return fail(ctx, ORB_E_CAPACITY, "cap exceeds side-array capacity");
  CK(cudaSetDevice(ctx->p.device));
  int rc = get_plan(ctx, w, h);
  if (rc) return rc;
  const OrbPlan& P = ctx->plan;
  ctx->launches = 0;

  // The kernels read level 0 with 16-byte loads: frames must be 16-byte aligned with 16-byte multiples as row pitch
  // and frame stride.  Host frames (and device frames that are not) go through the staging area, chunk by chunk on a
  // copy stream so that the transfer of chunk i+1 overlaps the kernels of chunk i; host outputs leave on a third
  // stream as soon as their chunk is described.
  // (and pitch > w: k_pyramid reads the byte after a row's last pixel, which must belong to the caller's buffer)
  const bool direct = !source && frames_on_device && ((uintptr_t)frames % 16 == 0) && pitch % 16 == 0 && frame_stride % 16 == 0 && pitch > (size_t)w;
  const uint8_t* src = direct ? frames : ctx->d_frames;
  const size_t stride = direct ? frame_stride : ctx->frames_slot_bytes;
  const int sp = direct ? (int)pitch : ctx->frames_pitch;
  orb_keypoint* o_kps = outputs_on_device ? kps : ctx->d_kps;
  float* o_ang = outputs_on_device ? angles : ctx->d_angles;
  orb_descriptor* o_desc = outputs_on_device ? desc : ctx->d_desc;
  int* o_n = outputs_on_device ? n_out : ctx->d_nout;

  if ((rc = update_tmaps(ctx, P, src, (size_t)sp, stride, direct ? n_frames : ctx->p.max_batch))) return rc;

  int total_quota = 0;
  for (int l = 0; l < P.nlevels; l++) total_quota += P.lv[l].quota;
  const int nwarps = std::min(total_quota, cap);
  // wave boundaries (wave_schedule): full waves; host-staged batches start with short waves and end with short waves
  std::vector<int> wave_begin;
  {
    int wave = (direct && outputs_on_device) ? ctx->chunk : ctx->chunk_staged;
    if (source && source->preferred_wave(ctx) > 0) wave = std::min(ctx->chunk, source->preferred_wave(ctx));
    wave_schedule(n_frames, wave, !direct, ORB_E2E_RAMP_DOWN && !direct && !outputs_on_device && !source, wave_begin);
  }
  const int nchunks = (int)wave_begin.size() - 1;
  const bool piped = !direct || !outputs_on_device;
  if (piped) {
    while ((int)ctx->ev_in.size() < nchunks) {
      cudaEvent_t a = nullptr, b = nullptr;
      CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      ctx->ev_in.push_back(a); ctx->ev_done.push_back(b);
    }
    // the copy streams start after whatever the caller's stream has queued so far (it may still use the staging area)
    CK(cudaEventRecord(ctx->ev_start, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_start, 0));
  }
  if (!direct && !source) {
    const cudaMemcpyKind kind = frames_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    for (int ci = 0; ci < nchunks; ci++) {
      const int c0 = wave_begin[ci], nc = wave_begin[ci + 1] - c0;
      uint8_t* dst = ctx->d_frames + (size_t)c0 * ctx->frames_slot_bytes;
      const uint8_t* from = frames + (size_t)c0 * frame_stride;
      if (pitch == (size_t)ctx->frames_pitch && frame_stride == ctx->frames_slot_bytes) {
        CK(cudaMemcpyAsync(dst, from, ctx->frames_slot_bytes * nc, kind, ctx->s_h2d));      // identical layout: one linear copy
      } else if (frame_stride == pitch * (size_t)h && ctx->frames_slot_bytes == (size_t)ctx->frames_pitch * h) {
        CK(cudaMemcpy2DAsync(dst, ctx->frames_pitch, from, pitch, w, (size_t)h * nc, kind, ctx->s_h2d));
      } else {
        for (int f = 0; f < nc; f++)
          CK(cudaMemcpy2DAsync(dst + (size_t)f * ctx->frames_slot_bytes, ctx->frames_pitch, from + (size_t)f * frame_stride,
                               pitch, w, h, kind, ctx->s_h2d));
      }
      CK(cudaEventRecord(ctx->ev_in[ci], ctx->s_h2d));
    }
  }

  for (int ci = 0; ci < nchunks; ci++) {
    const int c0 = wave_begin[ci], nc = wave_begin[ci + 1] - c0;
    if (source) {
      if ((rc = source->stage(ctx, ci, c0, nc, ctx->ev_in[ci]))) {
        // frames of earlier waves are still in flight: drain before handing the buffers back
        cudaDeviceSynchronize();
        return rc;
      }
    }
    if (!direct) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[ci], 0));
    Bufs B;
    fill_bufs(ctx, &B);
    B.frames = src + (size_t)c0 * stride; B.frame_stride = stride; B.pitch0 = sp; B.frame0 = c0;
    B.frames_bytes = direct ? stride * (size_t)(n_frames - c0) : ctx->frames_slot_bytes * (size_t)(ctx->p.max_batch - c0) + 16;
    B.out_kps = o_kps + (size_t)c0 * cap; B.out_angles = o_ang + (size_t)c0 * cap; B.out_desc = o_desc + (size_t)c0 * cap;
    B.out_n = o_n + c0; B.out_cap = cap;
    if (ctx->p.keep_side_arrays) {
      B.side_xy = ctx->d_side_xy + (size_t)c0 * cap; B.side_level = ctx->d_side_level + (size_t)c0 * cap;
      B.side_resp = ctx->d_side_resp + (size_t)c0 * cap;
    }
    if ((rc = launch_pyramid_fast(ctx, P, B, nc))) return rc;
    if ((rc = launch_select(ctx, P, B, nc))) return rc;
    DescribeJob J{0, nullptr, nullptr, 0};
    if ((rc = launch_describe(ctx, P, B, J, nwarps, nc))) return rc;
    if (nwarps == 0) CK(cudaMemsetAsync(B.out_n, 0, sizeof(int) * nc, ctx->stream));
    ctx->last_chunk_start = c0; ctx->last_chunk_n = nc;
    if (!outputs_on_device) {
      CK(cudaEventRecord(ctx->ev_done[ci], ctx->stream));
      CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[ci], 0));
      const size_t o = (size_t)c0 * cap, nrec = (size_t)nc * cap;
      CK(cudaMemcpyAsync(kps + o, ctx->d_kps + o, sizeof(orb_keypoint) * nrec, cudaMemcpyDeviceToHost, ctx->s_d2h));
      CK(cudaMemcpyAsync(angles + o, ctx->d_angles + o, sizeof(float) * nrec, cudaMemcpyDeviceToHost, ctx->s_d2h));
      CK(cudaMemcpyAsync(desc + o, ctx->d_desc + o, sizeof(orb_descriptor) * nrec, cudaMemcpyDeviceToHost, ctx->s_d2h));
      CK(cudaMemcpyAsync(n_out + c0, ctx->d_nout + c0, sizeof(int) * nc, cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
  }
  ctx->last_n = n_frames; ctx->last_cap = cap;
  ctx->last_frames = src; ctx->last_stride = stride; ctx->last_pitch = sp;
  ctx->last_outputs_ctx = !outputs_on_device;

  if (!outputs_on_device) {
    CK(cudaStreamSynchronize(ctx->s_d2h));
    return check_flags(ctx);
  }
  return ORB_OK;
}

extern "C" {

int orb_detect_and_compute_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h,
                                 size_t pitch, size_t frame_stride, int cap, orb_keypoint* kps, float* angles,
                                 orb_descriptor* desc, int* n_out, int outputs_on_device) {
  if (ctx && !frames) return fail(ctx, ORB_E_INVALID, "null buffer");
  return orb_internal_run_batch(ctx, frames, frames_on_device, n_frames, w, h, pitch, frame_stride, cap, kps, angles, desc, n_out,
                   outputs_on_device, nullptr);
}

int orb_detect_and_compute(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, int cap, orb_keypoint* kps,
                           float* angles, orb_descriptor* desc, int* n_out, int* n_per_level) {
  if (!ctx) return ORB_E_INVALID;
  int rc = orb_detect_and_compute_batch(ctx, img, 0, 1, w, h, pitch, pitch * (size_t)h, cap, kps, angles, desc, n_out, 0);
  if (rc) return rc;
  if (n_per_level) {
    int kc[ORB_MAX_LEVELS];
    CK(cudaMemcpy(kc, ctx->d_kept_count, sizeof(kc), cudaMemcpyDeviceToHost));
    int left = *n_out;
    for (int l = 0; l < ctx->p.nlevels; l++) { n_per_level[l] = std::min(kc[l], left); left -= n_per_level[l]; }
  }
  return ORB_OK;
}

int orb_get_level(orb_ctx* ctx, int frame, int level, uint8_t* dst, size_t dst_pitch, int* w, int* h) {
  if (!ctx || !dst) return ORB_E_INVALID;
  if (level < 0 || level >= ctx->p.nlevels) return fail(ctx, ORB_E_INVALID, "level out of range");
  if (frame < ctx->last_chunk_start || frame >= ctx->last_chunk_start + ctx->last_chunk_n || ctx->last_n == 0)
    return fail(ctx, ORB_E_INVALID, "frame %d is not resident (last chunk holds frames %d..%d)", frame, ctx->last_chunk_start,
                ctx->last_chunk_start + ctx->last_chunk_n - 1);
  CK(cudaSetDevice(ctx->p.device));
  const OrbLevel& G = ctx->plan.lv[level];
  if (dst_pitch < (size_t)G.w) return fail(ctx, ORB_E_INVALID, "dst_pitch too small");
  const uint8_t* s; size_t sp;
  if (level == 0) { s = ctx->last_frames + (size_t)frame * ctx->last_stride; sp = ctx->last_pitch; }
  else { s = ctx->d_pyr + (size_t)(frame - ctx->last_chunk_start) * ctx->plan.pyr_frame_bytes + G.lvl_ofs; sp = G.pitch; }
  CK(cudaMemcpy2DAsync(dst, dst_pitch, s, sp, G.w, G.h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (w) *w = G.w;
  if (h) *h = G.h;
  return ORB_OK;
}

int orb_get_side_arrays(orb_ctx* ctx, int frame, int n, orb_keypoint* level_xy, int32_t* level_id, float* response) {
  if (!ctx) return ORB_E_INVALID;
  if (!ctx->p.keep_side_arrays) return fail(ctx, ORB_E_INVALID, "context was created without keep_side_arrays");
  if (frame < 0 || frame >= ctx->last_n || n < 0 || n > ctx->last_cap) return fail(ctx, ORB_E_INVALID, "frame / n out of range");
  CK(cudaSetDevice(ctx->p.device));
  size_t o = (size_t)frame * ctx->last_cap;
  if (level_xy) CK(cudaMemcpyAsync(level_xy, ctx->d_side_xy + o, sizeof(orb_keypoint) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (level_id) CK(cudaMemcpyAsync(level_id, ctx->d_side_level + o, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (response) CK(cudaMemcpyAsync(response, ctx->d_side_resp + o, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}

int orb_get_candidates(orb_ctx* ctx, int frame, int level, int cap, orb_keypoint* xy, float* response, int* n_out) {
  if (!ctx || !n_out) return ORB_E_INVALID;
  if (level < 0 || level >= ctx->p.nlevels) return fail(ctx, ORB_E_INVALID, "level out of range");
  if (frame < ctx->last_chunk_start || frame >= ctx->last_chunk_start + ctx->last_chunk_n || ctx->last_n == 0)
    return fail(ctx, ORB_E_INVALID, "frame %d is not resident", frame);
  CK(cudaSetDevice(ctx->p.device));
  const int slot = frame - ctx->last_chunk_start;
  const OrbLevel& G = ctx->plan.lv[level];
  int n = 0;
  CK(cudaMemcpy(&n, ctx->d_cand_count + (size_t)slot * (ctx->zero_bytes_per_frame / sizeof(int)) + level, sizeof(int), cudaMemcpyDeviceToHost));
  *n_out = n;
  int m = std::min(std::min(n, cap), G.cand_cap);
  if (m > 0 && (xy || response)) {
    std::vector<unsigned long long> keys(m);
    CK(cudaMemcpy(keys.data(), ctx->d_cand + (size_t)slot * ctx->plan.cand_frame_elems + G.cand_ofs, 8 * (size_t)m,
                  cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
      uint32_t lo = (uint32_t)keys[i], hi = ~(uint32_t)(keys[i] >> 32);
      if (xy) xy[i] = orb_keypoint{(int)(lo & 0xffff), (int)(lo >> 16)};
      if (response) {
        uint32_t u = (hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi;
        float r;
        memcpy(&r, &u, 4);
        response[i] = ctx->p.select_policy == ORB_SELECT_HARRIS_TOP_N ? r : 0.0f;
      }
    }
  }
  return ORB_OK;
}

// ---- single-image stage entry points ----------------------------------------------------------
int orb_fast_detect(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, int nfeatures, orb_keypoint* kps, int* n_out) {
  if (!ctx || !kps || !n_out) return ORB_E_INVALID;
  if (nfeatures < 0 || nfeatures > ORB_SORT_CAP) return fail(ctx, ORB_E_CAPACITY, "nfeatures must be 0..%d", ORB_SORT_CAP);
  CK(cudaSetDevice(ctx->p.device));
  int rc = stage_image(ctx, img, w, h, pitch);
  if (rc) return rc;
  OrbPlan P; Bufs B;
  if ((rc = stage_plan(ctx, w, h, ORB_SELECT_RASTER_FIRST_N, nfeatures, &P, &B))) return rc;
  if ((rc = launch_pyramid_fast(ctx, P, B, 1))) return rc;
  if ((rc = launch_select(ctx, P, B, 1))) return rc;
  int m = 0;
  CK(cudaMemcpyAsync(&m, ctx->d_kept_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = check_flags(ctx))) return rc;
  std::vector<uint32_t> xy(std::max(m, 1));
  CK(cudaMemcpy(xy.data(), ctx->d_kept_xy, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
  for (int i = 0; i < m; i++) kps[i] = orb_keypoint{(int)(xy[i] & 0xffff), (int)(xy[i] >> 16)};
  *n_out = m;
  return ORB_OK;
}

int orb_harris(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, const orb_keypoint* kps, int n, float* response) {
  if (!ctx || (n > 0 && (!kps || !response))) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  int rc = stage_image(ctx, img, w, h, pitch);
  if (rc) return rc;
  for (int i = 0; i < n; i++)
    if (kps[i].x < 0 || kps[i].x >= w || kps[i].y < 0 || kps[i].y >= h) return fail(ctx, ORB_E_INVALID, "keypoint %d outside the image", i);
  for (int o = 0; o < n; o += ctx->list_cap) {
    int m = std::min(ctx->list_cap, n - o);
    CK(cudaMemcpyAsync(ctx->d_list_kps, kps + o, sizeof(orb_keypoint) * m, cudaMemcpyHostToDevice, ctx->stream));
    orbk::k_harris_list<<<(m + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_frames, ctx->frames_pitch, w, h, ctx->d_list_kps, m,
                                                                 ctx->p.harris_k, ctx->d_list_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(response + o, ctx->d_list_out, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return ORB_OK;
}

static int describe_list(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, const orb_keypoint* kps,
                         const float* angles_in, int n, float* angles_out, orb_descriptor* desc_out) {
  CK(cudaSetDevice(ctx->p.device));
  int rc = stage_image(ctx, img, w, h, pitch);
  if (rc) return rc;
  for (int i = 0; i < n; i++)
    if (kps[i].x < 0 || kps[i].x >= w || kps[i].y < 0 || kps[i].y >= h) return fail(ctx, ORB_E_INVALID, "keypoint %d outside the image", i);
  OrbPlan P; Bufs B;
  if ((rc = stage_plan(ctx, w, h, ORB_SELECT_RASTER_FIRST_N, 0, &P, &B))) return rc;
  if (desc_out && (rc = launch_pyramid_fast(ctx, P, B, 1))) return rc;   // builds the box-sum image of the frame
  for (int o = 0; o < n; o += ctx->list_cap) {
    int m = std::min(ctx->list_cap, n - o);
    CK(cudaMemcpyAsync(ctx->d_list_kps, kps + o, sizeof(orb_keypoint) * m, cudaMemcpyHostToDevice, ctx->stream));
    if (angles_in) CK(cudaMemcpyAsync(ctx->d_list_angles, angles_in + o, sizeof(float) * m, cudaMemcpyHostToDevice, ctx->stream));
    DescribeJob J{desc_out ? 2 : 1, ctx->d_list_kps, ctx->d_list_angles, m};
    if ((rc = launch_describe(ctx, P, B, J, m, 1))) return rc;
    if (angles_out) CK(cudaMemcpyAsync(angles_out + o, ctx->d_angles, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc_out) CK(cudaMemcpyAsync(desc_out + o, ctx->d_desc, sizeof(orb_descriptor) * m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return ORB_OK;
}

int orb_orientations(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, const orb_keypoint* kps, int n, float* angles) {
  if (!ctx || (n > 0 && (!kps || !angles))) return ORB_E_INVALID;
  return describe_list(ctx, img, w, h, pitch, kps, nullptr, n, angles, nullptr);
}

int orb_brief(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, const orb_keypoint* kps, const float* angles, int n,
              orb_descriptor* desc) {
  if (!ctx || (n > 0 && (!kps || !angles || !desc))) return ORB_E_INVALID;
  return describe_list(ctx, img, w, h, pitch, kps, angles, n, nullptr, desc);
}

int orb_nms_scores(orb_ctx* ctx, const float* scores, int w, int h, size_t pitch_bytes, int nms_window, int nfeatures, float threshold,
                   orb_keypoint* kps, int* n_out) {
  if (!ctx || !scores || !kps || !n_out) return ORB_E_INVALID;
  if (w < 1 || h < 1 || pitch_bytes < (size_t)w * 4 || pitch_bytes % 4) return fail(ctx, ORB_E_INVALID, "bad score map geometry");
  if (w > ctx->p.max_width || h > ctx->p.max_height)
    return fail(ctx, ORB_E_CAPACITY, "score map %dx%d exceeds context capacity %dx%d", w, h, ctx->p.max_width, ctx->p.max_height);
  if (nfeatures < 0 || nfeatures > ORB_SORT_CAP) return fail(ctx, ORB_E_CAPACITY, "nfeatures must be 0..%d", ORB_SORT_CAP);
  if (nms_window < 1 || nms_window > 7) return fail(ctx, ORB_E_INVALID, "nms_window must be 1..7");
  CK(cudaSetDevice(ctx->p.device));
  const size_t need = (size_t)w * h * sizeof(float);
  if (need > ctx->d_scores_bytes) {             // grow-only scratch for the uploaded map
    if (ctx->d_scores) CK(cudaFree(ctx->d_scores));
    ctx->d_scores = nullptr; ctx->d_scores_bytes = 0;
    CK(cudaMalloc(&ctx->d_scores, need));
    ctx->d_scores_bytes = need;
  }
  CK(cudaMemcpy2DAsync(ctx->d_scores, (size_t)w * 4, scores, pitch_bytes, (size_t)w * 4, h, cudaMemcpyHostToDevice, ctx->stream));
  ctx->last_n = 0;
  OrbPlan P; Bufs B;
  int rc;
  if ((rc = stage_plan(ctx, w, h, ORB_SELECT_RASTER_FIRST_N, nfeatures, &P, &B))) return rc;
  CK(cudaMemsetAsync(ctx->d_cand_count, 0, ctx->zero_bytes_per_frame, ctx->stream));
  orbk::k_nms_scores<<<dim3((w + 127) / 128, h), 128, 0, ctx->stream>>>(ctx->d_scores, w, w, h, nms_window / 2, threshold, ctx->d_cand,
                                                                      ctx->d_cand_count, P.lv[0].cand_cap);
  CK(cudaGetLastError());
  ctx->launches = 1;
  if ((rc = launch_select(ctx, P, B, 1))) return rc;
  int m = 0;
  CK(cudaMemcpyAsync(&m, ctx->d_kept_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = check_flags(ctx))) return rc;
  std::vector<uint32_t> xy(std::max(m, 1));
  CK(cudaMemcpy(xy.data(), ctx->d_kept_xy, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
  for (int i = 0; i < m; i++) kps[i] = orb_keypoint{(int)(xy[i] & 0xffff), (int)(xy[i] >> 16)};
  *n_out = m;
  return ORB_OK;
}

// ---- the reference's stand-alone filter wrappers ------------------------------------------------------------
static int scratch(orb_ctx* ctx, int slot, size_t bytes, void** out);

int orb_conv2d_u8(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, const float* kernel, int ksize, int border_reflect101,
                  float divisor, uint8_t* out, size_t out_pitch) {
  if (!ctx || !kernel || !out) return ORB_E_INVALID;
  if (ksize < 1 || ksize > 31 || !(ksize & 1)) return fail(ctx, ORB_E_INVALID, "kernel size must be odd, 1..31");
  const int ow = border_reflect101 ? w : w - ksize + 1, oh = border_reflect101 ? h : h - ksize + 1;
  if (ow < 1 || oh < 1 || out_pitch < (size_t)ow) return fail(ctx, ORB_E_INVALID, "image smaller than the kernel / bad output pitch");
  CK(cudaSetDevice(ctx->p.device));
  int rc = stage_image(ctx, img, w, h, pitch);
  if (rc) return rc;
  uint8_t* d_out = nullptr; float* d_k = nullptr;
  if ((rc = scratch(ctx, 0, (size_t)ow * oh, (void**)&d_out)) || (rc = scratch(ctx, 1, sizeof(float) * ksize * ksize, (void**)&d_k))) return rc;
  CK(cudaMemcpyAsync(d_k, kernel, sizeof(float) * ksize * ksize, cudaMemcpyHostToDevice, ctx->stream));
  orbk::k_conv2d_u8<<<dim3((ow + 127) / 128, oh), 128, sizeof(float) * ksize * ksize, ctx->stream>>>(
      ctx->d_frames, ctx->frames_pitch, w, h, d_k, ksize, border_reflect101, divisor, d_out, ow, ow, oh);
  CK(cudaGetLastError());
  CK(cudaMemcpy2DAsync(out, out_pitch, d_out, ow, ow, oh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}

int orb_gaussian_blur_1d(orb_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch, uint8_t* out, size_t out_pitch) {
  if (!ctx || !out) return ORB_E_INVALID;
  if (out_pitch < (size_t)w) return fail(ctx, ORB_E_INVALID, "bad output pitch");
  CK(cudaSetDevice(ctx->p.device));
  int rc = stage_image(ctx, img, w, h, pitch);
  if (rc) return rc;
  uint8_t* d_out = nullptr;
  if ((rc = scratch(ctx, 0, (size_t)w * h, (void**)&d_out))) return rc;
  orbk::k_gauss1d_u8<<<dim3((w + 127) / 128, h), 128, 0, ctx->stream>>>(ctx->d_frames, ctx->frames_pitch, w, h, d_out, w);
  CK(cudaGetLastError());
  CK(cudaMemcpy2DAsync(out, out_pitch, d_out, w, w, h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}

// grow-only device scratch of the context for the host-buffer paths of the matcher / debug calls (nothing is allocated
// per call once it is large enough; freed in orb_destroy)
static int scratch(orb_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes > ctx->scratch_bytes[slot]) {
    if (ctx->d_scratch[slot]) CK(cudaFree(ctx->d_scratch[slot]));
    ctx->d_scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0;
    CK(cudaMalloc(&ctx->d_scratch[slot], bytes));
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->d_scratch[slot];
  return ORB_OK;
}

// ---- descriptor matching -------------------------------------------------------------------------
// Tensor-core path (orb_match_tc.cuh): descriptors are expanded to +-1 int8 rows in a grow-only device buffer, two tensor
// maps (query rows, train rows; 128-byte swizzle) are encoded for the call; persistent CTAs walk over (pair, 256 queries) items.
static int match_tc(orb_ctx* ctx, const orb_descriptor* dq, int rows_q, long long stride_q, const orb_descriptor* dt, int rows_t,
                    long long stride_t, bool same_buffer, const int* dn, int nq, int nt, int npairs, long long out_stride, orb_match* dout) {
  if (npairs <= 0 || nq <= 0) return ORB_OK;
  if (rows_t > orbk::MT_MAX_INDEX) return fail(ctx, ORB_E_CAPACITY, "matcher: at most %d train descriptors per set", orbk::MT_MAX_INDEX);
  // expanded rows: [sets_q][rows_q][256] then (unless the train sets are the query sets shifted by one) [sets_t][rows_t][256]
  const int sets_q = same_buffer ? npairs + 1 : npairs, sets_t = same_buffer ? 0 : npairs;
  const size_t bytes_q = (size_t)sets_q * rows_q * orbk::MT_KB, bytes_t = (size_t)sets_t * std::max(rows_t, 1) * orbk::MT_KB;
  if (bytes_q + bytes_t > ctx->match_exp_bytes) {
    if (ctx->d_match_exp) CK(cudaFree(ctx->d_match_exp));
    ctx->d_match_exp = nullptr; ctx->match_exp_bytes = 0;
    CK(cudaMalloc(&ctx->d_match_exp, bytes_q + bytes_t + 1024));
    ctx->match_exp_bytes = bytes_q + bytes_t;
  }
  int8_t* eq = ctx->d_match_exp;
  int8_t* et = same_buffer ? eq + (size_t)rows_q * orbk::MT_KB : eq + bytes_q;
  orbk::k_match_expand<<<dim3((rows_q * 16 + 255) / 256, sets_q), 256, 0, ctx->stream>>>(dq, same_buffer ? dn : nullptr, rows_q, stride_q, rows_q, eq);
  if (!same_buffer && rows_t > 0)
    orbk::k_match_expand<<<dim3((rows_t * 16 + 255) / 256, sets_t), 256, 0, ctx->stream>>>(dt, nullptr, rows_t, stride_t, rows_t, et);
  CK(cudaGetLastError());
  CUtensorMap maps[2];
  int rc;
  if ((rc = encode_map(ctx, &maps[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, eq, orbk::MT_KB, rows_q, npairs, orbk::MT_KB, (size_t)rows_q * orbk::MT_KB,
                       128, 128, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = encode_map(ctx, &maps[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, et, orbk::MT_KB, std::max(rows_t, 1), npairs, orbk::MT_KB,
                       (size_t)std::max(same_buffer ? rows_q : rows_t, 1) * orbk::MT_KB, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  const int qblocks = (nq + orbk::MT_MB * orbk::MT_M - 1) / (orbk::MT_MB * orbk::MT_M);
  const long long n_items = (long long)qblocks * npairs;
  if (n_items > INT32_MAX) return fail(ctx, ORB_E_CAPACITY, "matcher: too many query blocks");
  const int grid = (int)std::min<long long>(n_items, ctx->sm_count);
  orbk::k_match_tc<<<grid, orbk::MT_THREADS, orbk::MT_SMEM, ctx->stream>>>(maps[0], maps[1], dn, nq, nt, qblocks, (int)n_items, out_stride, dout);
  CK(cudaGetLastError());
  ctx->launches += same_buffer ? 2 : 3;
  return ORB_OK;
}

int orb_debug_wave_schedule(int n_frames, int wave, int ramp_up, int ramp_down, int* begins, int cap) {
  std::vector<int> b;
  wave_schedule(n_frames, wave, ramp_up != 0, ramp_down != 0, b);
  for (int i = 0; i < (int)b.size() && i < cap; i++) begins[i] = b[i];
  return (int)b.size();
}

int orb_match_knn2(orb_ctx* ctx, const orb_descriptor* query, int nq, const orb_descriptor* train, int nt, int on_device,
                   orb_match* out) {
  if (!ctx) return ORB_E_INVALID;
  if (nq < 0 || nt < 0 || (nq > 0 && (!query || !out)) || (nt > 0 && !train)) return fail(ctx, ORB_E_INVALID, "bad match arguments");
  if (nq == 0) return ORB_OK;
  CK(cudaSetDevice(ctx->p.device));
  if (on_device) return match_tc(ctx, query, nq, 0, train, nt, 0, false, nullptr, nq, nt, 1, 0, out);
  orb_descriptor *dq = nullptr, *dt = nullptr; orb_match* dm = nullptr;
  int rc;
  if ((rc = scratch(ctx, 0, sizeof(orb_descriptor) * (size_t)nq, (void**)&dq))) return rc;
  if ((rc = scratch(ctx, 1, sizeof(orb_descriptor) * (size_t)std::max(nt, 1), (void**)&dt))) return rc;
  if ((rc = scratch(ctx, 2, sizeof(orb_match) * (size_t)nq, (void**)&dm))) return rc;
  CK(cudaMemcpyAsync(dq, query, sizeof(orb_descriptor) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
  if (nt) CK(cudaMemcpyAsync(dt, train, sizeof(orb_descriptor) * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = match_tc(ctx, dq, nq, 0, dt, nt, 0, false, nullptr, nq, nt, 1, 0, dm))) return rc;
  CK(cudaMemcpyAsync(out, dm, sizeof(orb_match) * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}

int orb_match_knn2_batch(orb_ctx* ctx, const orb_descriptor* desc, const int* n, int n_frames, int cap, int on_device,
                         orb_match* out) {
  if (!ctx) return ORB_E_INVALID;
  if (!desc || !n || !out || n_frames < 1 || cap < 1) return fail(ctx, ORB_E_INVALID, "bad match arguments");
  if (n_frames < 2) return ORB_OK;
  CK(cudaSetDevice(ctx->p.device));
  const long long s = cap;
  if (on_device) return match_tc(ctx, desc, cap, s, desc + cap, cap, s, true, n, cap, cap, n_frames - 1, s, out);
  orb_descriptor* dd = nullptr; int* dn = nullptr; orb_match* dm = nullptr;
  const size_t nd = (size_t)n_frames * cap, nm = (size_t)(n_frames - 1) * cap;
  int rc;
  if ((rc = scratch(ctx, 0, sizeof(orb_descriptor) * nd, (void**)&dd))) return rc;
  if ((rc = scratch(ctx, 1, sizeof(int) * n_frames, (void**)&dn))) return rc;
  if ((rc = scratch(ctx, 2, sizeof(orb_match) * nm, (void**)&dm))) return rc;
  CK(cudaMemcpyAsync(dd, desc, sizeof(orb_descriptor) * nd, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dn, n, sizeof(int) * n_frames, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dm, out, sizeof(orb_match) * nm, cudaMemcpyHostToDevice, ctx->stream));   // entries >= n[p] stay as they were
  if ((rc = match_tc(ctx, dd, cap, s, dd + cap, cap, s, true, dn, cap, cap, n_frames - 1, s, dm))) return rc;
  CK(cudaMemcpyAsync(out, dm, sizeof(orb_match) * nm, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}

void orb_ratio_test(const orb_match* m, int n, float ratio, uint8_t* keep) {
  // if (m.distance < 0.8 * n.distance), reference src/feature_matching.cpp:178 (float distances, double product)
  for (int i = 0; i < n; i++)
    keep[i] = m[i].idx2 >= 0 && (double)(float)m[i].dist1 < (double)ratio * (double)(float)m[i].dist2;
}

int orb_debug_bounds_check(orb_ctx* ctx, int* enabled, unsigned* failures, unsigned* first_line, unsigned* kernels_checked) {
  if (!ctx) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  CK(cudaDeviceSynchronize());
  unsigned v[4] = {0, 0, 0, 0};
  CK(cudaMemcpyFromSymbol(v, orbk::g_orb_bounds, sizeof(v)));
#ifdef ORB_BOUNDS_CHECK
  if (enabled) *enabled = 1;
#else
  if (enabled) *enabled = 0;
#endif
  if (failures) *failures = v[0];
  if (first_line) *first_line = v[1];
  if (kernels_checked) *kernels_checked = v[2];
  return ORB_OK;
}

int orb_debug_bounds_selftest(orb_ctx* ctx) {
  if (!ctx) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  orbk::k_bounds_selftest<<<1, 32, 0, ctx->stream>>>(31);   // lane 31 fails its check in a bounds-check build
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}

int orb_debug_eval_math(orb_ctx* ctx, int op, const float* a, const float* b, int n, float* out) {
  if (!ctx || !a || !out || n < 0 || op < 0 || op > 3 || (op == 0 && !b)) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  float *da = nullptr, *db = nullptr, *dout = nullptr;
  int rc;
  const size_t nb = sizeof(float) * (size_t)std::max(n, 1);
  if ((rc = scratch(ctx, 0, nb, (void**)&da)) || (rc = scratch(ctx, 1, nb, (void**)&db)) || (rc = scratch(ctx, 2, nb, (void**)&dout))) return rc;
  CK(cudaMemcpyAsync(da, a, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (b) CK(cudaMemcpyAsync(db, b, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (n > 0) orbk::k_eval_math<<<(n + 255) / 256, 256, 0, ctx->stream>>>(op, da, db, n, dout);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return ORB_OK;
}


}  // extern "C"
