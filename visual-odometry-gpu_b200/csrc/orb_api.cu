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
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/orb_b200.h"
#include "../../include/orb_brief_pattern.h"
#include "orb_kernels.cuh"
#include "orb_ingest_kernels.cuh"
#include "orb_lk_kernels.cuh"
#include "orb_plan.h"
#include "orb_png.h"

using orbk::Bufs;
using orbk::DescribeJob;


// == extern int bit_pattern_31_[256*4] of the reference (include/orb_pattern.hpp:2)
extern "C" { int bit_pattern_31_[256 * 4]; }
namespace {
struct PatternInit {
  PatternInit() { for (int i = 0; i < 1024; i++) bit_pattern_31_[i] = ORB_BRIEF_PATTERN_31[i]; }
} g_pattern_init;
char g_create_error[512] = "";
}  // namespace

struct orb_ctx {
  orb_params p;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  char err[512];
  int chunk = 1, chunk_staged = 1, max_kp = 0;   // frames per wave: arena capacity / wave size when host copies are involved
  // plan of the last shape + arena limits (plan of the max shape)
  OrbPlan plan, max_plan;
  bool plan_valid = false;
  int xtab_cap = 0, ytab_cap = 0;
  // arena
  uint8_t* d_frames = nullptr; size_t frames_slot_bytes = 0; int frames_pitch = 0;
  uint8_t* d_pyr = nullptr; uint16_t* d_box = nullptr; unsigned long long* d_cand = nullptr;
  int* d_cand_count = nullptr; size_t zero_bytes_per_frame = 0; uint32_t* d_kept_xy = nullptr; float* d_kept_r = nullptr; int* d_kept_count = nullptr;
  OrbTap *d_xtab = nullptr, *d_ytab = nullptr;
  uint32_t *d_tile_a = nullptr, *d_tile_b = nullptr, *d_tile_b1 = nullptr; int tile_a_cap = 0, tile_b_cap = 0;
  float* d_harris_w = nullptr; float4* d_pattern = nullptr; int* d_flags = nullptr; int* h_flags = nullptr;
  orb_keypoint* d_kps = nullptr; float* d_angles = nullptr; orb_descriptor* d_desc = nullptr; int* d_nout = nullptr;
  orb_keypoint* d_side_xy = nullptr; int* d_side_level = nullptr; float* d_side_resp = nullptr;
  orb_keypoint* d_list_kps = nullptr; float* d_list_angles = nullptr; float* d_list_out = nullptr; int list_cap = 0;
  float harris_w[49];
  // last detect call (for the read-back entry points)
  int last_n = 0, last_chunk_start = 0, last_chunk_n = 0, last_cap = 0;
  const uint8_t* last_frames = nullptr; size_t last_stride = 0; int last_pitch = 0;
  bool last_outputs_ctx = false;
  int launches = 0;
  // chunk pipeline: staging copies and result copies run on their own streams
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_start = nullptr, ev_chain = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  // frame ingest: pinned host area the decode threads fill (same layout as d_frames)
  uint8_t* h_ingest = nullptr; size_t h_ingest_bytes = 0;
  // device decode: compressed streams (pinned + device), inflated scanlines, per-frame descriptors and status
  uint8_t* h_comp = nullptr; uint8_t* d_comp = nullptr; uint8_t* d_raw = nullptr;
  size_t comp_slot = 0, raw_slot = 0; int ingest_cap = 0;
  orbk::InflateDesc* h_descs = nullptr; orbk::InflateDesc* d_descs = nullptr;
  int* h_inf_status = nullptr; int* d_inf_status = nullptr; uint32_t* d_adler = nullptr;
  static constexpr int N_INGEST = 4;
  cudaStream_t s_ingest[N_INGEST] = {nullptr, nullptr, nullptr, nullptr};
  // Lucas-Kanade tracker: two packed pyramids + point arrays
  uint8_t* d_lk = nullptr; size_t d_lk_bytes = 0;
  int lk_top = -1, lk_w[orbk::LK_MAX_LEVELS], lk_h[orbk::LK_MAX_LEVELS]; size_t lk_ofs[orbk::LK_MAX_LEVELS + 1], lk_pyr = 0;
  // optional per-kernel event timing
  bool profiling = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans; size_t spans_used = 0;
};

namespace {

int fail(orb_ctx* c, int code, const char* fmt, ...) {
  char* dst = c ? c->err : g_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, ORB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                              \
  } while (0)

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
  int tile = 0, atile = 0, kept = 0, xo = 0, yo = 0, eo = 0;
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
    G.scale = level_scale(p.scale_factor, l);
    G.lvl_ofs = lv; if (l > 0) lv += (unsigned long long)align_up(G.h * G.pitch, 256);
    G.box_ofs = bx; bx += (unsigned long long)align_up((G.h + 1) * G.bpitch, 128);
    G.cand_ofs = cd; cd += (unsigned long long)align_up(G.cand_cap, 32);
  }
  P->tiles_per_frame = tile; P->a_tiles_per_frame = atile; P->kept_per_frame = kept; P->edge_frame_elems = eo;
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
  }
  CK(cudaGetLastError());
  ctx->launches += 1;
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
    orbk::k_select<<<grid, orbk::K2_THREADS, (size_t)(npow2 + orbk::K2_SMEM_KEYS) * 8, ctx->stream>>>(P, B, npow2);
  }
  CK(cudaGetLastError());
  ctx->launches += 1;
  return ORB_OK;
}
int launch_describe(orb_ctx* ctx, const OrbPlan& P, const Bufs& B, const DescribeJob& J, int nwarps, int nframes) {
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

void stage_plan(orb_ctx* ctx, int w, int h, int policy, int quota, OrbPlan* P, Bufs* B) {
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
  B->out_kps = ctx->d_kps; B->out_angles = ctx->d_angles; B->out_desc = ctx->d_desc; B->out_n = ctx->d_nout;
  B->out_cap = ctx->list_cap;
}

}  // namespace

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
                  ctx->d_list_kps, ctx->d_list_angles, ctx->d_list_out};
  for (void* q : ptrs) if (q) cudaFree(q);
  for (auto& sp : ctx->spans) { if (sp.a) cudaEventDestroy(sp.a); if (sp.b) cudaEventDestroy(sp.b); }
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->h_ingest) cudaFreeHost(ctx->h_ingest);
  if (ctx->d_lk) cudaFree(ctx->d_lk);
  if (ctx->h_comp) cudaFreeHost(ctx->h_comp);
  if (ctx->h_descs) cudaFreeHost(ctx->h_descs);
  if (ctx->h_inf_status) cudaFreeHost(ctx->h_inf_status);
  if (ctx->d_comp) cudaFree(ctx->d_comp);
  if (ctx->d_raw) cudaFree(ctx->d_raw);
  if (ctx->d_descs) cudaFree(ctx->d_descs);
  if (ctx->d_inf_status) cudaFree(ctx->d_inf_status);
  if (ctx->d_adler) cudaFree(ctx->d_adler);
  for (cudaStream_t q : ctx->s_ingest) if (q) cudaStreamDestroy(q);
  for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_done) cudaEventDestroy(e);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
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
    // measured on B200: larger waves win (launch gaps and kernel tails outweigh L2 residency of the scratch): 512 frames
    // per wave when frames and results stay on the device, 128 when they are staged from / to the host so that the
    // copies of one wave hide behind the kernels of another (bounded by 3 GB of scratch)
    int chunk = p->chunk_frames > 0 ? p->chunk_frames : (int)std::max<size_t>(1, std::min<size_t>(512, ((size_t)3 << 30) / per_frame));
    ctx->chunk_staged = p->chunk_frames > 0 ? p->chunk_frames : 128;
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
    CK(cudaFuncSetAttribute(orbk::k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (ORB_SORT_CAP + orbk::K2_SMEM_KEYS) * 8));
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

// How the frames of a wave reach the staging area when they do not come from a caller buffer (frame ingest):
// stage() runs on the calling thread just before the wave's kernels are queued, may block on host work, and queues on
// ctx->s_h2d whatever brings frames [c0, c0 + nc) into ctx->d_frames.
struct WaveSource {
  // queues the work for frames [c0, c0 + nc) of wave ci and records `ready` behind it (on whichever stream it used)
  virtual int stage(orb_ctx* ctx, int ci, int c0, int nc, cudaEvent_t ready) = 0;
  virtual int preferred_wave(const orb_ctx*) const { return 0; }      // 0 = the context's staged wave size
  virtual ~WaveSource() {}
};

static int run_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h,
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

  int total_quota = 0;
  for (int l = 0; l < P.nlevels; l++) total_quota += P.lv[l].quota;
  const int nwarps = std::min(total_quota, cap);
  // wave boundaries: full waves of `chunk` frames; when frames are staged from the host the first waves are short
  // (chunk/8, chunk/4, chunk/2) so that the kernels start while most of the batch is still in flight over PCIe
  std::vector<int> wave_begin;
  {
    int wave = (direct && outputs_on_device) ? ctx->chunk : ctx->chunk_staged;
    if (source && source->preferred_wave(ctx) > 0) wave = std::min(ctx->chunk, source->preferred_wave(ctx));
    int c0 = 0, ramp = (!direct && n_frames > wave) ? std::max(1, wave / 8) : wave;
    while (c0 < n_frames) {
      wave_begin.push_back(c0);
      c0 += std::min(ramp, wave);
      if (ramp < wave) ramp *= 2;
    }
    wave_begin.push_back(n_frames);
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
    B.frames = src + (size_t)c0 * stride; B.frame_stride = stride; B.pitch0 = sp;
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

int orb_detect_and_compute_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h,
                                 size_t pitch, size_t frame_stride, int cap, orb_keypoint* kps, float* angles,
                                 orb_descriptor* desc, int* n_out, int outputs_on_device) {
  if (ctx && !frames) return fail(ctx, ORB_E_INVALID, "null buffer");
  return run_batch(ctx, frames, frames_on_device, n_frames, w, h, pitch, frame_stride, cap, kps, angles, desc, n_out,
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
  stage_plan(ctx, w, h, ORB_SELECT_RASTER_FIRST_N, nfeatures, &P, &B);
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
  stage_plan(ctx, w, h, ORB_SELECT_RASTER_FIRST_N, 0, &P, &B);
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

// ---- descriptor matching -------------------------------------------------------------------------
static int match_launch(orb_ctx* ctx, const orb_descriptor* dq, const orb_descriptor* dt, const int* dn, int nq, int nt, int npairs,
                        long long sq, long long st, long long so, orb_match* dout) {
  if (npairs <= 0 || nq <= 0) return ORB_OK;
  dim3 grid((nq + orbk::M_THREADS - 1) / orbk::M_THREADS, npairs);
  orbk::k_match<<<grid, orbk::M_THREADS, 0, ctx->stream>>>(dq, dt, dn, nq, nt, sq, st, so, dout);
  CK(cudaGetLastError());
  ctx->launches += 1;
  return ORB_OK;
}

int orb_match_knn2(orb_ctx* ctx, const orb_descriptor* query, int nq, const orb_descriptor* train, int nt, int on_device,
                   orb_match* out) {
  if (!ctx) return ORB_E_INVALID;
  if (nq < 0 || nt < 0 || (nq > 0 && (!query || !out)) || (nt > 0 && !train)) return fail(ctx, ORB_E_INVALID, "bad match arguments");
  if (nq == 0) return ORB_OK;
  CK(cudaSetDevice(ctx->p.device));
  if (on_device) return match_launch(ctx, query, train, nullptr, nq, nt, 1, 0, 0, 0, out);
  orb_descriptor *dq = nullptr, *dt = nullptr; orb_match* dm = nullptr;
  CK(cudaMalloc(&dq, sizeof(orb_descriptor) * (size_t)nq));
  CK(cudaMalloc(&dt, sizeof(orb_descriptor) * (size_t)std::max(nt, 1)));
  CK(cudaMalloc(&dm, sizeof(orb_match) * (size_t)nq));
  CK(cudaMemcpyAsync(dq, query, sizeof(orb_descriptor) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
  if (nt) CK(cudaMemcpyAsync(dt, train, sizeof(orb_descriptor) * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
  int rc = match_launch(ctx, dq, dt, nullptr, nq, nt, 1, 0, 0, 0, dm);
  if (!rc) {
    CK(cudaMemcpyAsync(out, dm, sizeof(orb_match) * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(dq); cudaFree(dt); cudaFree(dm);
  return rc;
}

int orb_match_knn2_batch(orb_ctx* ctx, const orb_descriptor* desc, const int* n, int n_frames, int cap, int on_device,
                         orb_match* out) {
  if (!ctx) return ORB_E_INVALID;
  if (!desc || !n || !out || n_frames < 1 || cap < 1) return fail(ctx, ORB_E_INVALID, "bad match arguments");
  if (n_frames < 2) return ORB_OK;
  CK(cudaSetDevice(ctx->p.device));
  const long long s = cap;
  if (on_device) return match_launch(ctx, desc, desc + cap, n, cap, cap, n_frames - 1, s, s, s, out);
  orb_descriptor* dd = nullptr; int* dn = nullptr; orb_match* dm = nullptr;
  const size_t nd = (size_t)n_frames * cap, nm = (size_t)(n_frames - 1) * cap;
  CK(cudaMalloc(&dd, sizeof(orb_descriptor) * nd));
  CK(cudaMalloc(&dn, sizeof(int) * n_frames));
  CK(cudaMalloc(&dm, sizeof(orb_match) * nm));
  CK(cudaMemcpyAsync(dd, desc, sizeof(orb_descriptor) * nd, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dn, n, sizeof(int) * n_frames, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dm, out, sizeof(orb_match) * nm, cudaMemcpyHostToDevice, ctx->stream));   // entries >= n[p] stay as they were
  int rc = match_launch(ctx, dd, dd + cap, dn, cap, cap, n_frames - 1, s, s, s, dm);
  if (!rc) {
    CK(cudaMemcpyAsync(out, dm, sizeof(orb_match) * nm, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(dd); cudaFree(dn); cudaFree(dm);
  return rc;
}

void orb_ratio_test(const orb_match* m, int n, float ratio, uint8_t* keep) {
  // if (m.distance < 0.8 * n.distance), reference src/feature_matching.cpp:178 (float distances, double product)
  for (int i = 0; i < n; i++)
    keep[i] = m[i].idx2 >= 0 && (double)(float)m[i].dist1 < (double)ratio * (double)(float)m[i].dist2;
}

int orb_debug_eval_math(orb_ctx* ctx, int op, const float* a, const float* b, int n, float* out) {
  if (!ctx || !a || !out || n < 0 || op < 0 || op > 3 || (op == 0 && !b)) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  float *da = nullptr, *db = nullptr, *dout = nullptr;
  CK(cudaMalloc(&da, sizeof(float) * std::max(n, 1)));
  CK(cudaMalloc(&db, sizeof(float) * std::max(n, 1)));
  CK(cudaMalloc(&dout, sizeof(float) * std::max(n, 1)));
  CK(cudaMemcpy(da, a, sizeof(float) * n, cudaMemcpyHostToDevice));
  if (b) CK(cudaMemcpy(db, b, sizeof(float) * n, cudaMemcpyHostToDevice));
  if (n > 0) orbk::k_eval_math<<<(n + 255) / 256, 256, 0, ctx->stream>>>(op, da, db, n, dout);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out, dout, sizeof(float) * n, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return ORB_OK;
}


}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Frame ingest: PNG files -> staging area -> the wave pipeline above.
namespace {

// whole file into `buf`; returns false with errno-style message
bool read_file(const char* path, std::vector<uint8_t>* buf, std::string* err) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { *err = std::string("cannot open ") + path; return false; }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); *err = std::string("cannot stat ") + path; return false; }
  buf->resize((size_t)st.st_size);
  size_t got = 0;
  while (got < buf->size()) {
    const ssize_t r = read(fd, buf->data() + got, buf->size() - got);
    if (r <= 0) break;
    got += (size_t)r;
  }
  close(fd);
  if (got != buf->size()) { *err = std::string("short read on ") + path; return false; }
  return true;
}

// Host-decode source: a pool of threads decodes the files in order straight into the pinned area; stage() waits for
// the frames of its wave and queues their copy.
struct HostDecodeSource : WaveSource {
  const char* const* paths; int n, w, h;
  uint8_t* area; size_t slot; int pitch;
  std::vector<std::thread> pool;
  std::atomic<int> next{0};
  std::atomic<bool> stop{false};
  std::mutex mu; std::condition_variable cv;
  std::vector<uint8_t> done;
  int ready = 0;                 // frames [0, ready) are decoded (guarded by mu)
  int err_code = 0; std::string err_msg;

  void work() {
    std::vector<uint8_t> file;
    orbpng::Scratch scratch;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || stop.load()) return;
      std::string msg;
      int code = 0;
      if (!read_file(paths[i], &file, &msg)) code = ORB_E_IO;
      else {
        const char* e = orbpng::decode_gray8(file.data(), file.size(), area + (size_t)i * slot, (size_t)pitch, w, h, &scratch);
        if (e) { code = ORB_E_FORMAT; msg = std::string(paths[i]) + ": " + e; }
      }
      std::lock_guard<std::mutex> lk(mu);
      if (code && !err_code) { err_code = code; err_msg = msg; stop.store(true); }
      done[i] = 1;
      cv.notify_all();
    }
  }
  void start(int n_threads) {
    done.assign(n, 0);
    for (int t = 0; t < n_threads; t++) pool.emplace_back([this] { work(); });
  }
  int stage(orb_ctx* ctx, int, int c0, int nc, cudaEvent_t ready_ev) override {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] {
        while (ready < n && done[ready]) ready++;
        return err_code != 0 || ready >= c0 + nc;
      });
      if (err_code) return fail(ctx, err_code, "%s", err_msg.c_str());
    }
    CK(cudaMemcpyAsync(ctx->d_frames + (size_t)c0 * slot, area + (size_t)c0 * slot, slot * nc, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaEventRecord(ready_ev, ctx->s_h2d));
    return ORB_OK;
  }
  ~HostDecodeSource() override {
    stop.store(true);
    for (auto& t : pool) t.join();
  }
};

const char* inflate_status_text(int st) {
  switch (st) {
    case orbk::INF_CORRUPT: return "corrupt deflate data";
    case orbk::INF_SIZE: return "image data does not match the frame size";
    case orbk::INF_TRUNCATED: return "truncated deflate data";
    case orbk::INF_TABLE: return "Huffman table larger than the device decoder holds";
    case orbk::INF_FILTER: return "unknown PNG filter type";
    case orbk::INF_CHECKSUM: return "zlib: incorrect data check";
    default: return "unknown decode error";
  }
}

// (re)allocates the device-decode areas for n frames of w x h
int ensure_device_decode(orb_ctx* ctx, int n, int w, int h) {
  const size_t raw = ((size_t)(w + 1) * h + orbk::UNF_LEAD + 48 + 15) / 16 * 16;
  const size_t comp = (raw + raw / 64 + 1024 + 511) / 512 * 512 + 512;
  if (n <= ctx->ingest_cap && raw <= ctx->raw_slot && comp <= ctx->comp_slot) return ORB_OK;
  if (ctx->h_comp) cudaFreeHost(ctx->h_comp);
  if (ctx->h_descs) cudaFreeHost(ctx->h_descs);
  if (ctx->h_inf_status) cudaFreeHost(ctx->h_inf_status);
  cudaFree(ctx->d_comp); cudaFree(ctx->d_raw); cudaFree(ctx->d_descs); cudaFree(ctx->d_inf_status); cudaFree(ctx->d_adler);
  ctx->d_adler = nullptr;
  ctx->h_comp = nullptr; ctx->h_descs = nullptr; ctx->h_inf_status = nullptr;
  ctx->d_comp = ctx->d_raw = nullptr; ctx->d_descs = nullptr; ctx->d_inf_status = nullptr;
  ctx->ingest_cap = 0;
  const int cap = std::max(n, ctx->p.max_batch);
  CK(cudaHostAlloc((void**)&ctx->h_comp, comp * cap, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&ctx->h_descs, sizeof(orbk::InflateDesc) * cap, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&ctx->h_inf_status, sizeof(int) * cap, cudaHostAllocDefault));
  CK(cudaMalloc((void**)&ctx->d_comp, comp * cap));
  CK(cudaMalloc((void**)&ctx->d_raw, raw * cap));
  CK(cudaMalloc((void**)&ctx->d_descs, sizeof(orbk::InflateDesc) * cap));
  CK(cudaMalloc((void**)&ctx->d_inf_status, sizeof(int) * cap));
  CK(cudaMalloc((void**)&ctx->d_adler, sizeof(uint32_t) * cap));
  CK(cudaMemset(ctx->d_comp, 0, comp * cap));
  // the inflate kernels are latency chains (one busy lane per warp): give their blocks the first free slots
  int prio_least = 0, prio_greatest = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  for (cudaStream_t& q : ctx->s_ingest) if (!q) CK(cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, prio_greatest));
  ctx->comp_slot = comp; ctx->raw_slot = raw; ctx->ingest_cap = cap;
  return ORB_OK;
}

// Device-decode source: host threads only read the files, check the framing and lay the deflate streams out in the pinned
// area; stage() uploads the compressed bytes of a wave and queues k_inflate + k_unfilter on the copy stream.
struct DeviceDecodeSource : WaveSource {
  const char* const* paths; int n, w, h;
  orb_ctx* ctx;
  std::vector<std::thread> pool;
  std::atomic<int> next{0};
  std::atomic<bool> stop{false};
  std::mutex mu; std::condition_variable cv;
  std::vector<uint8_t> done;
  std::vector<uint32_t> bytes;
  int ready = 0;
  int err_code = 0; std::string err_msg;

  void work() {
    std::vector<uint8_t> file;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || stop.load()) return;
      std::string msg;
      int code = 0;
      if (!read_file(paths[i], &file, &msg)) code = ORB_E_IO;
      else {
        orbpng::Info I;
        size_t nb = 0;
        const char* e = orbpng::extract_deflate(file.data(), file.size(), &I, ctx->h_comp + (size_t)i * ctx->comp_slot,
                                                ctx->comp_slot - 512, &nb);
        if (!e && (I.width != w || I.height != h)) e = "png: image size differs from the expected frame size";
        if (!e && (I.color_type != 0 || I.bit_depth != 8)) e = "png: the device decoder takes 8-bit gray files only";
        if (e) { code = ORB_E_FORMAT; msg = std::string(paths[i]) + ": " + e; }
        bytes[i] = (uint32_t)nb;
      }
      std::lock_guard<std::mutex> lk(mu);
      if (code && !err_code) { err_code = code; err_msg = msg; stop.store(true); }
      done[i] = 1;
      cv.notify_all();
    }
  }
  void start(int n_threads) {
    done.assign(n, 0);
    bytes.assign(n, 0);
    for (int t = 0; t < n_threads; t++) pool.emplace_back([this] { work(); });
  }
  // Inflate waves are independent of the ORB waves: a deflate stream is serial (one warp, ~15 ms per KITTI frame), so
  // the decoder wants as many streams in flight as the host has read, while the ORB kernels want short waves that
  // start early.  Inflate wave k covers frames [k * IW, (k + 1) * IW) and runs on ingest stream k % N_INGEST, so
  // consecutive inflate waves overlap; an ORB wave waits for the inflate wave that holds its last frame.
  static constexpr int IW = 256;
  int preferred_wave(const orb_ctx* c) const override { return c->chunk; }   // frames are resident before the ORB waves start
  int inflated = 0, inflate_waves = 0;
  cudaStream_t last_q = nullptr;
  int launch_inflate(orb_ctx* ctx, int f0, int nf) {
    cudaStream_t q = ctx->s_ingest[inflate_waves % orb_ctx::N_INGEST];
    if (inflate_waves < orb_ctx::N_INGEST) CK(cudaStreamWaitEvent(q, ctx->ev_start, 0));
    const uint32_t out_bytes = (uint32_t)((size_t)(w + 1) * h);
    // one strided copy for the wave: every slot up to the longest stream in it (stream + Adler trailer + zero pad)
    size_t up = 0;
    for (int i = f0; i < f0 + nf; i++) up = std::max(up, ((size_t)bytes[i] + 4 + 16 + 15) / 16 * 16);
    CK(cudaMemcpy2DAsync(ctx->d_comp + (size_t)f0 * ctx->comp_slot, ctx->comp_slot, ctx->h_comp + (size_t)f0 * ctx->comp_slot,
                         ctx->comp_slot, up, nf, cudaMemcpyHostToDevice, q));
    for (int i = f0; i < f0 + nf; i++) {
      ctx->h_descs[i] = orbk::InflateDesc{ctx->d_comp + (size_t)i * ctx->comp_slot, bytes[i], out_bytes, ctx->d_raw + (size_t)i * ctx->raw_slot + orbk::UNF_LEAD};
    }
    CK(cudaMemcpyAsync(ctx->d_descs + f0, ctx->h_descs + f0, sizeof(orbk::InflateDesc) * nf, cudaMemcpyHostToDevice, q));
    orbk::k_inflate<<<nf, 32, 0, q>>>(ctx->d_descs + f0, ctx->d_inf_status + f0, ctx->d_adler + f0);
    orbk::k_unfilter<<<(nf + orbk::UNF_WARPS - 1) / orbk::UNF_WARPS, orbk::UNF_WARPS * 32, 0, q>>>(
        ctx->d_raw + (size_t)f0 * ctx->raw_slot, ctx->raw_slot, ctx->d_frames + (size_t)f0 * ctx->frames_slot_bytes,
        ctx->frames_slot_bytes, ctx->frames_pitch, w, h, nf, ctx->d_inf_status + f0, ctx->d_adler + f0);
    CK(cudaGetLastError());
    ctx->launches += 2;
    inflate_waves++;
    last_q = q;
    return ORB_OK;
  }
  int stage(orb_ctx* ctx, int, int c0, int nc, cudaEvent_t ready_ev) override {
    cudaStream_t also[8];
    int n_also = 0;
    // Measured on B200: ORB kernels and inflate kernels sharing the SMs slow each other far more than the overlap gains
    // (126 ms vs 69 ms per 1024 KITTI frames), so every frame is inflated -- in waves that follow the host reads -- before
    // the first ORB wave starts.  ORB_INGEST_SEPARATE=0 restores the interleaved schedule for experiments.
    static const bool separate = !(getenv("ORB_INGEST_SEPARATE") && atoi(getenv("ORB_INGEST_SEPARATE")) == 0);
    const int want = separate ? n : c0 + nc;
    while (inflated < want) {
      if (last_q && n_also < 8 && inflated > c0) also[n_also++] = last_q;     // this ORB wave spans several inflate waves
      const int f0 = inflated, nf = std::min(IW, n - f0);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] {
          while (ready < n && done[ready]) ready++;
          return err_code != 0 || ready >= f0 + nf;
        });
        if (err_code) return fail(ctx, err_code, "%s", err_msg.c_str());
      }
      const int rc = launch_inflate(ctx, f0, nf);
      if (rc) return rc;
      inflated = f0 + nf;
    }
    // frames of this ORB wave come from the inflate wave queued last at the latest (earlier ORB waves have waited for
    // the earlier inflate waves, except for those first needed by this very wave)
    for (int k = 0; k < n_also; k++) {
      if (also[k] == last_q) continue;
      CK(cudaEventRecord(ctx->ev_chain, also[k]));
      CK(cudaStreamWaitEvent(last_q, ctx->ev_chain, 0));
    }
    CK(cudaEventRecord(ready_ev, last_q));
    return ORB_OK;
  }
  ~DeviceDecodeSource() override {
    stop.store(true);
    for (auto& t : pool) t.join();
  }
};

int png_status(const char* e) {
  if (!e) return ORB_OK;
  snprintf(g_create_error, sizeof(g_create_error), "%s", e);
  return ORB_E_FORMAT;
}

}  // namespace

extern "C" {

int orb_png_info(const uint8_t* file, size_t file_bytes, orb_image_info* info) {
  if (!file || !info) return ORB_E_INVALID;
  orbpng::Info I;
  const int rc = png_status(orbpng::read_info(file, file_bytes, &I));
  if (rc) return rc;
  info->width = I.width; info->height = I.height; info->bit_depth = I.bit_depth; info->channels = I.channels;
  return ORB_OK;
}

int orb_png_decode_gray8(const uint8_t* file, size_t file_bytes, uint8_t* dst, size_t pitch, int w, int h) {
  if (!file || !dst) return ORB_E_INVALID;
  return png_status(orbpng::decode_gray8(file, file_bytes, dst, pitch, w, h, nullptr));
}

int orb_imread_gray8(const char* path, uint8_t* dst, size_t pitch, int cap_w, int cap_h, int* w, int* h) {
  if (!path || !dst) return ORB_E_INVALID;
  std::vector<uint8_t> file;
  std::string msg;
  if (!read_file(path, &file, &msg)) { snprintf(g_create_error, sizeof(g_create_error), "%s", msg.c_str()); return ORB_E_IO; }
  orbpng::Info I;
  int rc = png_status(orbpng::read_info(file.data(), file.size(), &I));
  if (rc) return rc;
  if (w) *w = I.width;
  if (h) *h = I.height;
  if (I.width > cap_w || I.height > cap_h) { snprintf(g_create_error, sizeof(g_create_error), "%s: image %dx%d exceeds the buffer", path, I.width, I.height); return ORB_E_CAPACITY; }
  return png_status(orbpng::decode_gray8(file.data(), file.size(), dst, pitch, I.width, I.height, nullptr));
}

int orb_detect_and_compute_files(orb_ctx* ctx, const char* const* paths, int n_frames, int n_threads, int decode_on_device,
                                 int cap, orb_keypoint* kps, float* angles, orb_descriptor* desc, int* n_out,
                                 int outputs_on_device) {
  if (!ctx) return ORB_E_INVALID;
  if (!paths || n_frames < 1) return fail(ctx, ORB_E_INVALID, "no frame files");
  if (n_frames > ctx->p.max_batch) return fail(ctx, ORB_E_CAPACITY, "batch %d exceeds max_batch %d", n_frames, ctx->p.max_batch);
  CK(cudaSetDevice(ctx->p.device));
  // the first file fixes the frame size of the call
  int w = 0, h = 0;
  {
    uint8_t head[64];
    const int fd = open(paths[0], O_RDONLY);
    const ssize_t got = fd >= 0 ? read(fd, head, sizeof(head)) : -1;
    if (fd >= 0) close(fd);
    if (got < 33) return fail(ctx, ORB_E_IO, "cannot read %s", paths[0]);
    orbpng::Info I;
    const char* e = orbpng::read_info(head, (size_t)got, &I);
    if (e) return fail(ctx, ORB_E_FORMAT, "%s: %s", paths[0], e);
    w = I.width; h = I.height;
  }
  int rc = get_plan(ctx, w, h);
  if (rc) return rc;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = std::min(n_threads, n_frames);
  if (decode_on_device) {
    if ((rc = ensure_device_decode(ctx, n_frames, w, h))) return rc;
    DeviceDecodeSource src;
    src.paths = paths; src.n = n_frames; src.w = w; src.h = h; src.ctx = ctx;
    src.start(n_threads);
    rc = run_batch(ctx, nullptr, 0, n_frames, w, h, 0, 0, cap, kps, angles, desc, n_out, outputs_on_device, &src);
    if (rc) return rc;
    // decode failures are per-frame flags on the device: the results of a failed frame are meaningless, report it
    for (cudaStream_t q : ctx->s_ingest) CK(cudaStreamSynchronize(q));
    CK(cudaMemcpy(ctx->h_inf_status, ctx->d_inf_status, sizeof(int) * n_frames, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_frames; i++)
      if (ctx->h_inf_status[i]) return fail(ctx, ORB_E_FORMAT, "%s: %s", paths[i], inflate_status_text(ctx->h_inf_status[i]));
    return ORB_OK;
  }
  const size_t need = ctx->frames_slot_bytes * (size_t)n_frames;
  if (need > ctx->h_ingest_bytes) {
    if (ctx->h_ingest) CK(cudaFreeHost(ctx->h_ingest));
    ctx->h_ingest = nullptr; ctx->h_ingest_bytes = 0;
    CK(cudaHostAlloc((void**)&ctx->h_ingest, need, cudaHostAllocDefault));
    ctx->h_ingest_bytes = need;
    memset(ctx->h_ingest, 0, need);
  }
  HostDecodeSource src;
  src.paths = paths; src.n = n_frames; src.w = w; src.h = h;
  src.area = ctx->h_ingest; src.slot = ctx->frames_slot_bytes; src.pitch = ctx->frames_pitch;
  src.start(n_threads);
  return run_batch(ctx, nullptr, 0, n_frames, w, h, 0, 0, cap, kps, angles, desc, n_out, outputs_on_device, &src);
}

int orb_get_ingested_frame(orb_ctx* ctx, int frame, uint8_t* dst, size_t dst_pitch, int* w, int* h) {
  if (!ctx || !dst) return ORB_E_INVALID;
  if (!ctx->plan_valid || ctx->last_frames != ctx->d_frames || frame < 0 || frame >= ctx->last_n)
    return fail(ctx, ORB_E_INVALID, "frame %d is not in the staging area", frame);
  const int W = ctx->plan.lv[0].w, H = ctx->plan.lv[0].h;
  if (dst_pitch < (size_t)W) return fail(ctx, ORB_E_INVALID, "dst_pitch too small");
  CK(cudaSetDevice(ctx->p.device));
  CK(cudaMemcpy2DAsync(dst, dst_pitch, ctx->d_frames + (size_t)frame * ctx->frames_slot_bytes, ctx->frames_pitch, W, H,
                       cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (w) *w = W;
  if (h) *h = H;
  return ORB_OK;
}


int orb_debug_inflate(orb_ctx* ctx, const uint8_t* streams, const uint32_t* offsets, int n, uint8_t* out,
                      const uint32_t* out_offsets, int* status) {
  if (!ctx || !streams || !offsets || !out || !out_offsets || !status || n < 1) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  // every stream in its own zero-padded, 512-byte aligned slot; outputs 16-byte aligned
  std::vector<size_t> in_at(n + 1, 0), out_at(n + 1, 0);
  for (int i = 0; i < n; i++) {
    in_at[i + 1] = in_at[i] + ((size_t)(offsets[i + 1] - offsets[i]) + 16 + 511) / 512 * 512 + 512;
    out_at[i + 1] = out_at[i] + ((size_t)(out_offsets[i + 1] - out_offsets[i]) + 32 + 15) / 16 * 16;
  }
  std::vector<uint8_t> packed(in_at[n], 0);
  for (int i = 0; i < n; i++) memcpy(packed.data() + in_at[i], streams + offsets[i], offsets[i + 1] - offsets[i]);
  uint8_t *d_in = nullptr, *d_out = nullptr;
  orbk::InflateDesc* d_desc = nullptr;
  int* d_st = nullptr;
  CK(cudaMalloc((void**)&d_in, in_at[n]));
  CK(cudaMalloc((void**)&d_out, out_at[n]));
  CK(cudaMalloc((void**)&d_desc, sizeof(orbk::InflateDesc) * n));
  CK(cudaMalloc((void**)&d_st, sizeof(int) * n));
  std::vector<orbk::InflateDesc> descs(n);
  for (int i = 0; i < n; i++)
    descs[i] = orbk::InflateDesc{d_in + in_at[i], offsets[i + 1] - offsets[i], out_offsets[i + 1] - out_offsets[i], d_out + out_at[i]};
  CK(cudaMemcpy(d_in, packed.data(), in_at[n], cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_desc, descs.data(), sizeof(orbk::InflateDesc) * n, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xEE, out_at[n]));
  orbk::k_inflate<<<n, 32, 0, ctx->stream>>>(d_desc, d_st, nullptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(status, d_st, sizeof(int) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++)
    CK(cudaMemcpy(out + out_offsets[i], d_out + out_at[i], out_offsets[i + 1] - out_offsets[i], cudaMemcpyDeviceToHost));
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_desc); cudaFree(d_st);
  return ORB_OK;
}


// ---------------------------------------------------------------------------------------------
// Pyramidal Lucas-Kanade
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
  ctx->lk_top = top; ctx->lk_pyr = pyr;
  for (int l = 0; l <= top; l++) { ctx->lk_w[l] = lw[l]; ctx->lk_h[l] = lh[l]; ctx->lk_ofs[l] = lvl_ofs[l]; }
  ctx->launches = 0;
  for (int l = 1; l <= top; l++) {
    const dim3 blk(32, 8), grd((lw[l] + 31) / 32, (lh[l] + 7) / 8);
    orbk::k_lk_pyrdown<<<grd, blk, 0, q>>>(P.img[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], dP + lvl_ofs[l], lw[l], lh[l]);
    orbk::k_lk_pyrdown<<<grd, blk, 0, q>>>(N.img[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], dN + lvl_ofs[l], lw[l], lh[l]);
    ctx->launches += 2;
  }
  orbk::k_lk_track<<<(n + orbk::LK_WARPS - 1) / orbk::LK_WARPS, orbk::LK_WARPS * 32, 0, q>>>(
      P, N, top, d_prev, n, win, max_iter, eps2, min_eig, d_next, d_status, err ? d_err : nullptr);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(next_pts, d_next, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, q));
  CK(cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, q));
  if (err) CK(cudaMemcpyAsync(err, d_err, sizeof(float) * n, cudaMemcpyDeviceToHost, q));
  CK(cudaStreamSynchronize(q));
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
