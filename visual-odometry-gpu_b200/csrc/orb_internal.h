// orb_internal.h -- what the translation units of liborb_b200.so share: the context, error plumbing and the wave pipeline
// hook.  Not installed; the public surface is include/orb_b200.h.
//   orb_api.cu    : arena, plan, ORB launch pipeline, stage entry points, matcher      (kernels: orb_kernels.cuh)
//   orb_ingest.cu : PNG ingest, host and device decode                                 (kernels: orb_ingest_kernels.cuh)
//   orb_lk.cu     : pyramidal Lucas-Kanade tracker                                     (kernels: orb_lk_kernels.cuh)
#pragma once
#include <cuda.h>            // CUtensorMap only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/orb_b200.h"
#include "orb_plan.h"

namespace orbk { struct InflateDesc; constexpr int LK_MAX_LEVELS = 8; }

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
  int* d_edge2 = nullptr;
  void* d_scratch[3] = {nullptr, nullptr, nullptr}; size_t scratch_bytes[3] = {0, 0, 0};   // host-buffer matcher / debug calls (grow-only)
  int8_t* d_match_exp = nullptr; size_t match_exp_bytes = 0;   // tensor-core matcher: +-8 int8 rows (the tensor maps travel as kernel parameters)
  int sm_count = 0;                                      // persistent kernels: one CTA per SM
  float* d_scores = nullptr; size_t d_scores_bytes = 0;   // orb_nms_scores: the caller's score map (grow-only)
  int* d_cand_count = nullptr; size_t zero_bytes_per_frame = 0; uint32_t* d_kept_xy = nullptr; float* d_kept_r = nullptr; int* d_kept_count = nullptr;
  OrbTap *d_xtab = nullptr, *d_ytab = nullptr;
  uint32_t *d_tile_a = nullptr, *d_tile_b = nullptr, *d_tile_b1 = nullptr; int tile_a_cap = 0, tile_b_cap = 0;
  float* d_harris_w = nullptr; float4* d_pattern = nullptr; int* d_flags = nullptr; int* h_flags = nullptr;
  orb_keypoint* d_kps = nullptr; float* d_angles = nullptr; orb_descriptor* d_desc = nullptr; int* d_nout = nullptr;
  orb_keypoint* d_side_xy = nullptr; int* d_side_level = nullptr; float* d_side_resp = nullptr;
  orb_keypoint* d_list_kps = nullptr; float* d_list_angles = nullptr; float* d_list_out = nullptr; int list_cap = 0;
  float harris_w[49];
  // TMA tensor maps (orb_kernels.cuh: TM_*): device copy, host mirror and the geometry they were encoded for
  CUtensorMap* d_tmaps = nullptr; CUtensorMap h_tmaps[3 * ORB_MAX_LEVELS];
  void* tmap_encode = nullptr;     // cuTensorMapEncodeTiled
  struct TmapKey { const void* base; size_t pitch, stride; int n, W, H, nlevels, patch_radius; } tmap_key = {nullptr, 0, 0, 0, 0, 0, 0, 0};
  // last detect call (for the read-back entry points)
  int last_n = 0, last_chunk_start = 0, last_chunk_n = 0, last_cap = 0;
  const uint8_t* last_frames = nullptr; size_t last_stride = 0; int last_pitch = 0;
  bool last_outputs_ctx = false;
  int launches = 0;
  // chunk pipeline: staging copies and result copies run on their own streams
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_start = nullptr, ev_chain = nullptr;
  cudaStream_t s_side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool edges_pending = false;   // k_edges next to k_harris / k_select
  std::vector<cudaEvent_t> ev_in, ev_done;
  // frame ingest: pinned host area the decode threads fill (same layout as d_frames)
  uint8_t* h_ingest = nullptr; size_t h_ingest_bytes = 0;
  // device decode: compressed streams (pinned + device), inflated scanlines, per-frame descriptors and status
  uint8_t* h_comp = nullptr; uint8_t* d_comp = nullptr; uint8_t* d_raw = nullptr;
  size_t comp_slot = 0, raw_slot = 0; int ingest_cap = 0;
  orbk::InflateDesc* h_descs = nullptr; orbk::InflateDesc* d_descs = nullptr;
  int* h_inf_status = nullptr; int* d_inf_status = nullptr; uint32_t* d_adler = nullptr;
  uint32_t* h_crc = nullptr; uint32_t* d_crc = nullptr; int* h_crc_n = nullptr; int* d_crc_n = nullptr;   // IDAT chunk checksums
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

extern thread_local char g_orb_create_error[512];

inline int orb_fail(orb_ctx* c, int code, const char* fmt, ...) {
  char* dst = c ? c->err : g_orb_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}
#define fail orb_fail
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, ORB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                              \
  } while (0)

// How the frames of a wave reach the staging area when they do not come from a caller buffer (frame ingest):
// stage() runs on the calling thread just before the wave's kernels are queued, may block on host work, and queues on
// ctx->s_h2d whatever brings frames [c0, c0 + nc) into ctx->d_frames.
struct WaveSource {
  // queues the work for frames [c0, c0 + nc) of wave ci and records `ready` behind it (on whichever stream it used)
  virtual int stage(orb_ctx* ctx, int ci, int c0, int nc, cudaEvent_t ready) = 0;
  virtual int preferred_wave(const orb_ctx*) const { return 0; }      // 0 = the context's staged wave size
  virtual ~WaveSource() {}
};

// plan (level geometry, tables) of the context for frames of w x h; ORB_OK or an error
int orb_internal_get_plan(orb_ctx* ctx, int w, int h);
// the wave pipeline of orb_detect_and_compute_batch; with a source, frames come from source->stage() instead of `frames`
int orb_internal_run_batch(orb_ctx* ctx, const uint8_t* frames, int frames_on_device, int n_frames, int w, int h, size_t pitch,
                           size_t frame_stride, int cap, orb_keypoint* kps, float* angles, orb_descriptor* desc, int* n_out,
                           int outputs_on_device, WaveSource* source);
