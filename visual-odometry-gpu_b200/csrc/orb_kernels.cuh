// orb_kernels.cuh -- sm_100a kernels of the ORB hot path (u8 end to end; the extractor uses no tensor cores: nothing in it
// is a dense contraction -- the matcher in orb_match_tc.cuh is, and runs on them).  Per wave of frames one small memset
// and six launches:
//   k_pyramid  : per (frame, level >= 1, 128x64 tile): bilinear resize from level 0 + 5x5 Gaussian in shared memory,
//                level pixels written once.  Replaces ORB::buildPyramid (ref src/orb.cpp:111-120 / src/orb_cpu.cpp:278-290).
//   k_fast     : per (frame, level, 128x64 tile), tile staged by one TMA load: byte-SIMD prefilter + exact compass pretest,
//                FAST-n segment test + SAD score + 3x3 NMS -> candidate positions; 5x5 box-sum image and border strip
//                tables for BRIEF.  Replaces d_Fast (ref src/cuda/Fast.cu:30-209), d_NMS (src/cuda/NMS.cu:21-128) and the
//                host cv::integral (src/cuda/Brief.cu:101-105).
//   k_edges    : per (frame, level): the values of BRIEF boxes whose centre lies in the last two columns / rows
//                (decision D7) as tables, from k_fast's strip tables (side stream, next to k_harris / k_select).
//   k_harris   : one thread per candidate: Harris response into the candidate key.  Replaces HarrisScore
//                (ref src/cuda/HarrisScore.cu:23-89 + src/Sobel.cpp + src/GaussianBlur.cpp).
//   k_select   : per (frame, level): exact top-quota selection under the total order (response desc, y asc, x asc) by
//                64-bit radix select, then raster sort (ref std::nth_element at src/orb.cpp:73-86; raster cap at
//                src/orb_cpu.cpp:110).
//   k_describe : CTA per 32 kept keypoints (warp per keypoint; lane per keypoint for the libm part); patches and box-sum
//                windows staged per keypoint by cp.async rings: intensity-centroid orientation (ref d_Orientations,
//                src/cuda/Orientations.cu:22-63 == src/orb_cpu.cpp:139-183) and rotated BRIEF with ballot-packed words
//                (ref d_Brief, src/cuda/Brief.cu:40-95 == src/orb_cpu.cpp:203-258).
// plus k_nms_scores (NMS over a caller's score map, ref NMS() include/NMS.cuh:5) and two helpers for the single-image
// stage entry points (k_harris_list, k_eval_math).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_math.cuh"
#include "orb_plan.h"

namespace orbk {

// -DORB_BOUNDS_CHECK: every shared / global index the ORB kernels compute is checked against its buffer before use
// (compute-sanitizer is closed on the B200 pool; tests/test_gpu_bounds.py runs the parity suite on such a build).  A
// failing check is counted, with the source line of the first one, in g_orb_bounds; the access still happens.
__device__ unsigned int g_orb_bounds[4];   // [0] failed checks, [1] first failing line, [2] checks executed (low 32 bits)
#ifdef ORB_BOUNDS_CHECK
__device__ __noinline__ void orb_bounds_fail(int line) {
  if (atomicAdd(&g_orb_bounds[0], 1u) == 0) g_orb_bounds[1] = (unsigned)line;
}
#define ORB_CHECK(cond) do { if (!(cond)) orb_bounds_fail(__LINE__); } while (0)
#define ORB_CHECK_COUNT() do { if (threadIdx.x == 0) atomicAdd(&g_orb_bounds[2], 1u); } while (0)
#else
#define ORB_CHECK(cond) ((void)0)
#define ORB_CHECK_COUNT() ((void)0)
#endif
// byte range [p, p + n) inside [base, base + size)
#define ORB_CHECK_RANGE(p, n, base, size) ORB_CHECK((const char*)(p) >= (const char*)(base) && (const char*)(p) + (n) <= (const char*)(base) + (size))

struct Bufs {
  const uint8_t* frames;         // level 0 of the chunk's first frame
  unsigned long long frame_stride;
  int pitch0;
  uint8_t* pyr;                  // [chunk][pyr_frame_bytes]
  uint16_t* box;                 // [chunk][box_frame_elems]
  unsigned long long* cand;      // [chunk][cand_frame_elems]
  int* cand_count;               // per frame (stride zero_stride ints): [ORB_MAX_LEVELS] candidate counters, then the
  int zero_stride;               //   BRIEF border tables of every level (zeroed together before each wave)
  uint32_t* kept_xy;             // [chunk][kept_per_frame]   (y << 16 | x), level space
  float* kept_r;                 // [chunk][kept_per_frame]
  int* kept_count;               // [chunk][ORB_MAX_LEVELS]
  const OrbTap* xtab;
  const OrbTap* ytab;
  const uint32_t* tile_a;        // per tile of k_pyramid: level | tile_x << 4 | tile_y << 18
  const uint32_t* tile_b;        // same for k_fast
  const float4* pattern;         // 256 BRIEF tests (x1,y1,x2,y2) as floats
  int* flags;                    // bit 0: candidate overflow
  int* edge2;                    // [chunk][edge2_frame_elems]: BRIEF border-box tables (k_edges)
  orb_keypoint* out_kps;         // [chunk][out_cap]
  float* out_angles;
  orb_descriptor* out_desc;
  int* out_n;                    // [chunk]
  int out_cap;
  orb_keypoint* side_xy;         // nullable side arrays, [chunk][out_cap]
  int* side_level;
  float* side_resp;
  unsigned long long frames_bytes;   // bytes addressable from `frames` (bounds-check builds)
  const CUtensorMap* tmaps;      // TMA tensor maps, [3][ORB_MAX_LEVELS]: TM_PIX, TM_BOXW, TM_PATCH per level
  int frame0;                    // index of the wave's first frame inside the level-0 tensor (the whole source batch)
};

// Tensor maps (3-D: x, y, frame; out-of-range elements read as 0).  TMA wants the box start 16-byte aligned in x.
//   TM_PIX(l)  : u8 level image, box 160 x 72  -- the k_fast tile incl. halo (x0 - 16, y0 - 4)
constexpr int TM_PIX = 0, TM_BOXW = ORB_MAX_LEVELS, TM_PATCH = 2 * ORB_MAX_LEVELS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra WAIT_%=;\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
// one box of a 3-D tensor map -> shared memory (dst 128-byte aligned), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

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
// ROW(dy, v) fills v[0..8] with the pixels of the reflect-101 extended level at row offset dy, columns -4..4 from the
// keypoint.  The 9x9 neighbourhood is walked once, row by row: per window row the vertical [1 2 1] sums of the nine
// columns give Ix = V[c+1] - V[c-1], the horizontal [1 2 1] sums of the rows above / below give Iy.
__constant__ float c_harris_w[49];   // createGaussianKernel(7), ref src/GaussianBlur.cpp:7-37 (uploaded at orb_create)

// byte `sel` (0..3 of a, 4..7 of b) of the word pair as an exact float: 2^23 + byte assembled by one byte permute
// (0x4B0000xx), minus 2^23
__device__ __forceinline__ float byte2float(uint32_t a, uint32_t b, uint32_t sel) {
  // the constant 0x4B000000 supplies bytes 1..3; selector picks the pixel byte into byte 0
  return __fsub_rn(__uint_as_float(__byte_perm(a, b, sel)), 8388608.0f);
}

// All Sobel arithmetic is done in float on small integers (|values| <= 4 * 255, products < 2^24), i.e. exactly: the
// results equal the integer Sobel sums and products the oracle converts to float, so only the 147 weighted
// accumulations round -- in the reference order.
template <typename ROW>
__device__ __forceinline__ float harris_at(ROW row, float k) {
  float A = 0.f, B = 0.f, C = 0.f;
  float pm[9], pc[9], pp[9], hm[7], hc[7], hp[7];
  row(-4, pm);
  row(-3, pc);
#pragma unroll
  for (int j = 0; j < 7; j++) {
    hm[j] = __fadd_rn(__fmaf_rn(pm[j + 1], 2.0f, pm[j]), pm[j + 2]);
    hc[j] = __fadd_rn(__fmaf_rn(pc[j + 1], 2.0f, pc[j]), pc[j + 2]);
  }
#pragma unroll
  for (int dy = -3; dy <= 3; dy++) {
    row(dy + 1, pp);
#pragma unroll
    for (int j = 0; j < 7; j++) hp[j] = __fadd_rn(__fmaf_rn(pp[j + 1], 2.0f, pp[j]), pp[j + 2]);
    float V[9];
#pragma unroll
    for (int c = 0; c < 9; c++) V[c] = __fadd_rn(__fmaf_rn(pc[c], 2.0f, pm[c]), pp[c]);
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const float ix = __fsub_rn(V[j + 2], V[j]);          // Sobel x at (dy, j-3)
      const float iy = __fsub_rn(hp[j], hm[j]);            // Sobel y at (dy, j-3)
      const float g = c_harris_w[(dy + 3) * 7 + j];
      A = orbm::fadd(A, orbm::fmul(__fmul_rn(ix, ix), g));
      B = orbm::fadd(B, orbm::fmul(__fmul_rn(ix, iy), g));
      C = orbm::fadd(C, orbm::fmul(__fmul_rn(iy, iy), g));
    }
#pragma unroll
    for (int c = 0; c < 9; c++) { pm[c] = pc[c]; pc[c] = pp[c]; }
#pragma unroll
    for (int j = 0; j < 7; j++) { hm[j] = hc[j]; hc[j] = hp[j]; }
  }
  float det = orbm::fsub(orbm::fmul(A, C), orbm::fmul(B, B));
  float tr = orbm::fadd(A, C);
  return orbm::fsub(det, orbm::fmul(orbm::fmul(k, tr), tr));
}

// FAST-n corner score of one pixel (ref src/orb_cpu.cpp:60-100) for a pixel that already passed the
// compass pretest: 0 if no arc of n contiguous ring pixels is all brighter / all darker, else the SAD score.
// c points at the centre inside a shared-memory tile of row pitch SP bytes.
template <int SP>
__device__ __forceinline__ int fast_ring_score(const uint8_t* __restrict__ c, int thr, int n) {
  // FAST ring in the reference's order (src/orb_cpu.cpp:8-13)
  constexpr int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  constexpr int ring_dy[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};
  const int Ip = c[0];
  const int hi = Ip + thr, lo = Ip - thr;
  uint32_t nb = 0, nd = 0, sad = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int v = c[ring_dy[i] * SP + ring_dx[i]];
    nb = __funnelshift_l((uint32_t)(v - hi), nb, 1);   // shifts in the sign of v - hi: 1 = not brighter
    nd = __funnelshift_l((uint32_t)(lo - v), nd, 1);   // 1 = not darker
    sad = __usad(v, Ip, sad);
  }
  // ring pixel i sits at bit 15 - i: the circle runs the other way round, which the arc test does not see
  uint32_t mb = ~nb & 0xffffu, md = ~nd & 0xffffu;
  mb |= mb << 16;
  md |= md << 16;
  uint32_t rb, rd;
  if (n == 9) {                      // runs of >= 9: doubling steps 2, 4, 8, then one more
    rb = mb & (mb >> 1); rb &= rb >> 2; rb &= rb >> 4; rb &= mb >> 8;
    rd = md & (md >> 1); rd &= rd >> 2; rd &= rd >> 4; rd &= md >> 8;
  } else {
    rb = mb; rd = md;
    for (int j = 1; j < n; j++) { rb &= mb >> j; rd &= md >> j; }
  }
  return ((rb | rd) & 0xffffu) ? (int)sad : 0;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// bytes 1 and 3 of w as two 16-bit lanes, (w >> 8) & 0x00ff00ff, in one byte permute
__device__ __forceinline__ uint32_t odd_bytes(uint32_t w) { return __byte_perm(w, 0u, 0x4341); }
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

// =============================================================================================
// Kernel A: pyramid level l >= 1 = GaussianBlur5x5( resize_linear(level 0) )  (ref src/orb_cpu.cpp:283-290).
// One CTA per 128x64 output tile.  Bilinear taps come from host tables (no floating point on the device).  Resize:
// two adjacent columns per thread from two aligned source words per row (PRMT + IDP.2A) while the level shrinks by
// <= 3, one column per thread from byte loads below that.  Blur: 16-bit lanes, two pixels per register, horizontal
// sums kept in a five-row register window; each level is written exactly once.
constexpr int A_TW = 128, A_TH = 64, A_THREADS = 256;
constexpr int A_RW = A_TW + 4, A_RH = A_TH + 4;   // resized region incl. blur halo 2
constexpr int A_RP = 136;                          // shared pitch of resized rows (bytes)
#ifndef ORB_A_UNROLL
#define ORB_A_UNROLL 3
#endif
constexpr int A_UNROLL = ORB_A_UNROLL;             // rows of the column-pair resize loop in flight together
#ifndef ORB_A_SEG
#define ORB_A_SEG 4
#endif
constexpr int A_SEG = ORB_A_SEG;                   // output rows per thread in the blur (4: all 256 threads, 8: 128)
static_assert((A_TW / 8) * (A_TH / A_SEG) <= A_THREADS && A_TH % A_SEG == 0, "blur: one thread per (column group, row segment)");

__global__ void __launch_bounds__(A_THREADS) k_pyramid(const OrbPlan P, const Bufs B) {
  // [horizontal blur sums, 16 bytes per (row, 8-pixel group) -- the tap tables live in their first 2 KB until the resize is
  // over][resized rows]: 26.7 KB, still eight CTAs per SM
  __shared__ __align__(16) uint8_t s_all[A_RH * (A_TW / 8) * 16 + A_RH * A_RP];
  uint4* s_hs = (uint4*)s_all;
  uint32_t* s_xs = (uint32_t*)s_all;             // column taps: s0 | s1 << 16
  uint32_t* s_xa = s_xs + A_RW;                  //              a0 | a1 << 16
  uint4* s_yt = (uint4*)(s_all + 2 * A_RW * 4);  // row taps: byte offsets of the two source rows inside the frame, b0, b1
  uint8_t* s_res = s_all + A_RH * (A_TW / 8) * 16;
  static_assert((2 * A_RW * 4) % 16 == 0 && 2 * A_RW * 4 + A_RH * 16 <= A_RH * (A_TW / 8) * 16, "tap tables inside the hs area");
  const int tid = threadIdx.x, f = blockIdx.y;
  const uint32_t tt = __ldg(B.tile_a + blockIdx.x);
  const int l = tt & 15;
  const OrbLevel& G = P.lv[l];
  const int x0 = ((tt >> 4) & 0x3fff) * A_TW, y0 = (tt >> 18) * A_TH;
  const int w = G.w, h = G.h;
  const bool blur = P.blur_levels != 0;
  const int halo = blur ? 2 : 0;
  const int rw = A_TW + 2 * halo, rh = A_TH + 2 * halo;
  const uint8_t* __restrict__ src = B.frames + (size_t)f * B.frame_stride;
  const uint32_t sp = (uint32_t)B.pitch0;
  ORB_CHECK_COUNT();

  for (int i = tid; i < rw + rh; i += A_THREADS) {
    if (i < rw) {
      const OrbTap tx = B.xtab[G.xtab_ofs + reflect101(x0 - halo + i, w)];
      s_xs[i] = (uint32_t)tx.s0 | ((uint32_t)tx.s1 << 16);
      s_xa[i] = (uint32_t)(uint16_t)tx.a0 | ((uint32_t)(uint16_t)tx.a1 << 16);
    } else {
      const OrbTap ty = B.ytab[G.ytab_ofs + reflect101(y0 - halo + i - rw, h)];
      s_yt[i - rw] = make_uint4(ty.s0 * sp, ty.s1 * sp, (uint32_t)ty.a0, (uint32_t)ty.a1);
    }
  }
  __syncthreads();

  // resize (cv::resize INTER_LINEAR restated: 11-bit taps, >>4, >>16, +2 >>2): a thread owns one column of the region
  // (its taps stay in registers) and walks down half of the rows; row taps are broadcast from shared memory.
  auto resize_column = [&](int col, int r0, int r1) {
    if (x0 - halo + col >= w + halo) return;
    const uint32_t xs = s_xs[col], xa = s_xa[col];
    const int a0 = xa & 0xffff, a1 = xa >> 16;
    // the second horizontal tap is the next byte (s1 == s0 + 1) except at the clamped right border, where a1 == 0
    // and the byte after the row is only read, never used (the host guarantees it is addressable)
    const uint8_t* p0 = src + (xs & 0xffff);
    asm volatile("" : "+l"(p0));      // keep the column base in a register pair: two adds per row address, not three
    r1 = min(r1, h + halo - (y0 - halo));
#pragma unroll 3
    for (int ry = r0; ry < r1; ry++) {
      ORB_CHECK(ry >= 0 && ry < A_RH && col >= 0 && col < A_RW);
      const uint4 ty = s_yt[ry];
      const uint8_t* q0 = p0 + ty.x;
      const uint8_t* q1 = p0 + ty.y;
      ORB_CHECK_RANGE(q0, 2, B.frames, B.frames_bytes);      // the second tap may be the byte after the row's last pixel
      ORB_CHECK_RANGE(q1, 2, B.frames, B.frames_bytes);
      const int h0 = __ldg(q0) * a0 + __ldg(q0 + 1) * a1;
      const int h1 = __ldg(q1) * a0 + __ldg(q1 + 1) * a1;
      const int v = ((((int)ty.z * (h0 >> 4)) >> 16) + (((int)ty.w * (h1 >> 4)) >> 16) + 2) >> 2;
      s_res[ry * A_RP + col] = (uint8_t)v;   // <= 255 without a clamp: weights sum to <= 2049 per axis, so v <= (floor(2049 * (255 * 2049 >> 4) / 65536) + 2) >> 2 = 255
    }
  };
  // Two adjacent columns at once when the level shrinks by at most 3 (source columns of neighbours are then <= 3 apart):
  // the four source bytes of a row (two taps x two columns) lie in two aligned words, so a row pair costs 4 word loads,
  // 2 byte permutes and 4 two-way dot products (IDP.2A: a0 * S0 + a1 * S1) instead of 8 byte loads and 8 multiply-adds.
  auto resize_pair = [&](int col, int r0, int r1) {    // col even
    if (x0 - halo + col >= w + halo) return;
    const uint32_t xa0 = s_xa[col], xa1 = s_xa[col + 1];
    const int sa = s_xs[col] & 0xffff, sb = s_xs[col + 1] & 0xffff;
    const int base = min(min(sa, sb) & ~3, (int)sp - 8);          // both words inside the row (pitch is a multiple of 16)
    const int oa = sa - base, ob = sb - base;                      // <= 6: the second taps sit at <= 7
    const uint32_t sel = (uint32_t)(oa | ((oa + 1) << 4) | (ob << 8) | ((ob + 1) << 12));
    const uint8_t* p0 = src + base;
    asm volatile("" : "+l"(p0));
    r1 = min(r1, h + halo - (y0 - halo));
#pragma unroll (A_UNROLL)
    for (int ry = r0; ry < r1; ry++) {
      ORB_CHECK(ry >= 0 && ry < A_RH && col >= 0 && col + 1 < A_RP);
      const uint4 ty = s_yt[ry];
      const uint32_t* q0 = (const uint32_t*)(p0 + ty.x);
      const uint32_t* q1 = (const uint32_t*)(p0 + ty.y);
      ORB_CHECK_RANGE(q0, 8, B.frames, B.frames_bytes);
      ORB_CHECK_RANGE(q1, 8, B.frames, B.frames_bytes);
      ORB_CHECK(((uintptr_t)q0 & 3) == 0 && (uint32_t)(base + 8) <= sp);          // both words inside the row
      const uint32_t t0 = prmt(__ldg(q0), __ldg(q0 + 1), sel), t1 = prmt(__ldg(q1), __ldg(q1 + 1), sel);
      const int h0a = (int)__dp2a_lo(xa0, t0, 0u), h0b = (int)__dp2a_hi(xa1, t0, 0u);
      const int h1a = (int)__dp2a_lo(xa0, t1, 0u), h1b = (int)__dp2a_hi(xa1, t1, 0u);
      // (b0 * (h0 >> 4) >> 16) + (b1 * (h1 >> 4) >> 16) + 2 >> 2 for both columns in 16-bit lanes: the products are < 2^27, so
      // one byte permute takes ">> 16" of two of them at once, the lane sums stay < 2^13, and the two bits the final shift
      // drags from the upper lane land above the result byte (va, vb <= 255, see resize_column)
      const uint32_t pa0 = (uint32_t)((int)ty.z * (h0a >> 4)), pa1 = (uint32_t)((int)ty.w * (h1a >> 4));
      const uint32_t pb0 = (uint32_t)((int)ty.z * (h0b >> 4)), pb1 = (uint32_t)((int)ty.w * (h1b >> 4));
      const uint32_t v2 = (prmt(pa0, pb0, 0x7632) + prmt(pa1, pb1, 0x7632) + 0x00020002u) >> 2;
      *(uint16_t*)(s_res + ry * A_RP + col) = (uint16_t)prmt(v2, 0u, 0x4420);
    }
  };
  if (P.W <= 3 * w) {
    const int quarter = tid >> 6, rows_q = (rh + 3) >> 2;
    resize_pair(2 * (tid & 63), quarter * rows_q, min(rh, (quarter + 1) * rows_q));
    if (rw > A_TW)                                     // the 4 extra halo columns: 2 pairs x rh rows, one item per thread
      for (int extra = tid; extra < 2 * rh; extra += A_THREADS)
        resize_pair(A_TW + 2 * (extra & 1), extra >> 1, (extra >> 1) + 1);
  } else {
    const int half = tid >> 7, rows_half = (rh + 1) >> 1;
    resize_column(tid & 127, half * rows_half, min(rh, (half + 1) * rows_half));
    if (rw > A_TW)                                     // the 4 extra halo columns: 4 x rh pixels, one per thread
      for (int extra = tid; extra < 4 * rh; extra += A_THREADS)   // column 128 + extra % 4, row extra / 4
        resize_column(A_TW + (extra & 3), extra >> 2, (extra >> 2) + 1);
  }
  __syncthreads();

  uint8_t* dst = B.pyr + (size_t)f * P.pyr_frame_bytes + G.lvl_ofs;
  if (!blur) {   // ref src/orb.cpp:119: resize only
    for (int it = tid; it < A_TH * (A_TW / 8); it += A_THREADS) {
      const int oy = it / (A_TW / 8), g = it - oy * (A_TW / 8);
      if (y0 + oy >= h || x0 + 8 * g >= G.pitch) continue;
      *(uint2*)(dst + (size_t)(y0 + oy) * G.pitch + x0 + 8 * g) = *(const uint2*)(s_res + oy * A_RP + 8 * g);
    }
    return;
  }

  // [1 4 6 4 1] x [1 4 6 4 1] in two steps.  (1) The horizontal sums of every resized row are computed once (a thread takes
  // (row, 8-pixel group) items; 16-bit lanes holding pixels two apart: (o0,o2) (o1,o3) (o4,o6) (o5,o7)) and stored to shared
  // memory.  (2) A thread owns a column group and A_SEG output rows and walks down the A_SEG + 4 rows of sums with a
  // five-row register window.  One rounding: (sum + 128) >> 8  (cv::GaussianBlur 5x5 sigma 0 on u8).
  {
    const uint32_t M = 0x00ff00ffu;
    for (int it = tid; it < A_RH * (A_TW / 8); it += A_THREADS) {
      const int g = it & (A_TW / 8 - 1), row = it / (A_TW / 8);
      const uint8_t* r = s_res + row * A_RP + 8 * g;
      const uint2 w01 = *(const uint2*)r;
      const uint32_t w2 = *(const uint32_t*)(r + 8);
      const uint32_t v1 = prmt(w01.x, w01.y, 0x5432), v2 = prmt(w01.y, w2, 0x5432);
      const uint32_t q0 = w01.x & M, q1 = odd_bytes(w01.x), q2 = v1 & M, q3 = odd_bytes(v1), q4 = w01.y & M,
                     q5 = odd_bytes(w01.y), q6 = v2 & M, q7 = odd_bytes(v2), q8 = w2 & M, q9 = odd_bytes(w2);
      uint4 o;
      o.x = q0 + q4 + 4 * (q1 + q3) + 6 * q2;
      o.y = q1 + q5 + 4 * (q2 + q4) + 6 * q3;
      o.z = q4 + q8 + 4 * (q5 + q7) + 6 * q6;
      o.w = q5 + q9 + 4 * (q6 + q8) + 6 * q7;
      s_hs[it] = o;
    }
  }
  __syncthreads();
  if (tid < (A_TW / 8) * (A_TH / A_SEG)) {
    const int g = tid & (A_TW / 8 - 1), oy0 = (tid / (A_TW / 8)) * A_SEG;
    if (y0 + oy0 < h && x0 + 8 * g < G.pitch) {
      const uint32_t R = 0x00800080u;
      uint8_t* d = dst + (size_t)(y0 + oy0) * G.pitch + x0 + 8 * g;
      ORB_CHECK(oy0 + A_SEG + 4 <= A_RH);
      ORB_CHECK_RANGE(d, 8, dst, (size_t)G.h * G.pitch);                // (later rows of the segment are guarded by y < h)
      uint4 hs[5];
#pragma unroll
      for (int k = 0; k < A_SEG + 4; k++) {
        const uint4 o = s_hs[(oy0 + k) * (A_TW / 8) + g];
        hs[k % 5] = o;
        if (k >= 4 && y0 + oy0 + k - 4 < h) {      // output row oy0 + k - 4 from resized rows k-4 .. k
          const uint4 &m2 = hs[(k + 1) % 5], &m1 = hs[(k + 2) % 5], &c0 = hs[(k + 3) % 5], &p1 = hs[(k + 4) % 5];
          const uint32_t a = m2.x + o.x + 4 * (m1.x + p1.x) + 6 * c0.x + R;
          const uint32_t b = m2.y + o.y + 4 * (m1.y + p1.y) + 6 * c0.y + R;
          const uint32_t cc = m2.z + o.z + 4 * (m1.z + p1.z) + 6 * c0.z + R;
          const uint32_t dd = m2.w + o.w + 4 * (m1.w + p1.w) + 6 * c0.w + R;
          uint2 out;
          out.x = prmt(a, b, 0x7351);    // (a>>8 lane0, b>>8 lane0, a>>8 lane1, b>>8 lane1) = pixels 0,1,2,3
          out.y = prmt(cc, dd, 0x7351);
          *(uint2*)(d + (size_t)(k - 4) * G.pitch) = out;
        }
      }
    }
  }
}

// =============================================================================================
// Kernel B: FAST-n + SAD score + 3x3 NMS + 5x5 box sums on one 128x64 tile of one level (tile + 4-pixel halo staged by one
// TMA load, out-of-level pixels zero-filled).
//   pretest  : the >=3-of-4 compass test (ref src/orb_cpu.cpp:39-58): a byte-SIMD necessary condition (VABSDIFF4, 4 pixels /
//              instruction) drops ~2/3 of the 8-pixel row items; the survivors, compacted by ballot ranks, take the exact
//              test on packed half2 (2 pixels / instruction): second smallest / second largest of the four compass pixels
//              against Ip +- thr, or saturated counts;
//   ring     : passers are compacted into a shared list and finished one pixel per thread (arc + SAD);
//   NMS      : score == max of its 3x3 window, ties keep both (:126); survivors are appended to the level's
//              candidate list (k_harris adds their response);
//   box sums : 5x5 sums of the level (u16) for BRIEF, 16-bit lanes, written once; strip tables for the border rule.
constexpr int B_TW = 128, B_TH = 64, B_THREADS = 256;
constexpr int B_SP = 160;              // pixel tile pitch (bytes); pixel x sits at column x - x0 + 16
constexpr int B_PH = B_TH + 8;         // rows y0-4 .. y0+B_TH+3
constexpr int B_SCP = 136;             // score pitch (u16); pixel x sits at column x - x0 + 4
constexpr int B_SH = B_TH + 2;         // rows y0-1 .. y0+B_TH
#ifndef ORB_B_LIST
#define ORB_B_LIST 2704
#endif
#ifndef ORB_B_MINB
#define ORB_B_MINB 6
#endif
#ifndef ORB_PRETEST_SPLIT
#define ORB_PRETEST_SPLIT 2
#endif
#ifndef ORB_B_PREFILTER
#define ORB_B_PREFILTER 1
#endif
#ifndef ORB_B_DIRECT_EMIT
#define ORB_B_DIRECT_EMIT 1
#endif
constexpr int B_LIST = ORB_B_LIST;     // pretest passers kept in the list; denser tiles take the dense fallback
constexpr int B_SURV = 1024;
constexpr int B_PIX_BYTES = B_PH * B_SP, B_SCORE_BYTES = B_SH * B_SCP * 2;
constexpr int B_LIST_BYTES = B_LIST * 2;
constexpr int B_SMEM = B_PIX_BYTES + B_SCORE_BYTES + B_LIST_BYTES + B_SURV * 2 + 16 + B_TW * 4 + 64 + 56 * 4 + 16;   // 37.7 KB and 40 registers -> 6 CTAs / SM

__global__ void __launch_bounds__(B_THREADS, ORB_B_MINB) k_fast(const OrbPlan P, const Bufs B) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_pix = smem;                  // TMA destination: 128-byte aligned
  uint16_t* s_score = (uint16_t*)(smem + B_PIX_BYTES);
  uint16_t* s_list = (uint16_t*)(smem + B_PIX_BYTES + B_SCORE_BYTES);
  uint16_t* s_surv = (uint16_t*)(smem + B_PIX_BYTES + B_SCORE_BYTES + B_LIST_BYTES);
  int* s_ey = (int*)(s_surv + B_SURV);    // [B_TW] column-strip sums of this tile (phase 5/6)
  int* s_ctr = s_ey + B_TW;               // [0] pretest list, [1] survivor list, [2] global base, [3] item list
  uint16_t* s_tab = (uint16_t*)(s_ctr + 4);     // [32] passer bit -> tile offset (phase 2, list variant)
  uint32_t* s_vm = (uint32_t*)(s_tab + 32);     // [18][3] per column group: pair-lane validity mask, byte validity masks (phase 2)
  uint64_t* s_bar = (uint64_t*)(s_vm + 54 + 2); // mbarrier of the tile load
  uint16_t* s_items = s_surv;             // [<= 18 * 66] row items that survive the prefilter: lives in s_surv + the head of
                                          // s_ey until phase 2 is over (s_ey is cleared after that)

  const int tid = threadIdx.x, lane = tid & 31, f = blockIdx.y;
  const uint32_t tt = __ldg(B.tile_b + blockIdx.x);
  const int l = tt & 15;
  const OrbLevel& G = P.lv[l];
  const int x0 = ((tt >> 4) & 0x3fff) * B_TW, y0 = (tt >> 18) * B_TH;
  const int w = G.w, h = G.h;
  ORB_CHECK_COUNT();

  // ---- phase 0/1: stage the tile with one TMA load (box 160 x 72 at (x0 - 16, y0 - 4); everything outside the level
  // reads as 0 -- no valid output depends on it: FAST centres stay >= 3 pixels inside, the 5x5 sums that BRIEF may read
  // have their centre >= 2 pixels inside) and clear the score map while the copy is in flight ---------------------------
  if (tid == 0) {
    mbar_init(s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(s_bar, B_PIX_BYTES);
    tma_load_3d(s_pix, B.tmaps + TM_PIX + l, s_bar, x0 - 16, y0 - 4, l == 0 ? f + B.frame0 : f);
  }
  for (int i = tid; i < B_SCORE_BYTES / 16; i += B_THREADS) ((uint4*)s_score)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 4) s_ctr[tid] = 0;
#if !ORB_B_PREFILTER
  if (tid < B_TW) s_ey[tid] = 0;
#endif
  // passer bit b of a thread's mask word: row item b >> 3 (rows 14 apart), pixel 0,2,4,6,1,3,5,7 for b & 7 = 0..7
  if (tid < 32) s_tab[tid] = (uint16_t)((tid >> 3) * (14 * B_SP) + ((tid & 3) << 1) + ((tid >> 2) & 1));
#if ORB_B_PREFILTER
  {
    constexpr int NG = 18;
    if (tid < NG) {   // per column group: which of its 8 pixels are valid centres
      const int xs = x0 - 8 + 8 * tid;
      // valid centres: 3 <= x < w-3 (ref src/orb_cpu.cpp:35), inside tile + 1 halo column on each side
      const int lo = max(max(0, 3 - xs), x0 - 1 - xs), hi = min(min(8, w - 3 - xs), x0 + B_TW + 1 - xs);
      uint32_t vmask = 0, vb0 = 0, vb1 = 0;
      if (lo < hi) {
        const uint32_t v8m = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
  #pragma unroll
        for (int q = 0; q < 4; q++) {                        // pixel 2q -> bit 10+q, pixel 2q+1 -> bit 26+q (see below)
          vmask |= (((v8m >> (2 * q)) & 1u) << (10 + q)) | (((v8m >> (2 * q + 1)) & 1u) << (26 + q));
          vb0 |= ((v8m >> q) & 1u) << (8 * q + 7);
          vb1 |= ((v8m >> (4 + q)) & 1u) << (8 * q + 7);
        }
      }
      s_vm[3 * tid] = vmask; s_vm[3 * tid + 1] = vb0; s_vm[3 * tid + 2] = vb1;
    }
  }
#endif
  __syncthreads();          // the mbarrier is initialised for everybody
  mbar_wait(s_bar, 0);

#if ORB_B_PREFILTER
  // ---- phase 2: compass pretest (ref src/orb_cpu.cpp:39-58), 8 pixels per row item, in two steps -----------------
  // (a) prefilter, byte SIMD (4 pixels / instruction): a pixel can only have >= 3 of its 4 compass pixels brighter (or
  //     darker) by thr if at least one of top / bottom AND at least one of left / right differs from it by >= thr.  With
  //     d = |a - b| per byte (VABSDIFF4), (dT | dB) >= max(dT, dB), so "(dT | dB) >= thr and (dL | dR) >= thr" is a
  //     necessary condition; the byte compare is one add whose carry lands in bit 7.  About two thirds of the row items of
  //     a KITTI-like frame have no such pixel and are dropped here, at ~40 instructions per item.
  // (b) the exact test on packed half2 (2 pixels / instruction) only for the surviving items, compacted into a list so
  //     that every lane of a warp has an item.
  const int thr = P.fast_threshold, fn = P.fast_n;
  constexpr int NG = 18, NRT = 14, NK = (B_SH + NRT - 1) / NRT;
  {
    const int g = tid % NG, rt = tid / NG;
    const int pc = 8 + 8 * g;
    const uint32_t vb0 = rt < NRT ? s_vm[3 * g + 1] : 0u, vb1 = rt < NRT ? s_vm[3 * g + 2] : 0u;
    const uint32_t K7 = (uint32_t)(0x80 - min(thr, 128)) * 0x01010101u;
    const bool wide = thr > 128;               // then bit 7 of d itself (d >= 128) is the (weaker, still necessary) test
    uint32_t sb = 0;                           // bit k: row item k survives
#pragma unroll
    for (int k = 0; k < NK; k++) {
      const int sy = rt + NRT * k, y = y0 - 1 + sy;
      if ((vb0 | vb1) && sy < B_SH && y >= 3 && y < h - 3) {      // 3 <= y < h-3 (ref src/orb_cpu.cpp:34)
        const uint8_t* rc = s_pix + (sy + 3) * B_SP + pc;
        ORB_CHECK(rc - 3 * B_SP - 4 >= s_pix && rc + 3 * B_SP + 12 <= s_pix + B_PIX_BYTES);
        const uint32_t m = *(const uint32_t*)(rc - 4), p = *(const uint32_t*)(rc + 8);
        const uint2 cc = *(const uint2*)rc, tt = *(const uint2*)(rc - 3 * B_SP), bb = *(const uint2*)(rc + 3 * B_SP);
        const uint32_t v0 = __vabsdiffu4(tt.x, cc.x) | __vabsdiffu4(bb.x, cc.x), v1 = __vabsdiffu4(tt.y, cc.y) | __vabsdiffu4(bb.y, cc.y);
        const uint32_t h0 = __vabsdiffu4(prmt(m, cc.x, 0x4321), cc.x) | __vabsdiffu4(prmt(cc.x, cc.y, 0x6543), cc.x);
        const uint32_t h1 = __vabsdiffu4(prmt(cc.x, cc.y, 0x4321), cc.y) | __vabsdiffu4(prmt(cc.y, p, 0x6543), cc.y);
        uint32_t f0, f1;
        if (!wide) {
          f0 = (((v0 & 0x7f7f7f7fu) + K7) | v0) & (((h0 & 0x7f7f7f7fu) + K7) | h0);
          f1 = (((v1 & 0x7f7f7f7fu) + K7) | v1) & (((h1 & 0x7f7f7f7fu) + K7) | h1);
        } else {
          f0 = v0 & h0; f1 = v1 & h1;
        }
        if ((f0 & vb0) | (f1 & vb1)) sb |= 1u << k;
      }
    }
    // compact the surviving items of the warp into the item list (ballot ranks; one shared atomic per warp)
    unsigned bal[NK];
    int tot = 0;
#pragma unroll
    for (int k = 0; k < NK; k++) { bal[k] = __ballot_sync(0xffffffffu, (sb >> k) & 1u); tot += __popc(bal[k]); }
    if (tot) {
      int o = 0;
      if (lane == 0) o = atomicAdd(&s_ctr[3], tot);
      o = __shfl_sync(0xffffffffu, o, 0);
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int k = 0; k < NK; k++) {
        if ((sb >> k) & 1u) s_items[o + __popc(bal[k] & lt)] = (uint16_t)(((rt + NRT * k) << 5) | g);
        o += __popc(bal[k]);
      }
    }
  }
  __syncthreads();
  {
    const uint32_t K = 0x64646464u;   // half(1024 + p) = 0x6400 | p
    // Two equivalent formulations share the work between the ALU pipe (HMNMX2 / HSET2) and the half-precision adder:
    // (a) ">= 3 of 4 have v >= Ip + thr" <=> the second smallest of the four >= Ip + thr (min/max network);
    // (b) the count of such pixels is a sum of saturated differences, sat(v - Ip - thr + 1) in {0, 1} (small integers,
    //     exact in half precision), and ">= 3" is sat(count - 2): HADD2(.SAT) only.
    // The first ORB_PRETEST_SPLIT pixel pairs of an item use (a), the rest (b).
    // The darker side uses max(thr, 1) so that a pixel is never counted on both sides (ref src/orb_cpu.cpp:44-57).
    const __half2 one_m_thr = __float2half2_rn((float)(1 - thr)), one_m_dthr = __float2half2_rn((float)(1 - max(thr, 1)));
    const __half2 minus2 = __float2half2_rn(-2.0f);
    const __half2 thr2 = __float2half2_rn((float)thr), dthr2 = __float2half2_rn((float)max(thr, 1));
    const int n_items = s_ctr[3];
    for (int it = tid; it < n_items; it += B_THREADS) {
      const int code = s_items[it], sy = code >> 5, g = code & 31, pc = 8 + 8 * g;
      const uint32_t vmask = s_vm[3 * g];
      const uint8_t* rc = s_pix + (sy + 3) * B_SP + pc;
      ORB_CHECK(sy < B_SH && g < NG && rc - 3 * B_SP - 4 >= s_pix && rc + 3 * B_SP + 12 <= s_pix + B_PIX_BYTES);
      const uint32_t m = *(const uint32_t*)(rc - 4), p = *(const uint32_t*)(rc + 8);
      const uint2 cc = *(const uint2*)rc, tt = *(const uint2*)(rc - 3 * B_SP), bb = *(const uint2*)(rc + 3 * B_SP);
      const uint32_t C[4] = {prmt(cc.x, K, 0x4140), prmt(cc.x, K, 0x4342), prmt(cc.y, K, 0x4140), prmt(cc.y, K, 0x4342)};
      const uint32_t T[4] = {prmt(tt.x, K, 0x4140), prmt(tt.x, K, 0x4342), prmt(tt.y, K, 0x4140), prmt(tt.y, K, 0x4342)};
      const uint32_t Bm[4] = {prmt(bb.x, K, 0x4140), prmt(bb.x, K, 0x4342), prmt(bb.y, K, 0x4140), prmt(bb.y, K, 0x4342)};
      // pixel pairs three to the left / right of each centre pair
      const uint32_t p34 = prmt(prmt(cc.x, cc.y, 0x0043), K, 0x4140);
      const uint32_t Lf[4] = {prmt(m, K, 0x4241), prmt(prmt(m, cc.x, 0x0043), K, 0x4140), prmt(cc.x, K, 0x4241), p34};
      const uint32_t Rt[4] = {p34, prmt(cc.y, K, 0x4241), prmt(prmt(cc.y, p, 0x0043), K, 0x4140), prmt(p, K, 0x4241)};
      uint32_t flags = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const __half2 v0 = as_h2(T[q]), v4 = as_h2(Rt[q]), v8 = as_h2(Bm[q]), v12 = as_h2(Lf[q]), c2 = as_h2(C[q]);
        if (q < ORB_PRETEST_SPLIT) {   // min/max network (ALU pipe): 2nd smallest / 2nd largest of the four against Ip +- thr
          const __half2 a = __hmin2(v0, v4), b = __hmax2(v0, v4), c = __hmin2(v8, v12), d = __hmax2(v8, v12);
          const __half2 m1 = __hmax2(a, c), m2 = __hmin2(b, d);
          const __half2 s2 = __hmin2(m1, m2), l2 = __hmax2(m1, m2);
          const uint32_t fb = __hge2_mask(__hsub2(s2, c2), thr2);
          const uint32_t fd = __hge2_mask(__hsub2(c2, l2), dthr2);
          flags |= (fb | fd) & (0x04000400u << q);
        } else {                       // saturated counts (half-precision adds)
          const __half2 nhi = __hsub2(one_m_thr, c2), plo = __hadd2(c2, one_m_dthr);
          const __half2 nb = __hadd2(__hadd2(__hadd2_sat(v0, nhi), __hadd2_sat(v4, nhi)),
                                     __hadd2(__hadd2_sat(v8, nhi), __hadd2_sat(v12, nhi)));
          const __half2 nd = __hadd2(__hadd2(__hadd2_sat(plo, __hneg2(v0)), __hadd2_sat(plo, __hneg2(v4))),
                                     __hadd2(__hadd2_sat(plo, __hneg2(v8)), __hadd2_sat(plo, __hneg2(v12))));
          const __half2 pass = __hadd2(__hadd2_sat(nb, minus2), __hadd2_sat(nd, minus2));   // 0 or 1.0 (0x3c00) per pixel
          flags |= (*reinterpret_cast<const uint32_t*>(&pass)) & (0x04000400u << q);
        }
      }
      flags &= vmask;
      uint32_t m8 = ((flags >> 10) & 0xfu) | ((flags >> 22) & 0xf0u);   // bits 0-3: pixels 0,2,4,6; bits 4-7: pixels 1,3,5,7
      if (m8) {
        int off = atomicAdd(&s_ctr[0], __popc(m8));
        const int e0 = sy * B_SP + pc;
        do {
          const int b = __ffs(m8) - 1;
          m8 &= m8 - 1;
          if (off < B_LIST) s_list[off] = (uint16_t)(e0 + ((b & 3) << 1) + (b >> 2));
          off++;
        } while (m8);
      }
    }
  }
  __syncthreads();
  if (tid < B_TW) s_ey[tid] = 0;        // (the item list used this space; the strip sums start at phase 5)

#else
  // ---- phase 2: compass pretest, 8 pixels per item -------------------------------------------
  // Thread (g, rt) owns column group g (8 pixels from x0 - 8 + 8g) on rows rt, rt+14, ...; the column validity mask
  // is loop invariant.  Passers of all its rows are appended with one warp scan + one shared atomic per warp.
  const int thr = P.fast_threshold, fn = P.fast_n;
  {
    constexpr int NG = 18, NRT = 14, NK = (B_SH + NRT - 1) / NRT;
    static_assert(NRT == 14, "s_tab is filled for rows 14 apart");
    const uint32_t K = 0x64646464u;   // half(1024 + p) = 0x6400 | p
    // Two equivalent formulations share the work between the ALU pipe (HMNMX2 / HSET2) and the half-precision adder:
    // (a) ">= 3 of 4 have v >= Ip + thr" <=> the second smallest of the four >= Ip + thr (min/max network);
    // (b) the count of such pixels is a sum of saturated differences, sat(v - Ip - thr + 1) in {0, 1} (small integers,
    //     exact in half precision), and ">= 3" is sat(count - 2): HADD2(.SAT) only.
    // The first ORB_PRETEST_SPLIT pixel pairs of an item use (a), the rest (b) (measured: 2 is best, by ~1.5 %).
    // The darker side uses max(thr, 1) so that a pixel is never counted on both sides (ref src/orb_cpu.cpp:44-57).
    const __half2 one_m_thr = __float2half2_rn((float)(1 - thr)), one_m_dthr = __float2half2_rn((float)(1 - max(thr, 1)));
    const __half2 minus2 = __float2half2_rn(-2.0f);
    const __half2 thr2 = __float2half2_rn((float)thr), dthr2 = __float2half2_rn((float)max(thr, 1));
    const int g = tid % NG, rt = tid / NG;
    const int pc = 8 + 8 * g, xs = x0 - 8 + 8 * g;
    // valid centres: 3 <= x < w-3 (ref src/orb_cpu.cpp:35), inside tile + 1 halo column on each side
    const int lo = max(max(0, 3 - xs), x0 - 1 - xs), hi = min(min(8, w - 3 - xs), x0 + B_TW + 1 - xs);
    uint32_t vmask = 0;
    if (lo < hi && rt < NRT) {
      const uint32_t v8m = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
#pragma unroll
      for (int q = 0; q < 4; q++)                          // pixel 2q -> bit 10+q, pixel 2q+1 -> bit 26+q (see below)
        vmask |= (((v8m >> (2 * q)) & 1u) << (10 + q)) | (((v8m >> (2 * q + 1)) & 1u) << (26 + q));
    }
    static_assert(NK <= 8, "passer masks of a thread: one byte per row item, two words");
    uint32_t mlo = 0, mhi = 0;        // byte k: passers of row item k (k < 4 in mlo, the rest in mhi), bit order below
#pragma unroll
    for (int k = 0; k < NK; k++) {
      const int sy = rt + NRT * k, y = y0 - 1 + sy;
      uint32_t flags = 0;
      if (vmask && sy < B_SH && y >= 3 && y < h - 3) {      // 3 <= y < h-3 (ref src/orb_cpu.cpp:34)
        const uint8_t* rc = s_pix + (sy + 3) * B_SP + pc;
        ORB_CHECK(rc - 3 * B_SP - 4 >= s_pix && rc + 3 * B_SP + 12 <= s_pix + B_PIX_BYTES);
        const uint32_t m = *(const uint32_t*)(rc - 4), p = *(const uint32_t*)(rc + 8);
        const uint2 cc = *(const uint2*)rc, tt = *(const uint2*)(rc - 3 * B_SP), bb = *(const uint2*)(rc + 3 * B_SP);
        const uint32_t C[4] = {prmt(cc.x, K, 0x4140), prmt(cc.x, K, 0x4342), prmt(cc.y, K, 0x4140), prmt(cc.y, K, 0x4342)};
        const uint32_t T[4] = {prmt(tt.x, K, 0x4140), prmt(tt.x, K, 0x4342), prmt(tt.y, K, 0x4140), prmt(tt.y, K, 0x4342)};
        const uint32_t Bm[4] = {prmt(bb.x, K, 0x4140), prmt(bb.x, K, 0x4342), prmt(bb.y, K, 0x4140), prmt(bb.y, K, 0x4342)};
        // pixel pairs three to the left / right of each centre pair
        const uint32_t p34 = prmt(prmt(cc.x, cc.y, 0x0043), K, 0x4140);
        const uint32_t Lf[4] = {prmt(m, K, 0x4241), prmt(prmt(m, cc.x, 0x0043), K, 0x4140), prmt(cc.x, K, 0x4241), p34};
        const uint32_t Rt[4] = {p34, prmt(cc.y, K, 0x4241), prmt(prmt(cc.y, p, 0x0043), K, 0x4140), prmt(p, K, 0x4241)};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const __half2 v0 = as_h2(T[q]), v4 = as_h2(Rt[q]), v8 = as_h2(Bm[q]), v12 = as_h2(Lf[q]), c2 = as_h2(C[q]);
          if (q < ORB_PRETEST_SPLIT) {   // min/max network (ALU pipe): 2nd smallest / 2nd largest of the four against Ip +- thr
            const __half2 a = __hmin2(v0, v4), b = __hmax2(v0, v4), c = __hmin2(v8, v12), d = __hmax2(v8, v12);
            const __half2 m1 = __hmax2(a, c), m2 = __hmin2(b, d);
            const __half2 s2 = __hmin2(m1, m2), l2 = __hmax2(m1, m2);
            const uint32_t fb = __hge2_mask(__hsub2(s2, c2), thr2);
            const uint32_t fd = __hge2_mask(__hsub2(c2, l2), dthr2);
            flags |= (fb | fd) & (0x04000400u << q);
          } else {                       // saturated counts (half-precision adds)
            const __half2 nhi = __hsub2(one_m_thr, c2), plo = __hadd2(c2, one_m_dthr);
            const __half2 nb = __hadd2(__hadd2(__hadd2_sat(v0, nhi), __hadd2_sat(v4, nhi)),
                                       __hadd2(__hadd2_sat(v8, nhi), __hadd2_sat(v12, nhi)));
            const __half2 nd = __hadd2(__hadd2(__hadd2_sat(plo, __hneg2(v0)), __hadd2_sat(plo, __hneg2(v4))),
                                       __hadd2(__hadd2_sat(plo, __hneg2(v8)), __hadd2_sat(plo, __hneg2(v12))));
            const __half2 pass = __hadd2(__hadd2_sat(nb, minus2), __hadd2_sat(nd, minus2));   // 0 or 1.0 (0x3c00) per pixel
            flags |= (*reinterpret_cast<const uint32_t*>(&pass)) & (0x04000400u << q);
          }
        }
        flags &= vmask;
      }
      const uint32_t m8 = ((flags >> 10) & 0xfu) | ((flags >> 22) & 0xf0u);   // bits 0-3: pixels 0,2,4,6; bits 4-7: pixels 1,3,5,7
      if (k < 4) mlo |= m8 << (8 * k);
      else mhi |= m8 << (8 * (k - 4));
    }
    const int cnt = __popc(mlo) + __popc(mhi);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total) {
      int base = 0;
      if (lane == 31) base = atomicAdd(&s_ctr[0], total);
      base = __shfl_sync(0xffffffffu, base, 31);
      int off = base + incl - cnt;
      // one loop over all passers of the thread (row items flattened: the warp runs max-over-lanes of the thread totals,
      // not the sum over row items of the per-item maxima)
      const int e0 = rt * B_SP + pc;
      while (mlo) {
        const int b = __ffs(mlo) - 1;
        mlo &= mlo - 1;
        if (off < B_LIST) s_list[off] = (uint16_t)(e0 + s_tab[b]);
        off++;
      }
      while (mhi) {
        const int b = __ffs(mhi) - 1;
        mhi &= mhi - 1;
        if (off < B_LIST) s_list[off] = (uint16_t)(e0 + 4 * NRT * B_SP + s_tab[b]);
        off++;
      }
    }
  }
  __syncthreads();

#endif
  // ---- phase 3: ring test + SAD score for the passers, one pixel per thread --------------------
  // (a tile with more than B_LIST passers -- synthetic worst cases only -- is rescanned densely instead)
  const int n1 = s_ctr[0];
  const bool dense = n1 > B_LIST;
  if (!dense) {
    for (int j = tid; j < n1; j += B_THREADS) {
      const int idx = s_list[j];
      const int sy = idx / B_SP, pcx = idx - sy * B_SP;
      ORB_CHECK(j < B_LIST && sy >= 0 && sy < B_SH && pcx >= 15 && pcx <= 16 + B_TW);     // ring reach 3: rows sy .. sy+6, columns pcx-3 .. pcx+3
      ORB_CHECK(sy * B_SCP + pcx - 12 >= 0 && sy * B_SCP + pcx - 12 < B_SH * B_SCP);
      const int sc = fast_ring_score<B_SP>(s_pix + (sy + 3) * B_SP + pcx, thr, fn);
      if (sc) s_score[sy * B_SCP + pcx - 12] = (uint16_t)sc;
    }
  } else {
    for (int i = tid; i < B_SH * (B_TW + 2); i += B_THREADS) {
      const int sy = i / (B_TW + 2), pcx = 15 + (i - sy * (B_TW + 2));
      const int x = x0 + pcx - 16, y = y0 - 1 + sy;
      if (x < 3 || x >= w - 3 || y < 3 || y >= h - 3) continue;
      const uint8_t* c = s_pix + (sy + 3) * B_SP + pcx;
      const int Ip = c[0], hi = Ip + thr, lo = Ip - thr;
      const int v0 = c[-3 * B_SP], v4 = c[3], v8 = c[3 * B_SP], v12 = c[-3];
      const int br = (v0 >= hi) + (v4 >= hi) + (v8 >= hi) + (v12 >= hi);
      const int dk = (v0 < hi && v0 <= lo) + (v4 < hi && v4 <= lo) + (v8 < hi && v8 <= lo) + (v12 < hi && v12 <= lo);
      if (max(br, dk) < 3) continue;                       // ref src/orb_cpu.cpp:39-58
      const int sc = fast_ring_score<B_SP>(c, thr, fn);
      if (sc) s_score[sy * B_SCP + pcx - 12] = (uint16_t)sc;
    }
  }
  __syncthreads();

  // ---- phase 4: NMS over the corners of the tile interior ---------------------------------------
  unsigned long long* cand = B.cand + (size_t)f * P.cand_frame_elems + G.cand_ofs;
  int* gcount = B.cand_count + (size_t)f * B.zero_stride + l;
  const int nmsr = P.nms_radius;
  auto emit = [&](int idx, int slot) {   // key = raster position; k_harris adds the response in the high word
    const int sy = idx / B_SP, pcx = idx - sy * B_SP;
    const int lx = x0 + pcx - 16, ly = y0 - 1 + sy;
    ORB_CHECK(lx >= 3 && lx < w - 3 && ly >= 3 && ly < h - 3 && slot >= 0);
    if (slot < G.cand_cap) cand[slot] = (unsigned long long)(unsigned)((ly << 16) | lx);
  };
  auto nms_one = [&](int idx) {
    const int sy = idx / B_SP, pcx = idx - sy * B_SP;
    if (sy < 1 || sy > B_TH || pcx < 16 || pcx >= 16 + B_TW) return;   // halo positions belong to neighbours
    const uint16_t* s = s_score + sy * B_SCP + pcx - 12;
    ORB_CHECK(s - B_SCP - 1 >= s_score && s + B_SCP + 1 < s_score + B_SH * B_SCP);
    const int v = s[0];
    if (v == 0) return;
    if (nmsr) {   // ties keep both (ref src/orb_cpu.cpp:126)
      int mx = max(max(max(s[-B_SCP - 1], s[-B_SCP]), max(s[-B_SCP + 1], s[-1])),
                   max(max(s[1], s[B_SCP - 1]), max(s[B_SCP], s[B_SCP + 1])));
      if (v < mx) return;
    }
#if ORB_B_DIRECT_EMIT
    emit(idx, atomicAdd(gcount, 1));         // one global atomic per survivor (the compiler aggregates a warp's): no survivor
                                             // list, no barriers between NMS, emission and the box sums
#else
    const int slot = atomicAdd(&s_ctr[1], 1);
    if (slot < B_SURV) s_surv[slot] = (uint16_t)idx;
    else emit(idx, atomicAdd(gcount, 1));   // survivor list full: finish this one inline
#endif
  };
  if (!dense) {
    for (int j = tid; j < n1; j += B_THREADS) nms_one(s_list[j]);
  } else {
    for (int i = tid; i < B_TH * B_TW; i += B_THREADS) nms_one((1 + i / B_TW) * B_SP + 16 + (i % B_TW));
  }
#if !ORB_B_DIRECT_EMIT
  __syncthreads();
  const int nsurv = min(s_ctr[1], B_SURV);
  if (nsurv > 0) {
    if (tid == 0) s_ctr[2] = atomicAdd(gcount, nsurv);
    __syncthreads();
    const int base = s_ctr[2];
    for (int j = tid; j < nsurv; j += B_THREADS) emit(s_surv[j], base + j);
  }
#endif

  // ---- phase 5/6: 5x5 box sums (replace the int32 integral image of ref src/orb_cpu.cpp:207-208) and the strip
  // sums for BRIEF boxes that leave the image on the right / bottom (decision D7):
  //   ey[cx] = sum of rows [0,h-4) x cols [cx-2,cx+2]   (ey[0] instead holds column 0 over rows [0,h-4))
  //   rs[y]  = sum of row y over cols [0,w-4)
  // each tile adds its share; k_describe turns them into the values of the reference's wrapped integral taps.
  // Threads 0..127 own an 8-pixel column group and 8 output rows each: horizontal 5-sums of 12 input rows stay in
  // registers (16-bit lanes (o0,o2)(o1,o3)(o4,o6)(o5,o7)), five consecutive ones add up to a box row.  Threads
  // 192..255 produce the row sums meanwhile.
  int* ey = B.cand_count + (size_t)f * B.zero_stride + ORB_MAX_LEVELS + G.edge_ofs;
  int* rs = ey + G.edge_w;
  if (tid < B_TW) {
    const uint32_t M = 0x00ff00ffu;
    const int g = tid & 15, seg = tid >> 4;
    const int yb = y0 + 8 * seg;                                    // first output row of this thread
    if (yb < h && x0 + 8 * g < G.bpitch) {
      uint16_t* box = B.box + (size_t)f * P.box_frame_elems + G.box_ofs + (size_t)yb * G.bpitch + x0 + 8 * g;
      const uint8_t* rc = s_pix + (8 * seg + 2) * B_SP + 16 + 8 * g;   // input row yb - 2
      ORB_CHECK(rc - 4 >= s_pix && rc + 11 * B_SP + 12 <= s_pix + B_PIX_BYTES);
      ORB_CHECK_RANGE(box, 16, B.box + (size_t)f * P.box_frame_elems + G.box_ofs, (size_t)(G.h + 1) * G.bpitch * 2);
      ORB_CHECK(G.edge_ofs + G.edge_w + ((h + 3) & ~3) <= P.edge_frame_elems);
      const int nstrip = min(8, h - 4 - yb);                         // owned rows that lie above row h-4
      uint4 hs[5];
      uint4 strip = make_uint4(0, 0, 0, 0), acc = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int r = 0; r < 12; r++) {
        const uint32_t m = *(const uint32_t*)(rc - 4), p = *(const uint32_t*)(rc + 8);
        const uint2 cc = *(const uint2*)rc;
        rc += B_SP;
        const uint32_t w0 = prmt(m, cc.x, 0x5432), w1 = prmt(cc.x, cc.y, 0x5432), w2 = prmt(cc.y, p, 0x5432);
        const uint32_t q0 = w0 & M, q1 = odd_bytes(w0), q2 = cc.x & M, q3 = odd_bytes(cc.x), q4 = w1 & M, q5 = odd_bytes(w1),
                       q6 = cc.y & M, q7 = odd_bytes(cc.y), q8 = w2 & M, q9 = odd_bytes(w2);
        // horizontal 5-sums from shared pair sums: (o0,o2) = (p-2+p-1, p0+p1) + (p0+p1, p2+p3) + (p2, p4), ...
        const uint32_t t1 = q0 + q1, t2 = q2 + q3, t3 = q4 + q5, t4 = q6 + q7, t5 = q8 + q9;
        uint4 o;
        o.x = t1 + t2 + q4;
        o.y = q1 + t2 + t3;
        o.z = t3 + t4 + q8;
        o.w = q5 + t4 + t5;
        if (r >= 5) { const uint4 old = hs[r % 5]; acc.x -= old.x; acc.y -= old.y; acc.z -= old.z; acc.w -= old.w; }
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;      // sliding sum of the last five rows (16-bit lanes)
        hs[r % 5] = o;
        if (r >= 2 && r - 2 < nstrip) { strip.x += o.x; strip.y += o.y; strip.z += o.z; strip.w += o.w; }
        if (r >= 4 && yb + r - 4 < h) {                              // box row yb + r - 4
          uint4 v;   // lanes (o0,o2)(o1,o3) -> natural order
          v.x = prmt(acc.x, acc.y, 0x5410); v.y = prmt(acc.x, acc.y, 0x7632); v.z = prmt(acc.z, acc.w, 0x5410); v.w = prmt(acc.z, acc.w, 0x7632);
          *(uint4*)(box + (size_t)(r - 4) * G.bpitch) = v;
        }
      }
      if (nstrip > 0) {
        const int val[8] = {(int)(strip.x & 0xffff), (int)(strip.y & 0xffff), (int)(strip.x >> 16), (int)(strip.y >> 16),
                            (int)(strip.z & 0xffff), (int)(strip.w & 0xffff), (int)(strip.z >> 16), (int)(strip.w >> 16)};
#pragma unroll
        for (int j = 0; j < 8; j++) atomicAdd(&s_ey[16 * j + g], val[j]);   // column 8g+j lives at [16j+g]: lanes hit distinct banks
      }
    }
  } else if (tid >= B_THREADS - B_TH) {   // last two warps: row sums over columns < w-4 (and column 0 for the wrapped taps)
    const int iy = tid - (B_THREADS - B_TH), y = y0 + iy;
    if (y < h) {
      const int ncol = min(B_TW, w - 4 - x0);
      const uint8_t* r = s_pix + (iy + 4) * B_SP + 16;
      unsigned acc = 0;
      int c = 0;
      for (; c + 16 <= ncol; c += 16) {              // whole 16-byte groups: four dot products with ones
        const uint4 v = *(const uint4*)(r + c);
        acc = __dp4a(v.x, 0x01010101u, acc); acc = __dp4a(v.y, 0x01010101u, acc);
        acc = __dp4a(v.z, 0x01010101u, acc); acc = __dp4a(v.w, 0x01010101u, acc);
      }
      if (c < ncol) {                                // last, partial group of the level's right edge
        const uint4 v = *(const uint4*)(r + c);
        const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int nb = ncol - c - 4 * k;
          if (nb > 0) acc = __dp4a(nb >= 4 ? wd[k] : (wd[k] & ((1u << (8 * nb)) - 1u)), 0x01010101u, acc);
        }
      }
      ORB_CHECK(y >= 0 && y < h && (iy + 4) * B_SP + 16 + ((ncol + 15) & ~15) <= B_PIX_BYTES);
      if (ncol > 0) atomicAdd(rs + y, (int)acc);
      if (x0 == 0 && y < h - 4) atomicAdd(ey, (int)r[0]);
    }
  }
  __syncthreads();
  if (tid < B_TW) {
    const int x = x0 + tid, v = s_ey[16 * (tid & 7) + (tid >> 3)];
    ORB_CHECK(!(v && x >= 2 && x <= w - 3) || x < G.edge_w);
    if (v && x >= 2 && x <= w - 3) atomicAdd(ey + x, v);
  }
}

// =============================================================================================
// Kernel C: Harris response of every NMS survivor (decision D5), one thread per candidate, grid-stride over the
// candidates of a frame.  Replaces HarrisScore() (ref src/cuda/HarrisScore.cu:42-89, call site src/orb.cpp:65): the
// reference blurs three full-frame product images to read them at <= 2N points; here the 9x9 neighbourhood of each
// candidate is read from the level and the response goes into the high word of its key.  All Sobel arithmetic is exact
// float arithmetic on small integers; only the 147 weighted accumulations round, in the reference order.
#ifndef ORB_C_MINB
#define ORB_C_MINB 8
#endif
constexpr int C_THREADS = 128;

__global__ void __launch_bounds__(C_THREADS, ORB_C_MINB) k_harris(const OrbPlan P, const Bufs B) {
  const int f = blockIdx.y;
  const int* cc = B.cand_count + (size_t)f * B.zero_stride;
  int total = 0;
  for (int q = 0; q < P.nlevels; q++) total += min(cc[q], P.lv[q].cand_cap);
  for (int i = blockIdx.x * C_THREADS + threadIdx.x; i < total; i += gridDim.x * C_THREADS) {
    int l = 0, j = i;
    while (j >= min(cc[l], P.lv[l].cand_cap)) { j -= min(cc[l], P.lv[l].cand_cap); l++; }
    const OrbLevel& G = P.lv[l];
    unsigned long long* slot = B.cand + (size_t)f * P.cand_frame_elems + G.cand_ofs + j;
    const uint32_t xy = (uint32_t)*slot;
    const int x = xy & 0xffff, y = xy >> 16, w = G.w, h = G.h;
    ORB_CHECK(j < G.cand_cap && x >= 3 && x < w - 3 && y >= 3 && y < h - 3);
    const uint8_t* img;
    int pitch;
    if (l == 0) { img = B.frames + (size_t)f * B.frame_stride; pitch = B.pitch0; }
    else { img = B.pyr + (size_t)f * P.pyr_frame_bytes + G.lvl_ofs; pitch = G.pitch; }
    // candidates sit >= 3 pixels inside the level, the Sobel taps reach 4: only the outermost row / column of the
    // 9x9 neighbourhood can leave the level and is reflected (BORDER_REFLECT_101, ref src/Sobel.cpp:29)
    const int dxl = x >= 4 ? -4 : 4 - 2 * x, dxr = x + 4 < w ? 4 : 2 * (w - 1 - x) - 4;   // offsets of columns x-4, x+4
    const int dyt = y >= 4 ? -4 : 4 - 2 * y, dyb = y + 4 < h ? 4 : 2 * (h - 1 - y) - 4;
    const uint8_t* ctr = img + (size_t)y * pitch + x;
    float r;
    if (x >= 4 && x + 4 < w) {
      // columns x-4 .. x+4 of a row lie in three aligned words (rows are 16-byte aligned and pitch >= w, so the third
      // word, which holds column x+4 < w, is inside the row): 3 loads + 3 funnel shifts per row instead of 9 byte loads
      const int sh = 8 * ((x - 4) & 3);
      const uint8_t* base = ctr - 4 - ((x - 4) & 3);
      r = harris_at([&](int dy, float (&v)[9]) {
        const int oy = dy == -4 ? dyt : (dy == 4 ? dyb : dy);
        const uint32_t* q = (const uint32_t*)(base + oy * pitch);
        ORB_CHECK(((uintptr_t)q & 3) == 0 && (const uint8_t*)q >= img && (const uint8_t*)q + 12 <= img + (size_t)h * pitch && y + oy >= 0 && y + oy < h);
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
        const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh), c = w2 >> sh;
        const uint32_t F = 0x4B000000u;              // byte permutes below: (pixel byte, 0x00, 0x00, 0x4B)
        v[0] = byte2float(a, F, 0x7440); v[1] = byte2float(a, F, 0x7441); v[2] = byte2float(a, F, 0x7442); v[3] = byte2float(a, F, 0x7443);
        v[4] = byte2float(b, F, 0x7440); v[5] = byte2float(b, F, 0x7441); v[6] = byte2float(b, F, 0x7442); v[7] = byte2float(b, F, 0x7443);
        v[8] = byte2float(c, F, 0x7440);
      }, P.harris_k);
    } else {
      r = harris_at([&](int dy, float (&v)[9]) {
        const int oy = dy == -4 ? dyt : (dy == 4 ? dyb : dy);
        const uint8_t* q = ctr + oy * pitch;
        ORB_CHECK(y + oy >= 0 && y + oy < h && x + dxl >= 0 && x + dxr < w && x - 3 >= 0 && x + 3 < w);
        v[0] = (float)q[dxl];
#pragma unroll
        for (int c = 1; c < 8; c++) v[c] = (float)q[c - 4];
        v[8] = (float)q[dxr];
      }, P.harris_k);
    }
    *slot = ((unsigned long long)(~f2ord(r)) << 32) | xy;
  }
}

// ---------------------------------------------------------------------------------------------
// Selection: CTA per (frame, level).
constexpr int K2_THREADS = 512;
constexpr int K2_SMEM_KEYS = 6144;   // candidate keys of a level are staged in shared memory when they fit
constexpr int K2_MAX_ROWS = 4096;    // levels up to this height put their kept keys in raster order by a counting sort over rows

__global__ void __launch_bounds__(K2_THREADS) k_select(const OrbPlan P, const Bufs B, int npow2_max) {
  extern __shared__ __align__(16) unsigned long long s_dyn[];   // [npow2_max] sort buffer, [K2_SMEM_KEYS] keys, [npow2_max] row groups, [h + 1] row counters
  unsigned long long* s_sort = s_dyn;
  unsigned long long* s_keys = s_dyn + npow2_max;
  __shared__ int s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_k, s_n, s_done;
  const int tid = threadIdx.x, lane = tid & 31, l = blockIdx.x, f = blockIdx.y;
  const OrbLevel& G = P.lv[l];
  int n = B.cand_count[(size_t)f * B.zero_stride + l];
  if (n > G.cand_cap) {
    if (tid == 0) atomicOr(B.flags, 1);
    n = G.cand_cap;
  }
  const int m = min(G.quota, n);
  const unsigned long long* __restrict__ gkeys = B.cand + (size_t)f * P.cand_frame_elems + G.cand_ofs;
  const unsigned long long* keys = gkeys;
  unsigned long long T = ~0ull;
  if (n > m && m > 0) {
    if (n <= K2_SMEM_KEYS) {
      for (int i = tid; i < n; i += K2_THREADS) s_keys[i] = gkeys[i];
      keys = s_keys;
    }
    // m-th smallest key by MSB-first radix select, 8 bits per pass.  Lanes with equal digits are aggregated
    // (match_any) before the shared atomic; the scan stops as soon as the whole bin is needed.
    if (tid == 0) { s_prefix = 0; s_k = m; s_done = 0; }
    __syncthreads();
    for (int pass = 7; pass >= 0; pass--) {
      for (int i = tid; i < 256; i += K2_THREADS) s_hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const int shift = pass * 8;
      for (int i0 = 0; i0 < n; i0 += K2_THREADS) {
        const int i = i0 + tid;
        int digit = -1;
        if (i < n) {
          const unsigned long long k = keys[i];
          if (pass == 7 || (k >> (shift + 8)) == prefix) digit = (int)(k >> shift) & 255;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        if (digit >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
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
        int sel = __ffs(hit) - 1;
        if (tid == sel) {
          int before = incl - sum, b = 0;
          for (; b < 8; b++) { if (before + local[b] >= k) break; before += local[b]; }
          const unsigned long long np = (prefix << 8) | (unsigned)(tid * 8 + b);
          if (k - before == local[b]) {        // every key of this bin is kept: threshold = largest key of the bin
            s_prefix = shift ? ((np << shift) | ((1ull << shift) - 1ull)) : np;
            s_done = 1;
          } else {
            s_k = k - before;
            s_prefix = np;
          }
        }
      }
      __syncthreads();
      if (s_done) break;
    }
    T = s_prefix;
  }
  // gather the kept keys as (raster << 32 | response order) and put them in raster order
  if (tid == 0) s_n = 0;
  int npow2 = 1;
  while (npow2 < m) npow2 <<= 1;
  const bool by_rows = G.h <= K2_MAX_ROWS;       // counting sort over rows (below); bitonic sort for taller levels
  int* s_row = (int*)(s_dyn + npow2_max + K2_SMEM_KEYS + (by_rows ? npow2_max : 0));   // [h + 1]
  unsigned long long* s_out = s_dyn + npow2_max + K2_SMEM_KEYS;                        // [m] keys grouped by row
  if (by_rows) for (int i = tid; i <= G.h; i += K2_THREADS) s_row[i] = 0;
  else for (int i = tid; i < npow2; i += K2_THREADS) s_sort[i] = ~0ull;
  __syncthreads();
  if (m > 0)
    for (int i = tid; i < n; i += K2_THREADS) {
      unsigned long long k = keys[i];
      if (k <= T) {
        int pos = atomicAdd(&s_n, 1);
        ORB_CHECK(pos < m && npow2 <= npow2_max);
        if (pos < npow2) s_sort[pos] = (k << 32) | (k >> 32);
        if (by_rows) atomicAdd(&s_row[(int)((k >> 16) & 0xffff)], 1);
      }
    }
  __syncthreads();
  uint32_t* kxy = B.kept_xy + (size_t)f * P.kept_per_frame + G.kept_ofs;
  float* kr = B.kept_r + (size_t)f * P.kept_per_frame + G.kept_ofs;
  ORB_CHECK(G.kept_ofs + m <= P.kept_per_frame && n <= G.cand_cap && (n <= K2_SMEM_KEYS || keys == gkeys));
  const bool harris = P.select_policy == ORB_SELECT_HARRIS_TOP_N;
  if (by_rows) {
    // rows are few (<= K2_MAX_ROWS) and hold a handful of kept keys each: exclusive prefix over the row counts, scatter
    // into row groups, then every key finds its place inside its row by counting the smaller ones -- five barriers
    // instead of the ~45 of a bitonic sort of 512 keys
    constexpr int RPT = (K2_MAX_ROWS + K2_THREADS - 1) / K2_THREADS;     // rows per thread
    int loc[RPT], sum = 0;
#pragma unroll
    for (int j = 0; j < RPT; j++) { const int r = tid * RPT + j; loc[j] = r < G.h ? s_row[r] : 0; sum += loc[j]; }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_hist[tid >> 5] = incl;                             // (s_hist is free after the radix passes)
    __syncthreads();
    int wbase = 0;
    for (int wi = 0; wi < (tid >> 5); wi++) wbase += s_hist[wi];
    int run = wbase + incl - sum;
    __syncthreads();                                                     // every count has been read
#pragma unroll
    for (int j = 0; j < RPT; j++) { const int r = tid * RPT + j; if (r < G.h) s_row[r] = run; run += loc[j]; }   // cursor = first slot of the row
    __syncthreads();
    for (int i = tid; i < m; i += K2_THREADS) {
      const unsigned long long v = s_sort[i];
      s_out[atomicAdd(&s_row[(int)(v >> 48)], 1)] = v;                   // afterwards s_row[y] = end of row y = start of row y + 1
    }
    __syncthreads();
    for (int i = tid; i < m; i += K2_THREADS) {
      const unsigned long long v = s_out[i];
      const int y = (int)(v >> 48), b0 = y ? s_row[y - 1] : 0, e0 = s_row[y];
      int pos = b0;
      for (int j = b0; j < e0; j++) pos += s_out[j] < v;
      ORB_CHECK(pos < m && b0 <= i && i < e0);
      kxy[pos] = (uint32_t)(v >> 32);
      kr[pos] = harris ? ord2f(~(uint32_t)v) : 0.0f;
    }
  } else {
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
    for (int i = tid; i < m; i += K2_THREADS) {
      unsigned long long v = s_sort[i];
      kxy[i] = (uint32_t)(v >> 32);
      kr[i] = harris ? ord2f(~(uint32_t)v) : 0.0f;
    }
  }
  if (tid == 0) B.kept_count[f * ORB_MAX_LEVELS + l] = m;
}

// ---------------------------------------------------------------------------------------------
// Orientation + rotated BRIEF.
#ifndef ORB_BRIEF_UNROLL
#define ORB_BRIEF_UNROLL 8
#endif
constexpr int BRIEF_UNROLL = ORB_BRIEF_UNROLL;
#ifndef ORB_ORIENT_UNROLL
#define ORB_ORIENT_UNROLL 31
#endif
constexpr int ORIENT_UNROLL = ORB_ORIENT_UNROLL;
#ifndef ORB_K3_WARPS
#define ORB_K3_WARPS 4
#endif
#ifndef ORB_K3_WARP_LIBM
#define ORB_K3_WARP_LIBM 1
#endif
#ifndef ORB_K3_MINB
#define ORB_K3_MINB 7
#endif
constexpr int K3_WARPS = ORB_K3_WARPS;

struct EdgeSrc {
  const uint8_t* __restrict__ img; int pitch;
  int W, H;
  const int* __restrict__ ey;   // ey[cx]: rows [0,H-4) x cols [cx-2,cx+2]; ey[0]: column 0 over rows [0,H-4)
  const int* __restrict__ rs;   // rs[y]: row y over cols [0,W-4)
};

// Value tables for BRIEF boxes whose centre lies in the last two columns / rows of a level (decision D7), filled by
// k_edges from the strip tables: what the reference's four wrapped integral taps add up to.
//   bot[a * edge_w + cx]  : cy = H-2+a, 2 <= cx <= W-3      right[b * hh + cy] : cx = W-2+b, 2 <= cy <= H-3
//   corner[a * 2 + b]     : cy = H-2+a, cx = W-2+b          (hh = h rounded up to a multiple of 4)
__host__ __device__ inline int edge2_level_elems(int edge_w, int h) { return 2 * edge_w + 2 * ((h + 3) & ~3) + 4; }
struct EdgeTab {
  const int* __restrict__ tab; int W, H, edge_w, hh;
  __device__ __forceinline__ int at(int cx, int cy) const {
    const int a = cy - (H - 2), b = cx - (W - 2);
    const int idx = a >= 0 ? (b >= 0 ? 2 * edge_w + 2 * hh + 2 * a + b : a * edge_w + cx) : 2 * edge_w + b * hh + cy;
    ORB_CHECK(idx >= 0 && idx < edge2_level_elems(edge_w, H) && cx >= 2 && cx <= W - 1 && cy >= 2 && cy <= H - 1);
    return __ldg(tab + idx);
  }
};


// ---- BRIEF boxes that leave the image on the right / bottom ---------------------------------------
// The reference's sum5x5 (src/orb_cpu.cpp:190-201) indexes its (H+1)x(W+1) integral image flat, so for centres
// in the last two columns / rows the column overruns wrap into the next row and the row overruns fall off the
// end (decision D7: those read 0).  What its four taps then add up to are *strip* sums of the image:
//   bottom rows only : -(rows [0,cy-2) x cols [cx-2,cx+2])
//   right cols only  : -(rows [cy-2,cy+2] x cols [0,cx-2)) + wrapped column-0 taps (when cx == W-1)
//   corner           : +(rows [0,cy-2) x cols [0,cx-2)) - column-0 prefix
// The first two come from the strip tables k_fast accumulates (+ at most 15 pixels); k_edges writes them out per level.

// bottom-right corner box (cx > W-3 and cy > H-3), whole warp cooperates:
//   +(rows [0,cy-2) x cols [0,cx-2)) - (column 0 over rows [0,cy-1) when cx == W-1)
__device__ __noinline__ int box_corner(const EdgeSrc& E, int cx, int cy, int lane) {
  const int W = E.W, H = E.H, y0 = cy - 2;
  int s = 0;
#pragma unroll 4
  for (int v = lane; v < y0; v += 32) {
    s += E.rs[v];
    if (cx == W - 1) s += E.img[(size_t)v * E.pitch + (W - 4)];
  }
  s = warp_sum(s);
  if (cx == W - 1) {
    s -= E.ey[0] + E.img[(size_t)(H - 4) * E.pitch];
    if (y0 == H - 3) s -= E.img[(size_t)(H - 3) * E.pitch];
  }
  return s;
}

struct DescribeJob {             // where the keypoints of this launch come from
  int mode;                      // 0: kept lists of the pipeline; 1: explicit list, orientation only;
                                 // 2: explicit list + given angles, descriptors only
  const orb_keypoint* list_kps;  // modes 1, 2
  const float* list_angles;      // mode 2
  int list_n;
};

// Kernel E: the border-box tables of every level of every frame of the wave (after k_fast has finished the strip
// tables).  One CTA per (level, frame): ~2 (W + H) line entries spread over the threads, then one warp per corner entry.
constexpr int E_THREADS = 256;
__global__ void __launch_bounds__(E_THREADS) k_edges(const OrbPlan P, const Bufs B) {
  const int l = blockIdx.x, f = blockIdx.y;
  const OrbLevel& G = P.lv[l];
  const int W = G.w, H = G.h, ew = G.edge_w, hh = (H + 3) & ~3;
  const int n_line = 2 * ew + 2 * hh;
  const uint8_t* img;
  int pitch;
  if (l == 0) { img = B.frames + (size_t)f * B.frame_stride; pitch = B.pitch0; }
  else { img = B.pyr + (size_t)f * P.pyr_frame_bytes + G.lvl_ofs; pitch = G.pitch; }
  const int* ey = B.cand_count + (size_t)f * B.zero_stride + ORB_MAX_LEVELS + G.edge_ofs;
  const EdgeSrc E{img, pitch, W, H, ey, ey + ew};
  int* out = B.edge2 + (size_t)f * P.edge2_frame_elems + G.edge2_ofs;
  const bool big = W >= 7 && H >= 7;             // smaller levels cannot hold a keypoint (FAST needs a 3-pixel margin)
  ORB_CHECK_COUNT();
  ORB_CHECK(G.edge2_ofs + n_line + 4 <= P.edge2_frame_elems && G.edge_ofs + ew + hh <= P.edge_frame_elems && W <= ew);
  // bottom rows (cy = H-2, H-1; 2 <= cx <= W-3): -(strip above the box), the last row adds the five pixels of row H-4
  const uint8_t* row = img + (size_t)max(H - 4, 0) * pitch;
#pragma unroll 2
  for (int cx = threadIdx.x; cx < ew; cx += E_THREADS) {
    int v0 = 0, v1 = 0;
    if (big && cx >= 2 && cx <= W - 3) {
      const int s = E.ey[cx];
      v0 = -s;
      v1 = -(s + row[cx - 2] + row[cx - 1] + row[cx] + row[cx + 1] + row[cx + 2]);
    }
    out[cx] = v0;
    out[ew + cx] = v1;
  }
  // right columns (cx = W-2, W-1; 2 <= cy <= H-3): -(strip left of the box); the last column adds column W-4 and the
  // wrapped column-0 taps (or, when those fall off the integral image, the column-0 strip)
  const int ey0 = big ? E.ey[0] : 0;
#pragma unroll 2
  for (int cy = threadIdx.x; cy < hh; cy += E_THREADS) {
    int v0 = 0, v1 = 0;
    if (big && cy >= 2 && cy <= H - 3) {
      const int s = E.rs[cy - 2] + E.rs[cy - 1] + E.rs[cy] + E.rs[cy + 1] + E.rs[cy + 2];
      const uint8_t* c = img + (size_t)(cy - 2) * pitch + (W - 4);
      const int s1 = s + c[0] + c[pitch] + c[2 * pitch] + c[3 * pitch] + c[4 * pitch];
      int t;
      if (cy + 4 <= H) { const uint8_t* z = img + (size_t)(cy - 1) * pitch; t = z[0] + z[pitch] + z[2 * pitch] + z[3 * pitch] + z[4 * pitch]; }
      else t = -ey0;
      v0 = -s;
      v1 = t - s1;
    }
    out[2 * ew + cy] = v0;
    out[2 * ew + hh + cy] = v1;
  }
  if (threadIdx.x < 4 * 32) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;      // corner c: cy = H-2 + (c >> 1), cx = W-2 + (c & 1)
    const int v = big ? box_corner(E, W - 2 + (c & 1), H - 2 + (c >> 1), lane) : 0;
    if (lane == 0) out[n_line + c] = v;
  }
}

// ---- per-keypoint windows in shared memory -----------------------------------------------------------------------
// A warp copies what a keypoint needs into its own shared-memory slots with 16-byte cp.async copies, ahead of the
// keypoint it is working on, so that the global-memory latency is off the critical path and the gathers are shared-
// memory loads:
//   orientation : the (2r+1)^2 patch as rows of W_PP bytes (the 16-byte aligned superset of columns kx-r .. kx+r);
//   BRIEF       : rotated offsets stay within +-19 (pattern radius 18.4), so the 512 box sums of a keypoint lie in the
//                 39 x 39 window around it: 39 rows x 48 elements of the level's u16 box-sum image.
// The two uses are in different phases of the kernel and share the warp's buffer.
constexpr int W_BW = 48, W_BH = 39, W_SLOT = W_BW * W_BH * 2;   // 3744 bytes per box-window slot
constexpr int W_DEPTH = 2;                                        // box-window slots per warp
constexpr int W_WARP_BYTES = W_DEPTH * W_SLOT;                    // 7488 bytes per warp
constexpr int W_PP15 = 48, W_PDEPTH15 = 4;                        // patch 31: 31 rows x 48 bytes, four slots of 1536 bytes
constexpr int W_PSLOT15 = 1536;
constexpr int W_PPGEN = 80;                                       // other radii (<= 31): up to 63 rows x 80 bytes, one slot
static_assert(W_PDEPTH15 * W_PSLOT15 <= W_WARP_BYTES && 63 * W_PPGEN <= W_WARP_BYTES, "patch slots live in the box-window buffer");

__device__ __forceinline__ void cp_async_16(uint32_t sa, const void* g, bool ok, const void* safe) {
  // rows / columns outside the image are zero-filled (source size 0, address kept valid): they are never used
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(ok ? g : safe), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// box-sum window of keypoint (kx, ky): lane -> (row lane / 6 of a group of five rows, 16-byte column lane % 6);
// lanes 30, 31 idle; eight groups = 40 rows
__device__ __forceinline__ void window_issue(uint32_t slot_sa, const uint16_t* __restrict__ box, int bpitch, int h, int kx, int ky,
                                             int lane) {
  const int r0 = lane / 6, c0 = lane - 6 * r0;
  const int xa = ((kx - 19) & ~7) + 8 * c0, y0 = ky - 19 + r0;
  if (lane >= 30) return;
  const bool col_ok = xa >= 0 && xa < bpitch;
  uint32_t sa = slot_sa + (uint32_t)(r0 * (W_BW * 2) + c0 * 16);
  ORB_CHECK(r0 * (W_BW * 2) + c0 * 16 + 7 * 5 * (W_BW * 2) + 16 <= W_SLOT + (r0 < W_BH - 35 ? 0 : 5 * (W_BW * 2)));
  if (col_ok && ky - 19 >= 0 && ky + 19 < h) {   // this lane's column and all 39 rows lie inside the level
    const uint16_t* g = box + (ptrdiff_t)y0 * bpitch + xa;
    const size_t step = (size_t)(5 * bpitch) * 2;
    ORB_CHECK_RANGE(g, 16, box, (size_t)h * bpitch * 2);
    ORB_CHECK_RANGE((const char*)g + (r0 < W_BH - 35 ? 7 : 6) * step, 16, box, (size_t)h * bpitch * 2);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j < 7 || r0 < W_BH - 35)               // (the 40th row of the last group is not part of the window)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + j * 5 * (W_BW * 2)), "l"((const char*)g + j * step) : "memory");
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int y = y0 + 5 * j;
      ORB_CHECK(!(col_ok && y >= 0 && y < h) || (xa + 8 <= bpitch && (size_t)y * bpitch + xa + 8 <= (size_t)h * bpitch));
      if (j < 7 || r0 < W_BH - 35) cp_async_16(sa + j * 5 * (W_BW * 2), box + (ptrdiff_t)y * bpitch + xa, col_ok && y >= 0 && y < h, box);
    }
  }
}

// orientation patch of a keypoint whose patch lies inside the level (ref src/orb_cpu.cpp:152-156 sets the angle of the
// others to 0): rows ky-pr .. ky+pr, bytes from the 16-byte aligned column at or below kx-pr, PP bytes per row.
// PP == 48 (patch 31): lane -> (row lane / 3 of a group of ten rows, 16-byte column lane % 3), four groups = 40 rows.
template <int PP>
__device__ __forceinline__ void patch_issue(uint32_t slot_sa, const uint8_t* __restrict__ img, int pitch, int kx, int ky, int pr, int lane) {
  constexpr int CPR = PP / 16, ROWS_PER = 30 / CPR;
  const int r0 = lane / CPR, c0 = lane - CPR * r0;
  const int xa = ((kx - pr) & ~15) + 16 * c0;
  if (lane >= 30) return;
  const bool ok = xa < pitch;
  const uint8_t* g = ok ? img + (ptrdiff_t)(ky - pr + r0) * pitch + xa : img;
  const size_t step = ok ? (size_t)ROWS_PER * pitch : 0;
  const uint32_t sa = slot_sa + (uint32_t)(r0 * PP + c0 * 16), nb = ok ? 16u : 0u;
  ORB_CHECK(ky - pr >= 0 && kx - pr >= 0 && ((kx - pr) & 15) + 2 * pr + 1 <= PP && (!ok || xa + 16 <= pitch));
  if (PP == W_PP15) {                            // pr == 15: 31 rows
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < 3 || r0 == 0)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa + j * ROWS_PER * PP), "l"(g + j * step), "r"(nb) : "memory");
  } else {
    for (int r = r0, j = 0; r <= 2 * pr; r += ROWS_PER, j++)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa + j * ROWS_PER * PP), "l"(g + j * step), "r"(nb) : "memory");
  }
}

// Intensity-centroid moments of the (2*pr+1)^2 patch, ref src/orb_cpu.cpp:158-176, from the staged patch.  Moments are
// exact integers (|m| < 2^24), so integer accumulation in any order equals the reference's float accumulation.
// Lane j owns patch column j - pr (and j + 32 - pr): m10 = sum_c c * colsum(c), m01 = sum_r r * I.  For patch 31 one
// multiply-add per pixel accumulates both sums: I * (1 + (r' << 13)) with r' = r + 15 keeps the column sum
// (<= 31 * 255 < 2^13) in the low 13 bits and sum r' * I (<= 465 * 255 < 2^17) above it.
__device__ __forceinline__ void patch_moments(uint32_t slot_sa, int kx, int pr, int pp, int lane, int* m10_out, int* m01_out) {
  int m10 = 0, m01 = 0;
  const uint32_t base = slot_sa + (uint32_t)((kx - pr) & 15);
  ORB_CHECK(((kx - pr) & 15) + 2 * pr + 1 <= pp && (2 * pr + 1) * pp <= W_WARP_BYTES);
  if (pr == 15) {                                                 // patch 31 (include/orb.hpp:12): one column per lane
    if (lane <= 30) {
      uint32_t acc = 0;
#pragma unroll
      for (int r = 0; r <= 30; r++) acc += lds_u8(base + lane + r * W_PP15) * (1u + ((uint32_t)r << 13));
      const int colsum = (int)(acc & 0x1fffu);
      m01 = (int)(acc >> 13) - 15 * colsum;
      m10 = (lane - 15) * colsum;
    }
  } else {
    for (int c = lane - pr; c <= pr; c += 32) {
      int colsum = 0;
      for (int r = -pr; r <= pr; r++) {
        const int I = (int)lds_u8(base + (uint32_t)((r + pr) * pp + c + pr));
        colsum += I;
        m01 += r * I;
      }
      m10 += c * colsum;
    }
  }
  *m10_out = __reduce_add_sync(0xffffffffu, m10);   // REDUX.SUM: one instruction per warp sum
  *m01_out = __reduce_add_sync(0xffffffffu, m01);
}

// rotate one pattern point and round (ref src/orb_cpu.cpp:228-232): (lround(c*x - s*y), lround(s*x + c*y)).  The two
// products per coordinate pair are packed f32x2 multiplies (each half rounded like the scalar multiply); the sums
// stay scalar adds so that nothing can contract into an FMA.
__device__ __forceinline__ void rotate_round(unsigned long long cs, unsigned long long msc, float px, float py, int* dx, int* dy) {
  unsigned long long a, b;
  const unsigned long long xx = ((unsigned long long)__float_as_uint(px) << 32) | __float_as_uint(px);
  const unsigned long long yy = ((unsigned long long)__float_as_uint(py) << 32) | __float_as_uint(py);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(cs), "l"(xx));      // (c * x, s * x)
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(msc), "l"(yy));     // (-s * y, c * y)
  const float vx = __fadd_rn(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)b));                // c*x - s*y
  const float vy = __fadd_rn(__uint_as_float((uint32_t)(a >> 32)), __uint_as_float((uint32_t)(b >> 32)));  // s*x + c*y
  *dx = orbm::lround_f(vx);
  *dy = orbm::lround_f(vy);
}

// wc_sa: shared-memory byte address of the window element that holds the box sum at the keypoint itself, i.e. the box
// sum at (kx + dx, ky + dy) is the u16 at wc_sa + dy * 96 + dx * 2
__device__ __forceinline__ void brief_of(uint32_t wc_sa, const EdgeTab& T, int kx, int ky, float c, float s,
                                         const float4* __restrict__ pattern, int lane, uint32_t* out_words) {
  const int W = T.W, H = T.H;   // c, s = cos / sin of the keypoint angle (ref src/orb_cpu.cpp:217-218)
  uint32_t mine = 0;
  const unsigned long long cs = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(c);
  const unsigned long long msc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(-s);
  // if every box is interior the bound rule of src/orb_cpu.cpp:240-245 can never fire and the sums are plain lookups
  const bool interior = kx - 19 >= 2 && kx + 19 <= W - 3 && ky - 19 >= 2 && ky + 19 <= H - 3;
  if (interior) {
#pragma unroll (BRIEF_UNROLL)
    for (int wd = 0; wd < 8; wd++) {
      const float4 t = __ldg(pattern + wd * 32 + lane);
      int dx1, dy1, dx2, dy2;
      rotate_round(cs, msc, t.x, t.y, &dx1, &dy1);
      rotate_round(cs, msc, t.z, t.w, &dx2, &dy2);
      ORB_CHECK(abs(dx1) <= 19 && abs(dy1) <= 19 && abs(dx2) <= 19 && abs(dy2) <= 19);      // the window holds offsets -19 .. 19
      const uint32_t s1 = lds_u16(wc_sa + dy1 * (W_BW * 2) + dx1 * 2), s2 = lds_u16(wc_sa + dy2 * (W_BW * 2) + dx2 * 2);
      const uint32_t word = __ballot_sync(0xffffffffu, s1 < s2);   // bit i of word wd == test 32*wd + i
      if (lane == wd) mine = word;
    }
    *out_words = mine;
    return;
  }
  // Keypoint near a border.  The bound rule of :240-245 (against the integral image dims W+1, H+1) skips a test when a
  // box centre has cx < 2, cy < 2, cx > W-1 or cy > H-1: one unsigned compare per coordinate.  Centres in the last two
  // columns / rows (decision D7) take their value from the level's border-box tables (k_edges) instead of the window.
  const uint32_t xlim = (uint32_t)(W - 3), ylim = (uint32_t)(H - 3);
#pragma unroll 2
  for (int wd = 0; wd < 8; wd++) {
    const float4 t = __ldg(pattern + wd * 32 + lane);
    int dx1, dy1, dx2, dy2;
    rotate_round(cs, msc, t.x, t.y, &dx1, &dy1);
    rotate_round(cs, msc, t.z, t.w, &dx2, &dy2);
    const int cx1 = kx + dx1, cy1 = ky + dy1, cx2 = kx + dx2, cy2 = ky + dy2;
    const bool skip = (uint32_t)(cx1 - 2) > xlim || (uint32_t)(cy1 - 2) > ylim || (uint32_t)(cx2 - 2) > xlim || (uint32_t)(cy2 - 2) > ylim;
    const bool e1 = !skip && (cx1 > W - 3 || cy1 > H - 3), e2 = !skip && (cx2 > W - 3 || cy2 > H - 3);
    ORB_CHECK(abs(dx1) <= 19 && abs(dy1) <= 19 && abs(dx2) <= 19 && abs(dy2) <= 19);
    int s1 = (int)lds_u16(wc_sa + dy1 * (W_BW * 2) + dx1 * 2), s2 = (int)lds_u16(wc_sa + dy2 * (W_BW * 2) + dx2 * 2);
    if (e1) s1 = T.at(cx1, cy1);
    if (e2) s2 = T.at(cx2, cy2);
    uint32_t word = __ballot_sync(0xffffffffu, !skip && s1 < s2);
    if (lane == wd) mine = word;
  }
  *out_words = mine;
}

// A CTA of K3_WARPS warps owns K3_KPS = 32 consecutive keypoints of one frame and works in phases so that the scalar
// work runs once per lane for 32 keypoints instead of once per warp and keypoint:
//   (0) lane = keypoint: level / position lookup, image and table addresses of its level -> shared memory;
//   (1) warp per keypoint: integer patch moments out of staged patches (four keypoints ahead);
//   (2) warp 0, lane = keypoint: angle (libm: ~250 instructions), cos, sin, level-0 coordinates, record headers;
//   (3) warp per keypoint: rotated BRIEF out of staged box-sum windows (two keypoints ahead; the first two are
//       requested before phase 2).
constexpr int K3_KPS = 32;

struct KpSlot {                  // what phase 0 resolves per keypoint
  const uint8_t* img;            // level image
  const uint16_t* box;           // level box-sum image
  const int* tab;                // border-box tables of the level (k_edges)
  int pitch, bpitch, w, h, edge_w;
  int x, y, l;                   // level-space position, level (| 0x100: orientation patch leaves the level -> angle 0)
  union { int m10; float c; };   // moments (phase 1) -> cos / sin of the angle (phase 2)
  union { int m01; float s; };
};

__global__ void __launch_bounds__(K3_WARPS * 32, ORB_K3_MINB) k_describe(const OrbPlan P, const Bufs B, const DescribeJob J) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.y;
  const int k0 = blockIdx.x * K3_KPS;            // first keypoint (output slot) of this CTA
  __shared__ __align__(16) uint8_t s_win[K3_WARPS * W_WARP_BYTES];
  __shared__ int s_pref[ORB_MAX_LEVELS + 1];     // exclusive prefix of the frame's kept counts (levels are concatenated)
  __shared__ KpSlot s_kp[K3_KPS];
  int total;
  if (J.mode == 0) {
    if (threadIdx.x < 32) {
      const int c = lane < P.nlevels ? B.kept_count[f * ORB_MAX_LEVELS + lane] : 0;
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {                     // full 32-lane scan: lane 16 must see level 0 when nlevels == 16
        int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      if (lane <= ORB_MAX_LEVELS) s_pref[lane] = incl - c;   // lane == nlevels .. 16 hold the total
    }
    __syncthreads();
    total = min(s_pref[P.nlevels], B.out_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) B.out_n[f] = total;
  } else {
    total = J.list_n;
  }
  if (k0 >= total) return;
  const int n_cta = min(K3_KPS, total - k0);     // keypoints of this CTA; warp w owns q = w, w + K3_WARPS, ...
  ORB_CHECK_COUNT();
  const int pr = P.patch_radius;

  // ---- phase 0: lane = keypoint ------------------------------------------------------------------
  if (threadIdx.x < n_cta) {
    const int widx = k0 + threadIdx.x;
    int l = 0, x, y;
    if (J.mode == 0) {
      while (l + 1 < P.nlevels && widx >= s_pref[l + 1]) l++;
      const uint32_t xy = B.kept_xy[(size_t)f * P.kept_per_frame + P.lv[l].kept_ofs + (widx - s_pref[l])];
      x = xy & 0xffff; y = xy >> 16;
    } else {
      x = J.list_kps[widx].x; y = J.list_kps[widx].y;
    }
    const OrbLevel& G = P.lv[l];
    KpSlot k;
    if (l == 0) { k.img = B.frames + (size_t)f * B.frame_stride; k.pitch = B.pitch0; }
    else { k.img = B.pyr + (size_t)f * P.pyr_frame_bytes + G.lvl_ofs; k.pitch = G.pitch; }
    k.box = B.box + (size_t)f * P.box_frame_elems + G.box_ofs;
    k.tab = B.edge2 + (size_t)f * P.edge2_frame_elems + G.edge2_ofs;
    k.bpitch = G.bpitch; k.w = G.w; k.h = G.h; k.edge_w = G.edge_w;
    k.x = x; k.y = y;
    ORB_CHECK(x >= 0 && x < G.w && y >= 0 && y < G.h && G.edge2_ofs + edge2_level_elems(G.edge_w, G.h) <= P.edge2_frame_elems);
    // the reference sets the angle to 0 when the patch leaves the level (src/orb_cpu.cpp:152-156)
    k.l = (x - pr < 0 || x + pr >= G.w || y - pr < 0 || y + pr >= G.h) ? (l | 0x100) : l;
    k.m10 = 0; k.m01 = 0;
    s_kp[threadIdx.x] = k;
  }
  __syncthreads();
  const int n_mine = warp < n_cta ? (n_cta - warp + K3_WARPS - 1) / K3_WARPS : 0;
  const uint32_t my_sa = smem_u32(s_win) + (uint32_t)(warp * W_WARP_BYTES);

  // ---- phase 1: patch moments, one warp per keypoint, patches staged ahead ---------------------------
  if (J.mode != 2) {
    const int pp = pr == 15 ? W_PP15 : W_PPGEN, pdepth = pr == 15 ? W_PDEPTH15 : 1, pslot = pr == 15 ? W_PSLOT15 : 0;
    auto issue_patch = [&](int i) {
      if (i < n_mine) {
        const KpSlot& k = s_kp[warp + K3_WARPS * i];
        if (!(k.l & 0x100)) {
          if (pr == 15) patch_issue<W_PP15>(my_sa + (uint32_t)((i % W_PDEPTH15) * W_PSLOT15), k.img, k.pitch, k.x, k.y, 15, lane);
          else patch_issue<W_PPGEN>(my_sa, k.img, k.pitch, k.x, k.y, pr, lane);
        }
      }
      cp_async_commit();                         // (possibly empty) group i
    };
    for (int i = 0; i < pdepth; i++) issue_patch(i);
    for (int i = 0; i < n_mine; i++) {
      const int q = warp + K3_WARPS * i;
      if (pr == 15) cp_async_wait<W_PDEPTH15 - 1>(); else cp_async_wait<0>();
      __syncwarp();
      if (!(s_kp[q].l & 0x100)) {
        int m10, m01;
        patch_moments(my_sa + (uint32_t)((i % pdepth) * pslot), s_kp[q].x, pr, pp, lane, &m10, &m01);
        if (lane == 0) { s_kp[q].m10 = m10; s_kp[q].m01 = m01; }
      }
      __syncwarp();
      issue_patch(i + pdepth);
    }
    cp_async_wait<0>();
  }
  // the box-sum windows of this warp's first keypoints travel while warp 0 does the libm part
  auto issue_window = [&](int i) {
    if (i < n_mine) {
      const KpSlot& k = s_kp[warp + K3_WARPS * i];
      window_issue(my_sa + (uint32_t)((i % W_DEPTH) * W_SLOT), k.box, k.bpitch, k.h, k.x, k.y, lane);
    }
    cp_async_commit();
  };
  if (J.mode != 1) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < W_DEPTH; i++) issue_window(i);
  }
#if ORB_K3_WARP_LIBM
  __syncwarp();

  // ---- phase 2: lane = one of the warp's keypoints: angle (glibc-exact atan2f), cos / sin, record headers.  The
  // ~250 scalar instructions run once per warp for its <= 8 keypoints; in exchange no warp waits at a CTA barrier.
  if (lane < n_mine) {
    const int q2 = warp + K3_WARPS * lane;
    KpSlot& k = s_kp[q2];
    const int widx = k0 + q2, x = k.x, y = k.y, l = k.l & 0xff;
#else
  __syncthreads();

  // ---- phase 2: lane = keypoint: angle (glibc-exact atan2f), cos / sin, record headers -----------
  if (warp == 0 && lane < n_cta) {
    KpSlot& k = s_kp[lane];
    const int widx = k0 + lane, x = k.x, y = k.y, l = k.l & 0xff;
#endif
    const size_t o = (size_t)f * B.out_cap + widx;
    float angle;
    if (J.mode == 2) angle = J.list_angles[widx];
    else angle = (k.l & 0x100) ? 0.0f : orbm::atan2f_glibc((float)k.m01, (float)k.m10);   // ref src/orb_cpu.cpp:178
    if (J.mode != 1) { k.c = orbm::cosf_glibc(angle); k.s = orbm::sinf_glibc(angle); }      // :217-218 (overwrites the moments)
    if (J.mode != 2) B.out_angles[o] = angle;
    if (J.mode == 0) {
      // kp.x *= scale (int * float, truncated): ref src/orb.cpp:94-98
      const float sc = P.lv[l].scale;
      orb_keypoint kp;
      kp.x = __float2int_rz(orbm::fmul((float)x, sc));
      kp.y = __float2int_rz(orbm::fmul((float)y, sc));
      B.out_kps[o] = kp;
      if (B.side_xy) {
        B.side_xy[o] = orb_keypoint{x, y};
        B.side_level[o] = l;
        B.side_resp[o] = B.kept_r[(size_t)f * P.kept_per_frame + P.lv[l].kept_ofs + (widx - s_pref[l])];
      }
    }
  }
  if (J.mode == 1) return;
#if ORB_K3_WARP_LIBM
  __syncwarp();
#else
  __syncthreads();
#endif

  // ---- phase 3: rotated BRIEF, one warp per keypoint, windows W_DEPTH keypoints ahead ----------------
  uint32_t* out_words = (uint32_t*)B.out_desc + ((size_t)f * B.out_cap + k0) * 8;
  for (int i = 0; i < n_mine; i++) {
    const int q = warp + K3_WARPS * i;
    const KpSlot& k = s_kp[q];
    const int kx = k.x, ky = k.y;
    const EdgeTab T{k.tab, k.w, k.h, k.edge_w, (k.h + 3) & ~3};
    cp_async_wait<W_DEPTH - 1>();                 // group i has landed (this thread's copies) ...
    __syncwarp();                                 // ... and everybody else's
    uint32_t word;
    const uint32_t wc_sa = my_sa + (uint32_t)((i % W_DEPTH) * W_SLOT + (19 * W_BW + 19 + ((kx - 19) & 7)) * 2);
    ORB_CHECK(k0 + q < total && total <= B.out_cap && 19 + 19 + ((kx - 19) & 7) < W_BW);
    brief_of(wc_sa, T, kx, ky, k.c, k.s, B.pattern, lane, &word);
    if (lane < 8) out_words[q * 8 + lane] = word;
    __syncwarp();                                 // the slot is free again
    issue_window(i + W_DEPTH);
  }
}

// Harris response for an explicit keypoint list on a level-0 image (stage entry point orb_harris)
__global__ void k_harris_list(const uint8_t* __restrict__ img, int pitch, int w, int h, const orb_keypoint* kps, int n,
                              float k, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ky = kps[i].y, kx = kps[i].x;
  out[i] = harris_at([&](int dy, float (&v)[9]) {
    const uint8_t* q = img + (size_t)reflect101(ky + dy, h) * pitch;
#pragma unroll
    for (int c = 0; c < 9; c++) v[c] = (float)q[reflect101(kx + c - 4, w)];
  }, k);
}

// NMS over a caller's float score map (stage entry point orb_nms_scores; ref d_NMS, src/cuda/NMS.cu:21-128): a pixel at
// least `r` inside the map is kept iff its score exceeds `threshold` and no score in its (2r+1)^2 window is strictly
// greater (ties keep both; outside the map reads 0).  Survivors go to the level-0 candidate list as raster keys; k_select
// (raster policy) then keeps the first nfeatures of them in raster order.
__global__ void k_nms_scores(const float* __restrict__ sc, int pitch, int w, int h, int r, float threshold, unsigned long long* cand,
                             int* count, int cap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < r || y < r || x >= w - r || y >= h - r) return;
  const float v = sc[(size_t)y * pitch + x];
  if (!(v > threshold)) return;
  for (int dy = -r; dy <= r; dy++)
    for (int dx = -r; dx <= r; dx++)
      if (sc[(size_t)(y + dy) * pitch + x + dx] > v) return;
  const int slot = atomicAdd(count, 1);
  if (slot < cap) cand[slot] = (unsigned long long)(unsigned)((y << 16) | x);
}

// ---- the reference's stand-alone filter wrappers (stage entry points orb_conv2d_u8 / orb_gaussian_blur_1d) -----------
// conv2d (ref src/cuda/Convolution.cu:20-103): K x K correlation of the u8 image promoted to float, accumulated row by row
// with one FMA per tap (what `sum += tile * kernel` compiles to in the reference's default nvcc build), converted like
// cv::Mat::convertTo(CV_8U) (round half to even, saturate).  reflect: the image is first extended by K/2 with
// BORDER_REFLECT_101 (GaussianBlurCUDA, SobelCUDA, GaussianBlur), so the output has the input's size; otherwise valid mode.
__device__ __forceinline__ uint8_t float_to_u8(float v) { return (uint8_t)min(max(__float2int_rn(v), 0), 255); }

__global__ void k_conv2d_u8(const uint8_t* __restrict__ img, int pitch, int w, int h, const float* __restrict__ kernel, int K, int reflect,
                            float divisor, uint8_t* __restrict__ out, int opitch, int ow, int oh) {
  extern __shared__ float s_k[];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) s_k[i] = kernel[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= ow || y >= oh) return;
  const int r = K / 2;
  float sum = 0.0f;
  for (int i = 0; i < K; i++) {
    const uint8_t* row = img + (size_t)(reflect ? reflect101(y + i - r, h) : y + i) * pitch;
    for (int j = 0; j < K; j++) sum = __fmaf_rn((float)row[reflect ? reflect101(x + j - r, w) : x + j], s_k[i * K + j], sum);
  }
  if (divisor != 0.0f) sum = __fdiv_rn(sum, divisor);
  out[(size_t)y * opitch + x] = float_to_u8(sum);
}

// GaussianBlur1D (ref src/cuda/GaussianBlur1D.cu): [1 4 6 4 1] / 16 along x then along y in float.  Every intermediate is a
// small dyadic rational, so the float result equals v / 256 exactly with v the integer 5 x 5 sum; convertTo(CV_8U) rounds
// half to even.
__global__ void k_gauss1d_u8(const uint8_t* __restrict__ img, int pitch, int w, int h, uint8_t* __restrict__ out, int opitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int kw[5] = {1, 4, 6, 4, 1};
  int v = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint8_t* row = img + (size_t)reflect101(y + i - 2, h) * pitch;
    int hs = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) hs += kw[j] * row[reflect101(x + j - 2, w)];
    v += kw[i] * hs;
  }
  out[(size_t)y * opitch + x] = float_to_u8(__fdiv_rn((float)v, 256.0f));
}

// trips one check on purpose (tests/test_gpu_bounds.py: the counters of a bounds-check build do count)
__global__ void k_bounds_selftest(int n) { ORB_CHECK((int)threadIdx.x < n); }

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
