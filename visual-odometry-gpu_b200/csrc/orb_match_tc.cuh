// orb_match_tc.cuh -- exact Hamming 2-NN on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces flann->knnMatch(des1, des2, matches, 2) of the reference's VO loops (src/feature_matching.cpp:168,174-182;
// src/feature_tracking.cpp:205-219) by an exact brute-force search, written as a contraction: with descriptor bits mapped to +-1,
//     <a, b> = 256 - 2 * hamming(a, b),   i.e.   hamming = (256 - <a, b>) / 2,
// and an INT8 x INT8 -> INT32 product of +-1 vectors of length 256 is exact (|sum| <= 256).  One CTA owns 128 query
// descriptors of a frame pair and walks over the train descriptors in tiles of 256:
//   k_match_expand : 256-bit descriptors -> 256 int8 (+1 / -1) per descriptor, rows beyond a frame's count zeroed;
//   k_match_tc     : warp 0 (one lane) streams the int8 tiles with TMA (128-byte swizzle) into a two-stage ring,
//                    warp 1 (one lane) issues tcgen05.mma.kind::i8 (M 128 x N 256 x K 32, eight per tile) into one of two
//                    TMEM accumulators (2 x 256 columns), warps 2-5 read the finished accumulator with tcgen05.ld and keep
//                    the two smallest (distance, index) keys per query row while the next tile is being multiplied.
// Ties go to the lower train index (key = distance * 2^14 + index), exactly as the CPU oracle (orc_match_knn2).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_kernels.cuh"

namespace orbk {

constexpr int MT_M = 128, MT_N = 256, MT_KB = 256;          // queries per CTA, train descriptors per tile, bytes per descriptor
constexpr int MT_STAGES = 2;
constexpr int MT_SLAB_A = MT_M * 128, MT_SLAB_B = MT_N * 128;   // one 128-byte K slab of a tile (SWIZZLE_128B atom rows)
constexpr int MT_A_BYTES = 2 * MT_SLAB_A, MT_B_BYTES = 2 * MT_SLAB_B;
constexpr int MT_SMEM = MT_A_BYTES + MT_STAGES * MT_B_BYTES + 128 + 1024 + 1024;   // + barriers + merge area + slack for the 1024-byte alignment
constexpr int MT_EPI_WARPS = 8;                              // two warps per TMEM lane quarter, 128 columns of a tile each
constexpr int MT_THREADS = 64 + 32 * MT_EPI_WARPS;            // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-9: epilogue
constexpr int MT_MAX_INDEX = 1 << 14;                        // train indices must fit the key's low 14 bits

// ---- descriptors -> +-1 int8 rows -----------------------------------------------------------------------------------
// grid (ceil(rows * 16 / 256), sets); thread = one 16-bit group of one descriptor -> 16 bytes.  n == nullptr: all rows valid.
__global__ void k_match_expand(const orb_descriptor* __restrict__ desc, const int* __restrict__ n, int rows_in, long long in_stride,
                               int rows_out, int8_t* __restrict__ out) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = i >> 4, g = i & 15;
  if (r >= rows_out) return;
  uint4 v = make_uint4(0, 0, 0, 0);
  const int nf = n ? min(n[f], rows_in) : rows_in;
  if (r < nf) {
    const uint32_t bits = ((const uint16_t*)(desc + (size_t)f * in_stride + r))[g];
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t s = (((bits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;   // bit b of the nibble -> byte b (0 / 1)
      w[q] = 0xffffffffu ^ (s * 0xfeu);                                               // 1 -> 0x01 (+1), 0 -> 0xff (-1)
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  *(uint4*)(out + ((size_t)f * rows_out + r) * MT_KB + g * 16) = v;
}

// ---- tcgen05 helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// 2-D box (128 bytes x rows) of a 3-D tensor map -> shared memory (128-byte swizzle)
__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* map, uint64_t* bar, int k_byte, int row, int set) {
  tma_load_3d(dst, map, bar, k_byte, row, set);
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor, kind::i8: D = S32, A = B = signed int8, both K-major, N at [17,23) (>> 3), M at [24,29) (>> 4)
constexpr uint32_t MT_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MT_N >> 3) << 17) | ((uint32_t)(MT_M >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(MT_IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i of the warp = TMEM lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 64 consecutive 32-bit columns -> 64 registers per thread
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, int (&v)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// grid (ceil(max queries / 128), pairs).  maps[0]: query rows, maps[1]: train rows (3-D: 256 bytes, rows, pairs; box 128 x 128).
// n_arr != nullptr: pair p matches the n_arr[p] descriptors of frame p against the n_arr[p + 1] of frame p + 1.
__global__ void __launch_bounds__(MT_THREADS, 1) k_match_tc(const CUtensorMap* __restrict__ maps, const int* __restrict__ n_arr, int nq_fixed,
                                                            int nt_fixed, long long out_stride, orb_match* __restrict__ out) {
  extern __shared__ uint8_t mt_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)mt_smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles: 1024-byte aligned
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + MT_A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + MT_A_BYTES + MT_STAGES * MT_B_BYTES);
  uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6, *afull = bars + 8;
  uint32_t* s_tmem = (uint32_t*)(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.y, m0 = blockIdx.x * MT_M;
  const int nq = n_arr ? n_arr[p] : nq_fixed, nt = n_arr ? n_arr[p + 1] : nt_fixed;
  if (m0 >= nq) return;                                   // whole CTA: nothing allocated yet
  const int ntiles = (nt + MT_N - 1) / MT_N;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); mbar_init(tfull + i, 1); mbar_init(tempty + i, MT_EPI_WARPS); }
    mbar_init(afull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {                                        // TMEM: 512 columns = two 128 x 256 int32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (lane == 0) {                                      // ===== TMA producer =====
      mbar_expect_tx(afull, MT_A_BYTES);
      tma_load_rows(s_a, maps + 0, afull, 0, m0, p);
      tma_load_rows(s_a + MT_SLAB_A, maps + 0, afull, 128, m0, p);
      for (int t = 0; t < ntiles; t++) {
        const int st = t & 1;
        mbar_wait(empty + st, ((t >> 1) & 1) ^ 1);        // the MMAs that read this stage have finished
        uint8_t* b = s_b + st * MT_B_BYTES;
        mbar_expect_tx(full + st, MT_B_BYTES);
        tma_load_rows(b, maps + 1, full + st, 0, t * MT_N, p);
        tma_load_rows(b + MT_SLAB_A, maps + 1, full + st, 0, t * MT_N + 128, p);
        tma_load_rows(b + MT_SLAB_B, maps + 1, full + st, 128, t * MT_N, p);
        tma_load_rows(b + MT_SLAB_B + MT_SLAB_A, maps + 1, full + st, 128, t * MT_N + 128, p);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                      // ===== MMA issuer =====
      mbar_wait(afull, 0);
      for (int t = 0; t < ntiles; t++) {
        const int st = t & 1;
        mbar_wait(tempty + st, ((t >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
        mbar_wait(full + st, (t >> 1) & 1);               // the tile has landed
        tc_fence_after();
        const uint32_t a0 = smem_u32(s_a), b0 = smem_u32(s_b + st * MT_B_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {                  // K = 256 bytes = 8 x 32; 4 steps per 128-byte slab
          const uint32_t ao = a0 + (ks >> 2) * MT_SLAB_A + (ks & 3) * 32, bo = b0 + (ks >> 2) * MT_SLAB_B + (ks & 3) * 32;
          umma_i8(tmem + st * MT_N, umma_desc_k_sw128(ao), umma_desc_k_sw128(bo), ks > 0);
        }
        umma_commit(empty + st);                          // smem stage free once these MMAs are done
        umma_commit(tfull + st);                          // accumulator ready
      }
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32 * (w % 4) .. + 31 (its query rows) and columns 128 * half .. + 127 of every
    // tile; chunks of 32 columns, the next chunk's tcgen05.ld in flight while the current one is folded into the keys =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    int k1 = 0x7fffffff, k2 = 0x7fffffff;                 // two smallest keys: distance << 14 | train index
    auto fold = [&](const int (&v)[64], int kb, int nvalid) {   // key = (256 - dot) / 2 * 2^14 + j = (256 - dot) * 2^13 + j
      if (nvalid >= 64) {
#pragma unroll
        for (int i = 0; i < 64; i++) {
          const int key = kb + i - v[i] * 8192;
          k2 = min(k2, max(k1, key));
          k1 = min(k1, key);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; i++) {
          const int key = i < nvalid ? kb + i - v[i] * 8192 : 0x7fffffff;
          k2 = min(k2, max(k1, key));
          k1 = min(k1, key);
        }
      }
    };
    for (int t = 0; t < ntiles; t++) {
      const int st = t & 1;
      mbar_wait(tfull + st, (t >> 1) & 1);
      tc_fence_after();
      const int c0 = half * 128;                          // first column of this warp inside the tile
      const int jn = min(MT_N, nt - t * MT_N) - c0;       // valid train columns from c0 on (<= 0: nothing for this warp)
      const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + st * MT_N + c0;
      const int kb = (256 << 13) + t * MT_N + c0;
      int va[64], vb[64];
      if (jn > 0) {
        tmem_ld_32x64(ta, va);
        tmem_ld_wait();
        if (jn > 64) tmem_ld_32x64(ta + 64, vb);          // in flight while the first 64 columns are folded
        fold(va, kb, jn);
        if (jn > 64) {
          tmem_ld_wait();
          fold(vb, kb + 64, jn - 64);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + st);
    }
    // the two column halves of a row meet in shared memory
    int* s_merge = (int*)(bars + 16);
    const int r = quarter * 32 + lane;
    if (half == 1) { s_merge[2 * r] = k1; s_merge[2 * r + 1] = k2; }
    asm volatile("bar.sync 1, %0;" ::"r"(32 * MT_EPI_WARPS) : "memory");
    if (half == 0 && m0 + r < nq) {
      const int b1 = s_merge[2 * r], b2 = s_merge[2 * r + 1];
      const int m1 = min(k1, b1), m2 = min(max(k1, b1), min(k2, b2));
      orb_match m;
      m.idx1 = m1 == 0x7fffffff ? -1 : (m1 & (MT_MAX_INDEX - 1)); m.dist1 = m1 == 0x7fffffff ? 0x7fffffff : (m1 >> 14);
      m.idx2 = m2 == 0x7fffffff ? -1 : (m2 & (MT_MAX_INDEX - 1)); m.dist2 = m2 == 0x7fffffff ? 0x7fffffff : (m2 >> 14);
      out[(size_t)p * out_stride + m0 + r] = m;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace orbk
