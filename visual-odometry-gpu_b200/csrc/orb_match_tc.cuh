// orb_match_tc.cuh -- exact Hamming 2-NN on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces flann->knnMatch(des1, des2, matches, 2) of the reference's VO loops (src/feature_matching.cpp:168,174-182;
// src/feature_tracking.cpp:205-219) by an exact brute-force search, written as a contraction: with descriptor bits mapped to +-1,
//     <a, b> = 256 - 2 * hamming(a, b),   i.e.   hamming = (256 - <a, b>) / 2,
// and an INT8 x INT8 -> INT32 product of such vectors of length 256 is exact.
//   k_match_expand : 256-bit descriptors -> 256 int8 (+8 / -8) per descriptor, rows beyond a frame's count zeroed;
//   k_match_tc     : persistent, one CTA per SM walking over work items (frame pair, block of 256 query descriptors).
//                    The item's two 128-row query blocks stay in shared memory while the train descriptors stream past in
//                    tiles of 128 through a TMA ring (128-byte swizzle), so each train tile read from L2 serves 256 queries.
//                    warp 0 is the TMA producer, warp 1 issues tcgen05.mma.kind::i8 (M 128 x N 128 x K 32; eight per tile
//                    and query block, plus a ninth that adds the column index, see the kernel) into four TMEM accumulators
//                    (2 query blocks x 2 buffers x 128 columns) -- both warps run converged, one elected lane executes the
//                    asynchronous instruction; warps 2-9 (one per TMEM lane quarter and query block) read finished
//                    accumulators with tcgen05.ld.pack::16b and keep the two best per query row with packed 16-bit min / max
//                    while the next tile is being multiplied.  Pipeline state (barrier phases) carries over from item to
//                    item, so the prologue of an item (query block load) overlaps the tail of the previous one.
// Ties go to the lower train index, exactly as the CPU oracle (orc_match_knn2).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_kernels.cuh"

#ifndef ORB_MT_PROBE
#define ORB_MT_PROBE 0   // timing probes (results wrong): 1 no accumulator reads, 2 one MMA per tile, 4 train tiles loaded once, 8 issuer does not wait for the epilogue
#endif

namespace orbk {

constexpr int MT_M = 128, MT_N = 128, MT_KB = 256;          // rows of a query block, train descriptors per tile, bytes per descriptor
constexpr int MT_MB = 2;                                    // query blocks per work item
constexpr int MT_STAGES = 4;
constexpr int MT_SLAB_A = MT_M * 128, MT_SLAB_B = MT_N * 128;   // one 128-byte K slab of a tile (SWIZZLE_128B atom rows)
constexpr int MT_A_BYTES = 2 * MT_SLAB_A, MT_B_BYTES = 2 * MT_SLAB_B;
constexpr int MT_SMEM = MT_MB * MT_A_BYTES + MT_SLAB_A + MT_SLAB_B + MT_STAGES * MT_B_BYTES + 256 + 1024;   // + index slabs + barriers + alignment slack
constexpr int MT_EPI_WARPS = 4 * MT_MB;                      // one warp per TMEM lane quarter and query block
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
  const int nf = n ? max(0, min(n[f], rows_in)) : rows_in;
  if (r < nf) {
    const uint32_t bits = ((const uint16_t*)(desc + (size_t)f * in_stride + r))[g];
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t s = (((bits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;   // bit b of the nibble -> byte b (0 / 1)
      w[q] = 0xf8f8f8f8u ^ (s * 0xf0u);                                               // 1 -> 0x08 (+8), 0 -> 0xf8 (-8)
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
// 2-D box (128 bytes x rows) of a 3-D tensor map -> shared memory (128-byte swizzle), issued by one elected lane
__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* map, uint64_t* bar, int k_byte, int row, int set) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(k_byte), "r"(row), "r"(set)
               : "memory");
}
// the same box, only pulled into L2 (no destination): takes the HBM latency off the next item's query load
__device__ __forceinline__ void tma_prefetch_rows(const CUtensorMap* map, int k_byte, int row, int set) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];\n\t}\n" ::"l"(map), "r"(k_byte), "r"(row), "r"(set)
               : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::i8: D = S32 (2) at [4,6), A = B = signed int8 (1) at [7,10) and
// [10,13), both K-major, N >> 3 at [17,23), M >> 4 at [24,29).  (kind::f8f6f4 with E4M3 operands and FP32 accumulators is
// exact here too and was measured: no faster, and its accumulators do not fit the 16-bit packed read of the epilogue.)
constexpr uint32_t MT_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MT_N >> 3) << 17) | ((uint32_t)(MT_M >> 4) << 24);
#define ORB_MT_KIND "kind::i8"
// The producer and issuer warps run their loops CONVERGED and one elected lane executes the asynchronous instruction inside
// the asm: addresses, descriptors and coordinates are then warp-uniform for the compiler and stay in uniform registers (issued
// from inside `if (lane == 0)` every UTCxMMA / UTMALDG is wrapped in a R2UR.BROADCAST loop of ~18 instructions, and that
// single-thread instruction stream, not the tensor pipe, set the pace: ~72 clk per MMA).
__device__ __forceinline__ void umma_8bit(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1." ORB_MT_KIND " [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(MT_IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* b, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(
                   smem_u32(b)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_elect(uint64_t* b) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(smem_u32(b))
               : "memory");
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
// 32 lanes x 128 consecutive columns, low 16 bits of each, two columns per register -> 64 registers per thread
__device__ __forceinline__ void tmem_ld_32x128_pack16(uint32_t taddr, int (&v)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the loaded registers are valid only after tcgen05.wait::ld: pin every use behind the wait (no instructions emitted)
__device__ __forceinline__ void tmem_ld_wait(int (&v)[64]) {
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 64; i++) asm volatile("" : "+r"(v[i]));
}

// grid: persistent (<= SM count).  map_q: query rows, map_t: train rows (3-D: 256 bytes, rows, pairs; box 128 x 128), passed as
// kernel parameters (no device copy of the descriptors, no copy to wait for).
// n_arr != nullptr: pair p matches the n_arr[p] descriptors of frame p against the n_arr[p + 1] of frame p + 1.
// Work item i = (pair i / qblocks, query rows 256 * (i % qblocks) ..).
//
// The accumulator already IS the comparison key: the expanded rows hold +-8, so the eight data MMAs leave 64 * <a, b> (a
// multiple of 128), and a ninth MMA over a constant K slab (query side: -1 in its first byte, train side: column - 64) adds
// 64 - column:   acc = 64 * dot + 64 - column,   larger = nearer, ties to the lower column,
// which the epilogue folds with three integer min / max per distance; tiles are merged in order into global keys
// distance * 2^14 + index (smaller = nearer), so an equal distance in a later tile never displaces an earlier one.
__global__ void __launch_bounds__(MT_THREADS, 1) k_match_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
                                                            const int* __restrict__ n_arr, int nq_fixed,
                                                            int nt_fixed, int qblocks, int n_items, long long out_stride,
                                                            orb_match* __restrict__ out) {
  extern __shared__ uint8_t mt_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)mt_smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles: 1024-byte aligned
  uint8_t* s_a = smem;                                    // [MT_MB] query blocks
  uint8_t* s_ax = s_a + MT_MB * MT_A_BYTES;               // constant K slabs of the index term
  uint8_t* s_bx = s_ax + MT_SLAB_A;
  uint8_t* s_b = s_bx + MT_SLAB_B;                        // [MT_STAGES] train tiles
  uint64_t* bars = (uint64_t*)(s_b + MT_STAGES * MT_B_BYTES);
  uint64_t *bfull = bars, *bempty = bars + MT_STAGES, *afull = bars + 2 * MT_STAGES, *aempty = afull + MT_MB;
  uint64_t *tfull = aempty + MT_MB, *tempty = tfull + 2 * MT_MB;   // [query block][buffer]
  uint32_t* s_tmem = (uint32_t*)(tempty + 2 * MT_MB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MT_STAGES; i++) { mbar_init(bfull + i, 1); mbar_init(bempty + i, 1); }
    for (int i = 0; i < MT_MB; i++) { mbar_init(afull + i, 1); mbar_init(aempty + i, 1); }
    for (int i = 0; i < 2 * MT_MB; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
    mbar_fence_init();
  }
  if (warp == 1) {                                        // TMEM: 512 columns = four 128 x 128 int32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // index slabs, in the layout TMA's 128-byte swizzle gives the data slabs: 16-byte chunk c of row r sits at chunk c ^ (r & 7);
  // only the first K step (32 bytes = chunks 0, 1) of a row is ever multiplied
  for (int i = threadIdx.x; i < (MT_M + MT_N) * 8; i += MT_THREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c == 0) v.x = r < MT_M ? 0xffu : (uint32_t)((r - MT_M - 64) & 0xff);   // queries: -1, train column j: j - 64
    const int rr = r < MT_M ? r : r - MT_M;
    *(uint4*)((r < MT_M ? s_ax : s_bx) + rr * 128 + ((c ^ (rr & 7)) << 4)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // every role walks the same item list and derives the same (nq, nt, blocks, tiles) for an item
  auto item_shape = [&](int item, int& p, int& m0, int& nq, int& nt, int& mbs, int& ntiles) {
    p = item / qblocks;
    m0 = (item - p * qblocks) * (MT_MB * MT_M);
    nq = n_arr ? max(0, min(n_arr[p], nq_fixed)) : nq_fixed;        // counts beyond the capacity of a set are clipped, as in k_match_expand
    nt = n_arr ? max(0, min(n_arr[p + 1], nt_fixed)) : nt_fixed;
    ntiles = (nt + MT_N - 1) / MT_N;
    mbs = nq <= m0 ? 0 : min(MT_MB, (nq - m0 + MT_M - 1) / MT_M);
    if (ntiles == 0) mbs = -mbs;                          // nothing to multiply: the epilogue still writes "no neighbour"
  };

  if (warp == 0) {
    {                                                     // ===== TMA producer (whole warp, converged) =====
      uint32_t b_it = 0, a_it = 0;                        // tiles / items loaded so far (ring position and phase)
      auto load_b = [&](int p, int t) {
        const int st = b_it % MT_STAGES;
        mbar_wait(bempty + st, ((b_it / MT_STAGES) & 1) ^ 1);   // the MMAs that read this stage have finished
        uint8_t* b = s_b + st * MT_B_BYTES;
#if ORB_MT_PROBE & 4
        if (b_it >= MT_STAGES) { mbar_arrive_elect(bfull + st); b_it++; return; }
#endif
        mbar_expect_tx_elect(bfull + st, MT_B_BYTES);
        tma_load_rows(b, &map_t, bfull + st, 0, t * MT_N, p);
        tma_load_rows(b + MT_SLAB_B, &map_t, bfull + st, 128, t * MT_N, p);
        b_it++;
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int p, m0, nq, nt, mbs, ntiles;
        item_shape(item, p, m0, nq, nt, mbs, ntiles);
        if (mbs <= 0) continue;
        const int pre = min(ntiles, MT_STAGES - 1);       // train tiles requested before the query blocks: their stages free up first
        for (int t = 0; t < pre; t++) load_b(p, t);
        for (int mb = 0; mb < mbs; mb++) {
          mbar_wait(aempty + mb, (a_it & 1) ^ 1);         // the previous item's MMAs on this block have finished
          uint8_t* a = s_a + mb * MT_A_BYTES;
          mbar_expect_tx_elect(afull + mb, MT_A_BYTES);
          tma_load_rows(a, &map_q, afull + mb, 0, m0 + mb * MT_M, p);
          tma_load_rows(a + MT_SLAB_A, &map_q, afull + mb, 128, m0 + mb * MT_M, p);
        }
        for (int mb = mbs; mb < MT_MB; mb++) {            // unused block: keep its barriers in step
          mbar_wait(aempty + mb, (a_it & 1) ^ 1);
          mbar_arrive_elect(afull + mb);
        }
        a_it++;
        if (item + (int)gridDim.x < n_items) {            // next item's query rows -> L2 while this item is being multiplied
          int p2, m2, nq2, nt2, mbs2, ntiles2;
          item_shape(item + gridDim.x, p2, m2, nq2, nt2, mbs2, ntiles2);
          for (int mb = 0; mb < mbs2; mb++) {
            tma_prefetch_rows(&map_q, 0, m2 + mb * MT_M, p2);
            tma_prefetch_rows(&map_q, 128, m2 + mb * MT_M, p2);
          }
        }
        for (int t = pre; t < ntiles; t++) load_b(p, t);
      }
    }
  } else if (warp == 1) {
    {                                                     // ===== MMA issuer (whole warp, converged) =====
      const uint64_t ax = umma_desc_k_sw128(smem_u32(s_ax)), bx = umma_desc_k_sw128(smem_u32(s_bx));
      const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(s_a)), b_desc0 = umma_desc_k_sw128(smem_u32(s_b));
      uint32_t b_it = 0, a_it = 0, acc_cnt[MT_MB] = {};   // tiles landed / items started / tiles multiplied per query block
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int p, m0, nq, nt, mbs, ntiles;
        item_shape(item, p, m0, nq, nt, mbs, ntiles);
        if (mbs <= 0) continue;
        for (int t = 0; t < ntiles; t++, b_it++) {
          const int st = b_it % MT_STAGES;
          mbar_wait(bfull + st, (b_it / MT_STAGES) & 1);  // the tile has landed
          const uint64_t b_desc = b_desc0 + (uint64_t)((st * MT_B_BYTES) >> 4);   // the address field counts 16-byte units
          for (int mb = 0; mb < MT_MB; mb++) {
            if (t == 0) mbar_wait(afull + mb, a_it & 1);
            if (mb < mbs) {
              const int buf = acc_cnt[mb] & 1;
#if !(ORB_MT_PROBE & 8)
              mbar_wait(tempty + 2 * mb + buf, ((acc_cnt[mb] >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
#endif
              tc_fence_after();
              const uint64_t a_desc = a_desc0 + (uint64_t)((mb * MT_A_BYTES) >> 4);
              const uint32_t d = tmem + (2 * mb + buf) * MT_N;
#pragma unroll
              for (int ks = 0; ks < ((ORB_MT_PROBE & 2) ? 1 : 8); ks++) {   // K = 256 bytes = 8 x 32; 4 steps per 128-byte slab
                umma_8bit(d, a_desc + (uint64_t)(((ks >> 2) * MT_SLAB_A + (ks & 3) * 32) >> 4),
                          b_desc + (uint64_t)(((ks >> 2) * MT_SLAB_B + (ks & 3) * 32) >> 4), ks > 0);
              }
              umma_8bit(d, ax, bx, 1);                      // + 64 - column
              umma_commit(tfull + 2 * mb + buf);          // accumulator ready
              acc_cnt[mb]++;
            }
            if (t == ntiles - 1) umma_commit(aempty + mb);   // the query block may be replaced once everything issued so far is done
          }
          umma_commit(bempty + st);                       // smem stage free once these MMAs are done
        }
        a_it++;
      }
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32 * (w % 4) .. + 31 = query rows of block mb = (w - 2) / 4 and all 128 columns
    // of every tile.  An accumulator fits 16 bits (|64 * dot + 64 - j| <= 16448), so ONE tcgen05.ld.pack::16b brings a whole
    // tile row as 64 registers of two columns each, and the two best are kept per 16-bit half with packed min / max (1.5
    // instructions per distance); as every value carries its own column, the halves are simply compared at the end of the
    // tile.  The next tile's load is in flight while the current one is folded. =====
    const int quarter = warp & 3, mb = (warp - 2) >> 2;
    const uint32_t NONE2 = 0x80008000u;                   // -32768 | -32768: below any accumulator
    uint32_t acc_it = 0;                                  // tiles of this warp's query block folded so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int p, m0, nq, nt, mbs, ntiles;
      item_shape(item, p, m0, nq, nt, mbs, ntiles);
      if (mbs == 0) continue;
      const int row = m0 + mb * MT_M + quarter * 32 + lane;
      if (mbs < 0) {                                      // no train descriptors
        if (row < nq) out[(size_t)p * out_stride + row] = orb_match{-1, 0x7fffffff, -1, 0x7fffffff};
        continue;
      }
      if (mb >= mbs) continue;
      int g1 = 0x7fffffff, g2 = 0x7fffffff;               // two smallest global keys: distance << 14 | train index
      // acc = 64 * dot + 64 - j  ->  key = (256 - dot) / 2 * 2^14 + base + j = 2^21 + (w - j) * 128 + base + j,  w = 64 - acc
      auto to_key = [&](int acc, int base) {
        const int w = 64 - acc, j = w & 127;
        return acc == -32768 ? 0x7fffffff : w * 128 + ((1 << 21) + base) - j * 127;
      };
      auto fold_tile = [&](const int (&v)[64], int nvalid, int base) {
        uint32_t a1 = NONE2, a2 = NONE2, b1 = NONE2, b2 = NONE2;   // two largest per 16-bit half, even / odd registers
        if (nvalid >= MT_N) {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            a2 = __vmaxs2(a2, __vmins2(a1, (uint32_t)v[i]));
            a1 = __vmaxs2(a1, (uint32_t)v[i]);
            b2 = __vmaxs2(b2, __vmins2(b1, (uint32_t)v[i + 1]));
            b1 = __vmaxs2(b1, (uint32_t)v[i + 1]);
          }
        } else {                                          // last tile of the pair: register i = columns 2 i (low half), 2 i + 1
#pragma unroll
          for (int i = 0; i < 64; i++) {
            const uint32_t x = 2 * i + 1 < nvalid ? (uint32_t)v[i] : 2 * i < nvalid ? (((uint32_t)v[i] & 0xffffu) | 0x80000000u) : NONE2;
            a2 = __vmaxs2(a2, __vmins2(a1, x));
            a1 = __vmaxs2(a1, x);
          }
        }
        const uint32_t t1 = __vmaxs2(a1, b1), t2 = __vmaxs2(__vmins2(a1, b1), __vmaxs2(a2, b2));
        const int lo1 = (int)(short)(t1 & 0xffffu), hi1 = (int)t1 >> 16, lo2 = (int)(short)(t2 & 0xffffu), hi2 = (int)t2 >> 16;
        const int f1 = to_key(max(lo1, hi1), base), f2 = to_key(max(min(lo1, hi1), max(lo2, hi2)), base);
        g2 = min(max(g1, f1), min(g2, f2));
        g1 = min(g1, f1);
      };
      const uint32_t ta0 = tmem + ((uint32_t)(quarter * 32) << 16) + 2 * mb * MT_N;
      int va[64], vb[64];
      // one step: the current tile's registers land, its accumulator is handed back, the next tile's load is started, and
      // only then the current registers are folded
      auto step = [&](int t, int (&cur)[64], int (&nxt)[64]) {
        const int buf = acc_it & 1;
        tmem_ld_wait(cur);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty + 2 * mb + buf);
        if (t + 1 < ntiles) {
          mbar_wait(tfull + 2 * mb + (buf ^ 1), ((acc_it + 1) >> 1) & 1);
          tc_fence_after();
#if !(ORB_MT_PROBE & 1)
          tmem_ld_32x128_pack16(ta0 + (buf ^ 1) * MT_N, nxt);
#endif
        }
#if !(ORB_MT_PROBE & 1)
        fold_tile(cur, min(MT_N, nt - t * MT_N), t * MT_N);
#endif
        acc_it++;
      };
      mbar_wait(tfull + 2 * mb + (acc_it & 1), (acc_it >> 1) & 1);
      tc_fence_after();
#if !(ORB_MT_PROBE & 1)
      tmem_ld_32x128_pack16(ta0 + (acc_it & 1) * MT_N, va);
#endif
      for (int t = 0; t < ntiles; t += 2) {
        step(t, va, vb);
        if (t + 1 < ntiles) step(t + 1, vb, va);
      }
      if (row < nq) {
        orb_match m;
        m.idx1 = g1 == 0x7fffffff ? -1 : (g1 & (MT_MAX_INDEX - 1)); m.dist1 = g1 == 0x7fffffff ? 0x7fffffff : (g1 >> 14);
        m.idx2 = g2 == 0x7fffffff ? -1 : (g2 & (MT_MAX_INDEX - 1)); m.dist2 = g2 == 0x7fffffff ? 0x7fffffff : (g2 >> 14);
        out[(size_t)p * out_stride + row] = m;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace orbk
