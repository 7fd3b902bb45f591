// orb_ingest_kernels.cuh -- device side of the frame ingest: inflate (RFC 1951) and PNG unfilter on the
// GPU, so that a frame crosses PCIe compressed and the host only reads files and checks chunk CRCs.  Replaces what
// cv::imread(path, IMREAD_GRAYSCALE) does inside libpng/zlib (reference src/feature_matching.cpp:55,59).
//
//   k_inflate  : one warp per deflate stream (= one frame).  Literal runs are decoded by all 32 lanes (self-synchronising
//                segments, see inf_decode_segment); lengths / distances / block headers are walked by lane 0 with a
//                shared-memory lookup table that yields up to two literals per probe; the other lanes do everything else
//                that is not serial -- coalesced 16-byte refills of the input ring, 16-byte flushes of the output window,
//                and the per-block table construction (code assignment by match_any ranks, replicated fills, the
//                two-literal augmentation pass).
//   k_unfilter : one warp per frame, lane = scanline, lanes skewed by one pixel so that left / up / up-left of the PNG
//                predictors (Sub, Up, Average, Paeth) are a register, a shuffle from the lane above, and the previous
//                shuffle.  Writes level 0 in the staging layout the ORB kernels read.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orbk {

struct InflateDesc {
  const uint8_t* in;    // raw deflate stream, 16-byte aligned, readable (zero padded) up to the next multiple of 512 bytes + 16
  uint32_t in_bytes;
  uint32_t out_bytes;   // exact size the stream must inflate to
  uint8_t* out;         // 16-byte aligned
};

enum : int { INF_OK = 0, INF_CORRUPT = 1, INF_SIZE = 2, INF_TRUNCATED = 3, INF_TABLE = 4, INF_FILTER = 5, INF_CHECKSUM = 6, INF_CRC = 7 };

constexpr int INF_LIT_ROOT = 11, INF_DIST_ROOT = 8;
constexpr int INF_LIT_SUB = 1024, INF_DIST_SUB = 512;
#ifndef ORB_INF_RING
#define ORB_INF_RING 8192
#endif
constexpr int INF_RING = ORB_INF_RING; // output window kept in shared memory (bytes, power of two)
constexpr int INF_IN_WORDS = 512;     // input ring (32-bit words, power of two)
constexpr int INF_MARGIN = 600;       // a decode step never adds more than 258 + 2 bytes; flush well before the ring wraps
// match sources nearer than this are read from the window (valid up to INF_RING - 2 back), the others from global memory
// (everything older than INF_RING - INF_MARGIN + 258 has been flushed when a step starts)
constexpr int INF_NEAR = INF_RING - 300;

// Literal/length table entries (u32).  The literal run reads only [3:0] and [5:4] and must find zeros there for anything
// that is not a literal, so that such a probe consumes nothing and emits nothing:
//   literal(s) : [3:0] bits to consume  [5:4] number of literals (1, 2)  [15:12] bits of the first literal
//                [23:16] first literal  [31:24] second literal
//   others     : [7] = 1  [11:8] bits to consume  [14:12] kind  and in [31:16]
//                length: [27:16] base, [31:28] extra bits      link: [27:16] subtable offset, [31:28] its index bits
// Distance table entries: [3:0] bits to consume  [6:4] kind  [11:8] extra bits (links: index bits)  [31:16] base / offset
enum : uint32_t { IK_BAD = 0, IK_LEN = 3, IK_EOB = 4, IK_LINK = 5, IK_DIST = 6 };

struct InflateShared {
  uint32_t lit[(1 << INF_LIT_ROOT) + INF_LIT_SUB];
  uint32_t dist[(1 << INF_DIST_ROOT) + INF_DIST_SUB];
  __align__(16) uint8_t ring[INF_RING];
  __align__(16) uint32_t in[INF_IN_WORDS];
  uint8_t lens[320];
  uint16_t codes[320];
  uint32_t count[16], next[16];
  int build_status;
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ uint32_t inf_symbol_entry(int sym, bool is_dist, int bits) {
  if (is_dist) return sym < 30 ? ((uint32_t)c_dist_base[sym] << 16) | ((uint32_t)c_dist_extra[sym] << 8) | (IK_DIST << 4) | bits : 0u;
  if (sym < 256) return ((uint32_t)sym << 16) | ((uint32_t)bits << 12) | (1u << 4) | bits;
  if (sym == 256) return (IK_EOB << 12) | ((uint32_t)bits << 8) | 0x80u;
  return sym < 286 ? ((uint32_t)c_len_extra[sym - 257] << 28) | ((uint32_t)c_len_base[sym - 257] << 16) | (IK_LEN << 12) | ((uint32_t)bits << 8) | 0x80u
                   : 0x80u;
}
// link entries: the index bits sit above the offset in both formats, so atomicMax over unplaced links keeps the widest
__device__ __forceinline__ uint32_t inf_link_entry(bool is_dist, int sub_bits, int root) {
  return is_dist ? ((uint32_t)sub_bits << 8) | (IK_LINK << 4) | root : ((uint32_t)sub_bits << 28) | (IK_LINK << 12) | ((uint32_t)root << 8) | 0x80u;
}
__device__ __forceinline__ int inf_link_offset(bool is_dist, uint32_t e) { return is_dist ? e >> 16 : (e >> 16) & 0xfff; }
__device__ __forceinline__ int inf_link_bits(bool is_dist, uint32_t e) { return is_dist ? (e >> 8) & 15 : e >> 28; }

// Canonical code lengths (S.lens[first .. first + n)) -> two-level table.  Whole warp; returns 0 or an INF_ status.
__device__ __forceinline__ int inf_build_table(InflateShared& S, int first, int n, bool is_dist, uint32_t* table, int root, int sub_cap, int lane) {
  const int nroot = 1 << root;
  if (lane < 16) S.count[lane] = 0;
  if (is_dist) for (int i = lane; i < nroot; i += 32) table[i] = 0u;   // a distance code may be incomplete (one code only)
  __syncwarp();
  // histogram of the lengths: one leader per distinct length in each group of 32 symbols
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    const int l = s < n ? S.lens[first + s] : 0;
    const unsigned same = __match_any_sync(0xffffffffu, l);
    if (l && (same & ((1u << lane) - 1)) == 0) S.count[l] += __popc(same);
    __syncwarp();
  }
  if (lane == 0) {
    int left = 1, used = 0;
    uint32_t code = 0;
    S.count[0] = 0;
    int st = 0;
    for (int l = 1; l <= 15; l++) {
      code = (code + S.count[l - 1]) << 1;
      S.next[l] = code;
      left = (left << 1) - (int)S.count[l];
      if (left < 0) st = INF_CORRUPT;                      // over-subscribed
      used += S.count[l];
    }
    if (left > 0 && !(is_dist && used <= 1)) st = INF_CORRUPT;   // incomplete (zlib accepts only the one-code distance tree)
    S.build_status = st;
  }
  __syncwarp();
  if (S.build_status) return S.build_status;
  // canonical codes in symbol order: rank among the equal lengths of the group + running counter per length
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    const int l = s < n ? S.lens[first + s] : 0;
    const unsigned same = __match_any_sync(0xffffffffu, l);
    if (l) {
      const uint32_t code = S.next[l] + __popc(same & ((1u << lane) - 1));
      S.codes[s] = (uint16_t)(__brev(code) >> (32 - l));      // the bit reader is LSB first
    }
    __syncwarp();
    if (l && (same & ((1u << lane) - 1)) == 0) S.next[l] += __popc(same);
    __syncwarp();
  }
  // codes that fit the root table: every replica of every symbol.  Short codes (>= 32 replicas) are spread over the
  // warp, the others go one symbol per lane.
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    const int l = s < n ? S.lens[first + s] : 0;
    unsigned wide = __ballot_sync(0xffffffffu, l && l + 5 <= root);
    while (wide) {
      const int src = __ffs(wide) - 1;
      wide &= wide - 1;
      const int ls = __shfl_sync(0xffffffffu, l, src);
      const uint32_t e = inf_symbol_entry(s0 + src, is_dist, ls);
      const int c = S.codes[s0 + src];
      for (int i = c + (lane << ls); i < nroot; i += 32 << ls) table[i] = e;
    }
    if (l && l <= root && l + 5 > root) {
      const uint32_t e = inf_symbol_entry(s, is_dist, l);
      for (int i = S.codes[s]; i < nroot; i += 1 << l) table[i] = e;
    }
  }
  __syncwarp();
  // longer codes: a link per root prefix; the subtable is indexed by the longest code below it
  bool any_long = false;
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    if (s < n && S.lens[first + s] > root) table[S.codes[s] & (nroot - 1)] = 0u;   // entry of the previous block's table
  }
  __syncwarp();
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    const int l = s < n ? S.lens[first + s] : 0;
    if (l > root) {
      atomicMax(&table[S.codes[s] & (nroot - 1)], inf_link_entry(is_dist, l - root, root));   // offset 0 = not placed yet
      any_long = true;
    }
  }
  any_long = __any_sync(0xffffffffu, any_long);
  if (any_long) {
    __syncwarp();
    if (lane == 0) {
      int top = nroot, st = 0;
      for (int s = 0; s < n; s++) {
        const int l = S.lens[first + s];
        if (l <= root) continue;
        const int pre = S.codes[s] & (nroot - 1);
        const uint32_t e = table[pre];
        if (inf_link_offset(is_dist, e) == 0) {
          const int size = 1 << inf_link_bits(is_dist, e);
          if (top + size > nroot + sub_cap) { st = INF_TABLE; break; }
          table[pre] = e | ((uint32_t)top << 16);
          top += size;
        }
      }
      S.build_status = st;
    }
    __syncwarp();
    if (S.build_status) return S.build_status;
    for (int s0 = 0; s0 < n; s0 += 32) {
      const int s = s0 + lane;
      const int l = s < n ? S.lens[first + s] : 0;
      if (l > root) {
        const int c = S.codes[s];
        const uint32_t link = table[c & (nroot - 1)];
        const int base = inf_link_offset(is_dist, link), size = 1 << inf_link_bits(is_dist, link);
        const uint32_t e = inf_symbol_entry(s, is_dist, l - root);
        for (int i = c >> root; i < size; i += 1 << (l - root)) table[base + i] = e;
      }
    }
  }
  __syncwarp();
  if (!is_dist) {
    // second literal: when the bits left in the root index decode another complete literal, take both in one probe
    for (int i = lane; i < nroot; i += 32) {
      const uint32_t e1 = table[i];
      if ((e1 & 0xb0u) != 0x10u) continue;              // not a single literal
      const int l1 = (e1 >> 12) & 15;
      const uint32_t e2 = table[i >> l1];
      const int l2 = (e2 >> 12) & 15;
      if (!(e2 & 0x80u) && l1 + l2 <= root)             // a literal entry (its first literal is what follows here)
        table[i] = (e2 << 8 & 0xff000000u) | (e1 & 0x00fff000u) | (2u << 4) | (l1 + l2);
    }
    __syncwarp();
  }
  return 0;
}

// ---- literal runs with all 32 lanes -----------------------------------------------------------
// A Huffman stream is serial only until decoders that start at wrong bit positions fall into step with the true symbol
// grid, which prefix codes do after a few symbols.  A round cuts the next 32 x INF_SEG bits into 32 segments; lane i decodes
// the symbols that START inside segment i.  Pass 1 starts every lane at its segment boundary (a guess); after that lane i
// restarts from the position where lane i-1 ended until no start changes any more -- lane 0 starts at the true position, so
// the fixed point is the true decode (two or three passes in practice).  Counts are prefix-summed and a last pass writes
// the bytes.  A lane stops at the first entry that is not a literal; the serial code takes over at the first such position.
#ifndef ORB_INF_SEG
#define ORB_INF_SEG 256
#endif
constexpr int INF_SEG = ORB_INF_SEG;    // bits per lane and round
constexpr int INF_SEG_CAP = 220;        // bytes per lane and round: 32 lanes fit the free part of the window whatever the code lengths
constexpr int INF_PAR_MIN = 192;        // a round that yields fewer bytes (match-heavy data) pauses the parallel rounds
constexpr int INF_PAR_PAUSE = 8;

#ifndef ORB_INF_VOTE
#define ORB_INF_VOTE 4
#endif
constexpr int INF_STEPS_PER_VOTE = ORB_INF_VOTE;   // probes between two checks of "is any lane still inside its segment"
#ifndef ORB_INF_COUNT_PAIRS
#define ORB_INF_COUNT_PAIRS 1
#endif
// All 32 lanes call this together (lanes with run = false only keep the loop company): the loop is uniform and the body
// is predicated rather than branched, because a divergent body costs every lane the sum of all paths.
// Every pass steps by table probes (one or two literals; a second literal that starts in the next segment is left to the
// next lane).  Stepping the counting passes one symbol at a time would let two decoders fall into step on ANY common symbol
// boundary instead of only on common probe boundaries, but measured slower (10.7 vs 8.2 ms per KITTI frame): the extra
// probes cost more than the extra fix-up passes.
template <bool WRITE>
__device__ __forceinline__ void inf_decode_segment(InflateShared& S, uint32_t start, uint32_t seg_end, uint32_t out_at, bool run,
                                                   uint32_t* end, uint32_t* cnt, bool* stop) {
  constexpr bool PAIRS = WRITE || ORB_INF_COUNT_PAIRS;
  uint32_t pos = start, count = 0;
  bool st = false;
  bool active = run && pos < seg_end;
  uint32_t wi = pos >> 5;
  uint64_t bb = (uint64_t)(S.in[wi & (INF_IN_WORDS - 1)] >> (pos & 31));
  int nb = 32 - (int)(pos & 31);
  wi++;
  uint32_t w = S.in[wi & (INF_IN_WORDS - 1)];            // next input word, fetched one step ahead of its use
  while (__any_sync(0xffffffffu, active)) {
#pragma unroll
    for (int rep = 0; rep < INF_STEPS_PER_VOTE; rep++) {
      if (nb <= 32) { bb |= (uint64_t)w << nb; nb += 32; wi++; }
      w = S.in[wi & (INF_IN_WORDS - 1)];
      uint32_t e = S.lit[(uint32_t)bb & ((1u << INF_LIT_ROOT) - 1)];
      if (active && (e & 0x80u) && ((e >> 12) & 7) == IK_LINK) {
        // a long code may still be a literal: fold it into a single-literal entry of root + sub bits
        const uint32_t e2 = S.lit[((e >> 16) & 0xfff) + ((uint32_t)(bb >> INF_LIT_ROOT) & ((1u << (e >> 28)) - 1))];
        if (!(e2 & 0x80u)) {
          const uint32_t cbt = INF_LIT_ROOT + (e2 & 15);
          e = (e2 & 0x00ff0000u) | (cbt << 12) | (1u << 4) | cbt;
        }
      }
      const bool flagged = (e & 0x80u) != 0;
      const int l1 = (e >> 12) & 15;
      int cb = e & 15, n = (e >> 4) & 3;
      if (!PAIRS || (n == 2 && pos + l1 >= seg_end)) { n = min(n, 1); cb = l1; }
      const bool take = active && !flagged;
      st = st || (active && flagged);
      if (!take) { cb = 0; n = 0; }
      if (WRITE && take) {
        S.ring[(out_at + count) & (INF_RING - 1)] = (uint8_t)(e >> 16);
        if (n == 2) S.ring[(out_at + count + 1) & (INF_RING - 1)] = (uint8_t)(e >> 24);
      }
      count += n; pos += cb;
      bb >>= cb; nb -= cb;
      if (take && count >= INF_SEG_CAP) st = true;      // (degenerate 1-bit codes) ends the round like a non-literal
      active = take && pos < seg_end && count < INF_SEG_CAP;
    }
  }
  *end = pos; *cnt = count; *stop = st;
}

__global__ void __launch_bounds__(32) k_inflate(const InflateDesc* __restrict__ descs, int* __restrict__ status,
                                                uint32_t* __restrict__ trailer /* may be null: the 4 bytes after the stream */) {
  extern __shared__ __align__(16) unsigned char inf_smem[];      // sizeof(InflateShared), opted in by the host when > 48 KB
  InflateShared& S = *reinterpret_cast<InflateShared*>(inf_smem);
  const int lane = threadIdx.x;
  const InflateDesc D = descs[blockIdx.x];
  const uint32_t in_words = (D.in_bytes + 3) >> 2;
  const uint4* in4 = reinterpret_cast<const uint4*>(D.in);

  // warp-uniform state
  uint32_t in_loaded = 0;      // words of the stream that are in the input ring
  uint32_t flushed = 0;        // bytes of the output that are in global memory (multiple of 16)
  uint32_t bitpos = 0;         // lane 0's position in the stream, in bits (broadcast after every serial section)
  int huff = 0;                // lane 0 is inside a Huffman block whose tables are built
  int pause = 0;               // rounds to go before the next parallel literal round is tried
  // lane-0 state
  uint64_t bitbuf = 0;
  int nbits = 0;
  uint32_t in_pos = 0, out_pos = 0;
  int stored_left = 0, last = 0;
  int saw_match = 0;
  uint32_t consumed = 0;
  bool in_block = false;
  int st = INF_OK;
  enum { GO_ON = 0, BUILD = 1, DONE = 2 };

  for (;;) {
    // ---- cooperative part: flush the finished 16-byte groups of the window, top up the input ring ----
    {
      const uint32_t target = min(out_pos, D.out_bytes) & ~15u;
      for (uint32_t p = flushed + lane * 16; p < target; p += 512)
        *reinterpret_cast<uint4*>(D.out + p) = *reinterpret_cast<const uint4*>(S.ring + (p & (INF_RING - 1)));
      flushed = max(flushed, target);
      // (lane 0 may still hold the two words before in_pos in its bit buffer, and a parallel round re-reads them)
      while (in_loaded < in_words && in_loaded - in_pos <= INF_IN_WORDS - 128 - 4) {
        const uint4 v = __ldg(in4 + (in_loaded >> 2) + lane);
        *reinterpret_cast<uint4*>(&S.in[(in_loaded + lane * 4) & (INF_IN_WORDS - 1)]) = v;
        in_loaded += 128;
      }
      __syncwarp();
    }
    // ---- parallel literal round ----
    bool par = false;
    uint32_t par_total = 0;
    if (huff && pause == 0 && (uint64_t)bitpos + 32 * INF_SEG + 96 <= (uint64_t)in_loaded * 32 &&
        INF_RING - INF_MARGIN - (out_pos - flushed) >= 32 * (INF_SEG_CAP + 2) && out_pos <= D.out_bytes) {
      par = true;
      const uint32_t base = bitpos;
      uint32_t start = base + lane * INF_SEG;
      const uint32_t seg_end = base + (lane + 1) * INF_SEG;
      uint32_t end, cnt;
      bool stop;
      inf_decode_segment<false>(S, start, seg_end, 0, true, &end, &cnt, &stop);
      for (int iter = 0; iter < 40; iter++) {
        const unsigned stopmask = __ballot_sync(0xffffffffu, stop);
        const int first = stopmask ? __ffs(stopmask) - 1 : 32;
        const uint32_t prev_end = __shfl_up_sync(0xffffffffu, end, 1);
        const uint32_t ns = lane == 0 ? base : prev_end;
        const bool changed = lane <= first && ns != start;      // lanes behind the first stop do not matter
        if (!__any_sync(0xffffffffu, changed)) break;
        uint32_t e1, c1;
        bool s1;
        inf_decode_segment<false>(S, changed ? ns : start, seg_end, 0, changed, &e1, &c1, &s1);
        if (changed) { start = ns; end = e1; cnt = c1; stop = s1; }
      }
      const unsigned stopmask = __ballot_sync(0xffffffffu, stop);
      const int first = stopmask ? __ffs(stopmask) - 1 : 32;
      const bool alive = lane <= first;
      uint32_t incl = alive ? cnt : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      {
        uint32_t e2, c2;
        bool s2;
        inf_decode_segment<true>(S, start, seg_end, out_pos + incl - cnt, alive, &e2, &c2, &s2);
      }
      bitpos = __shfl_sync(0xffffffffu, end, min(first, 31));
      out_pos += total;
      par_total = total;
      __syncwarp();
      if (lane == 0) {
        // the serial reader continues at the new position
        in_pos = bitpos >> 5;
        bitbuf = (uint64_t)(S.in[in_pos & (INF_IN_WORDS - 1)] >> (bitpos & 31));
        nbits = 32 - (int)(bitpos & 31);
        in_pos++;
      }
    } else if (pause > 0) {
      pause--;
    }
    int why = GO_ON;
    saw_match = 0;
    if (lane == 0) {
      const bool in_done = in_loaded >= in_words;
#define INF_REFILL()                                                             \
  do {                                                                           \
    if (nbits < 32) {                                                            \
      const uint32_t w_ = in_pos < in_words ? S.in[in_pos & (INF_IN_WORDS - 1)] : 0u; \
      bitbuf |= (uint64_t)w_ << nbits;                                           \
      nbits += 32;                                                               \
      in_pos++;                                                                  \
    }                                                                            \
  } while (0)
#define INF_TAKE(n_) (tmp = (uint32_t)bitbuf & ((1u << (n_)) - 1), bitbuf >>= (n_), nbits -= (n_), tmp)
      uint32_t tmp;
      for (;;) {
        // every step below needs at most two refills: stop for input (unless it is all here) and for window space
        if (out_pos - flushed > INF_RING - INF_MARGIN) break;
        if (!in_done && in_loaded - in_pos < 4) break;
        if (in_pos > in_words + 4) { st = INF_TRUNCATED; why = DONE; break; }
        if (out_pos > D.out_bytes) { st = INF_SIZE; why = DONE; break; }
        if (!in_block) {
          if (last) {
            consumed = in_pos * 4 - (uint32_t)(nbits >> 3);      // bytes really consumed
            if (consumed > D.in_bytes) st = INF_TRUNCATED;
            why = DONE;
            break;
          }
          // a dynamic header can take ~570 bytes: have them in the ring before starting
          if (!in_done && in_loaded - in_pos < 160) break;
          INF_REFILL();
          last = INF_TAKE(1);
          const int type = INF_TAKE(2);
          if (type == 0) {
            const int drop = nbits & 7;
            bitbuf >>= drop; nbits -= drop;
            INF_REFILL();
            const uint32_t len = INF_TAKE(16);
            INF_REFILL();
            const uint32_t nlen = INF_TAKE(16);
            if ((len ^ 0xffffu) != nlen) { st = INF_CORRUPT; why = DONE; break; }
            stored_left = (int)len;
            in_block = true;
            if (!stored_left) in_block = false;
            continue;
          }
          if (type == 3) { st = INF_CORRUPT; why = DONE; break; }
          if (type == 1) {
            for (int i = 0; i < 144; i++) S.lens[i] = 8;
            for (int i = 144; i < 256; i++) S.lens[i] = 9;
            for (int i = 256; i < 280; i++) S.lens[i] = 7;
            for (int i = 280; i < 288; i++) S.lens[i] = 8;
            for (int i = 288; i < 320; i++) S.lens[i] = 5;   // 32 distance codes, 30 and 31 never valid
            S.codes[318] = 288; S.codes[319] = 32;     // hlit, hdist handed to the build step
          } else {
            const int hlit = INF_TAKE(5) + 257, hdist = INF_TAKE(5) + 1, hclen = INF_TAKE(4) + 4;
            if (hlit > 286 || hdist > 30) { st = INF_CORRUPT; why = DONE; break; }
            // code-length code: at most 7 bits, decoded by a direct 128-entry table kept in the (not yet built) dist area
            uint32_t* cl = S.dist;
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < hclen; i++) {
              INF_REFILL();
              const int v = INF_TAKE(3);
              S.lens[c_cl_order[i]] = (uint8_t)v;      // staged in lens[0..19), overwritten by the real lengths below
              cnt[v]++;
            }
            for (int i = hclen; i < 19; i++) S.lens[c_cl_order[i]] = 0;
            int left = 1;
            uint32_t nxt[8], code = 0;
            cnt[0] = 0;
            bool bad = false;
            for (int l = 1; l <= 7; l++) {
              code = (code + cnt[l - 1]) << 1;
              nxt[l] = code;
              left = (left << 1) - cnt[l];
              if (left < 0) bad = true;
            }
            if (bad || left > 0) { st = INF_CORRUPT; why = DONE; break; }
            for (int i = 0; i < 128; i++) cl[i] = 0;
            for (int s = 0; s < 19; s++) {
              const int l = S.lens[s];
              if (!l) continue;
              const uint32_t r = __brev(nxt[l]++) >> (32 - l);
              for (int i = r; i < 128; i += 1 << l) cl[i] = (s << 8) | l;
            }
            int n = 0;
            const int total = hlit + hdist;
            bool fail = false;
            while (n < total) {
              INF_REFILL();
              const uint32_t e = cl[(uint32_t)bitbuf & 127];
              const int cb = e & 0xff;
              if (!cb) { fail = true; break; }
              bitbuf >>= cb; nbits -= cb;
              const int sym = e >> 8;
              if (sym < 16) { S.lens[n++] = (uint8_t)sym; continue; }
              int rep, val = 0;
              if (sym == 16) {
                if (n == 0) { fail = true; break; }
                val = S.lens[n - 1];
                rep = 3 + INF_TAKE(2);
              } else if (sym == 17) rep = 3 + INF_TAKE(3);
              else rep = 11 + INF_TAKE(7);
              if (n + rep > total) { fail = true; break; }
              while (rep--) S.lens[n++] = (uint8_t)val;
            }
            if (fail || S.lens[256] == 0) { st = INF_CORRUPT; why = DONE; break; }
            S.codes[318] = (uint16_t)hlit; S.codes[319] = (uint16_t)hdist;
          }
          in_block = true;
          stored_left = -1;                  // marks a Huffman block
          why = BUILD;
          break;
        }
        if (stored_left >= 0) {
          // stored block: bytes straight through the bit reader (rare in PNG files)
          INF_REFILL();
          S.ring[out_pos & (INF_RING - 1)] = (uint8_t)INF_TAKE(8);
          out_pos++;
          if (--stored_left == 0) in_block = false;
          continue;
        }
        // ---- Huffman block: the serial walk ----
        INF_REFILL();
        const uint32_t e0 = S.lit[(uint32_t)bitbuf & ((1u << INF_LIT_ROOT) - 1)];
        if (par) {
          // literals belong to the next parallel round: only lengths, links and end-of-block are handled here
          if (!(e0 & 0x80u)) break;
        } else if (!(e0 & 0x80u)) {             // (a match right after a match skips the set-up of the literal run)
          // Literal run.  No bookkeeping inside: a probe emits at most 2 bytes and takes at most INF_LIT_ROOT bits, so
          // the window room and the input words at hand bound the number of probes up front; a probe that meets a
          // length / end-of-block / link entry reads zeros in the fields used here and does nothing, and the run ends
          // at the next group boundary.  Three probes per refill check: 33 valid bits cover 3 x 11 (an empty buffer takes
          // two words to get there).
          const uint32_t limit = in_done ? in_words : in_loaded;
          int n = (int)(INF_RING - INF_MARGIN - (out_pos - flushed)) >> 1;
          n = min(n, in_pos + 3 < limit ? (int)(limit - in_pos - 3) * 2 : 0);
          if (n > 0) {
            uint32_t wnext = S.in[in_pos & (INF_IN_WORDS - 1)];      // next input word, loaded ahead of its use
            uint32_t seen = 0;
            do {
              if (nbits <= 32) {
                bitbuf |= (uint64_t)wnext << nbits;
                nbits += 32;
                in_pos++;
                wnext = S.in[in_pos & (INF_IN_WORDS - 1)];
                if (nbits == 32) {           // the buffer was empty: one word is one bit short of three probes
                  bitbuf |= (uint64_t)wnext << 32;
                  nbits = 64;
                  in_pos++;
                  wnext = S.in[in_pos & (INF_IN_WORDS - 1)];
                }
              }
#pragma unroll
              for (int k = 0; k < 3; k++) {
                const uint32_t e = S.lit[(uint32_t)bitbuf & ((1u << INF_LIT_ROOT) - 1)];
                seen |= e;
                S.ring[out_pos & (INF_RING - 1)] = (uint8_t)(e >> 16);
                S.ring[(out_pos + 1) & (INF_RING - 1)] = (uint8_t)(e >> 24);
                out_pos += (e >> 4) & 3;
                const int cb = e & 15;
                bitbuf >>= cb; nbits -= cb;
              }
              n -= 3;
            } while (!(seen & 0x80u) && n > 0);
            INF_REFILL();
          }
        }
        uint32_t e = S.lit[(uint32_t)bitbuf & ((1u << INF_LIT_ROOT) - 1)];
        if (!(e & 0x80u)) {
          // one or two literals: the second byte lands on a position the next step overwrites when it is not used
          S.ring[out_pos & (INF_RING - 1)] = (uint8_t)(e >> 16);
          S.ring[(out_pos + 1) & (INF_RING - 1)] = (uint8_t)(e >> 24);
          out_pos += (e >> 4) & 3;
          const int cb = e & 15;
          if (!cb) { st = INF_CORRUPT; why = DONE; break; }
          bitbuf >>= cb; nbits -= cb;
          continue;
        }
        uint32_t kind = (e >> 12) & 7;
        if (kind == IK_LINK) {
          bitbuf >>= INF_LIT_ROOT; nbits -= INF_LIT_ROOT;
          e = S.lit[((e >> 16) & 0xfff) + ((uint32_t)bitbuf & ((1u << (e >> 28)) - 1))];
          if (!(e & 0x80u)) {
            const int cb = e & 15;
            if (!cb) { st = INF_CORRUPT; why = DONE; break; }
            S.ring[out_pos & (INF_RING - 1)] = (uint8_t)(e >> 16);
            out_pos++;
            bitbuf >>= cb; nbits -= cb;
            continue;
          }
          kind = (e >> 12) & 7;
        }
        {
          const int cb = (e >> 8) & 15;
          bitbuf >>= cb; nbits -= cb;
        }
        if (kind == IK_EOB) { in_block = false; continue; }
        if (kind != IK_LEN) { st = INF_CORRUPT; why = DONE; break; }
        const uint32_t length = ((e >> 16) & 0xfff) + INF_TAKE(e >> 28);
        INF_REFILL();
        uint32_t d = S.dist[(uint32_t)bitbuf & ((1u << INF_DIST_ROOT) - 1)];
        if (((d >> 4) & 7) == IK_LINK) {
          bitbuf >>= INF_DIST_ROOT; nbits -= INF_DIST_ROOT;
          d = S.dist[(d >> 16) + ((uint32_t)bitbuf & ((1u << ((d >> 8) & 15)) - 1))];
        }
        if (((d >> 4) & 7) != IK_DIST) { st = INF_CORRUPT; why = DONE; break; }
        {
          const int cb = d & 15;
          bitbuf >>= cb; nbits -= cb;
        }
        const uint32_t distance = (d >> 16) + INF_TAKE((d >> 8) & 15);
        if (distance > out_pos) { st = INF_CORRUPT; why = DONE; break; }
        saw_match = 1;
        uint32_t src = out_pos - distance;
        if (distance < INF_NEAR) {
          // source still in the window
          if (distance >= length) {
            uint32_t i = 0;
            for (; i + 4 <= length; i += 4) {
              const uint8_t b0 = S.ring[(src + i) & (INF_RING - 1)], b1 = S.ring[(src + i + 1) & (INF_RING - 1)];
              const uint8_t b2 = S.ring[(src + i + 2) & (INF_RING - 1)], b3 = S.ring[(src + i + 3) & (INF_RING - 1)];
              S.ring[(out_pos + i) & (INF_RING - 1)] = b0; S.ring[(out_pos + i + 1) & (INF_RING - 1)] = b1;
              S.ring[(out_pos + i + 2) & (INF_RING - 1)] = b2; S.ring[(out_pos + i + 3) & (INF_RING - 1)] = b3;
            }
            for (; i < length; i++) S.ring[(out_pos + i) & (INF_RING - 1)] = S.ring[(src + i) & (INF_RING - 1)];
          } else {
            for (uint32_t i = 0; i < length; i++) S.ring[(out_pos + i) & (INF_RING - 1)] = S.ring[(src + i) & (INF_RING - 1)];
          }
        } else {
          // source already flushed (the loop head keeps out_pos - flushed below the window size minus the margin)
          const uint8_t* g = D.out + src;
          uint32_t i = 0;
          for (; i + 4 <= length; i += 4) {
            const uint8_t b0 = __ldcg(g + i), b1 = __ldcg(g + i + 1), b2 = __ldcg(g + i + 2), b3 = __ldcg(g + i + 3);
            S.ring[(out_pos + i) & (INF_RING - 1)] = b0; S.ring[(out_pos + i + 1) & (INF_RING - 1)] = b1;
            S.ring[(out_pos + i + 2) & (INF_RING - 1)] = b2; S.ring[(out_pos + i + 3) & (INF_RING - 1)] = b3;
          }
          for (; i < length; i++) S.ring[(out_pos + i) & (INF_RING - 1)] = __ldcg(g + i);
        }
        out_pos += length;
      }
#undef INF_REFILL
#undef INF_TAKE
    }
    why = __shfl_sync(0xffffffffu, why, 0);
    out_pos = __shfl_sync(0xffffffffu, out_pos, 0);
    in_pos = __shfl_sync(0xffffffffu, in_pos, 0);
    bitpos = __shfl_sync(0xffffffffu, in_pos * 32u - (uint32_t)nbits, 0);
    huff = __shfl_sync(0xffffffffu, (int)(in_block && stored_left < 0), 0);
    // short literal runs between matches do not pay for a parallel round: go serial for a while
    if (par && par_total < INF_PAR_MIN && __shfl_sync(0xffffffffu, saw_match, 0)) pause = INF_PAR_PAUSE;
    __syncwarp();
    if (why == BUILD) {
      const int hlit = S.codes[318], hdist = S.codes[319];
      __syncwarp();
      int rc = inf_build_table(S, hlit, hdist, true, S.dist, INF_DIST_ROOT, INF_DIST_SUB, lane);
      if (!rc) rc = inf_build_table(S, 0, hlit, false, S.lit, INF_LIT_ROOT, INF_LIT_SUB, lane);
      if (rc) { st = rc; why = DONE; }
    }
    if (why == DONE) break;
  }
  st = __shfl_sync(0xffffffffu, st, 0);
  if (st == INF_OK && out_pos != D.out_bytes) st = INF_SIZE;
  // tail of the window
  if (st == INF_OK) {
    const uint32_t target = out_pos & ~15u;
    for (uint32_t p = flushed + lane * 16; p < target; p += 512)
      *reinterpret_cast<uint4*>(D.out + p) = *reinterpret_cast<const uint4*>(S.ring + (p & (INF_RING - 1)));
    for (uint32_t p = target + lane; p < out_pos; p += 32) D.out[p] = S.ring[p & (INF_RING - 1)];
  }
  if (lane == 0) {
    status[blockIdx.x] = st;
    // zlib framing: the big-endian Adler-32 of the output follows the stream (k_unfilter checks it)
    if (trailer && st == INF_OK) {
      const uint8_t* t = D.in + consumed;
      trailer[blockIdx.x] = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PNG reconstruction (PNG spec 9.2) of 8-bit gray scanlines: raw = h x (1 filter byte + w filtered bytes) -> level 0.
// Lane r of the warp owns scanline row0 + r and runs r pixels behind the lane above it.
__device__ __forceinline__ int inf_paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

constexpr int UNF_WARPS = 4;
constexpr int UNF_LEAD = 16;      // bytes in front of a frame's scanlines in its slot (the prefetch may look behind row 0)

__global__ void __launch_bounds__(UNF_WARPS * 32) k_unfilter(const uint8_t* __restrict__ raw, size_t raw_slot, uint8_t* frames,
                                                             size_t frame_slot, int pitch, int w, int h, int n_frames,
                                                             int* status, const uint32_t* __restrict__ adler_expect) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * UNF_WARPS + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  if (status[f] != INF_OK) return;
  const uint8_t* R = raw + (size_t)f * raw_slot + UNF_LEAD;
  uint8_t* Dst = frames + (size_t)f * frame_slot;
  const int rb = w + 1;
  bool bad = false;
  // Adler-32 of the scanlines as they pass: a = 1 + sum(x_i), b = n + sum((n - i) * x_i)  (mod 65521)
  const uint32_t n_raw = (uint32_t)rb * h;
  unsigned long long sum_a = 0, sum_b = 0;
  for (int row0 = 0; row0 < h; row0 += 32) {
    const int row = row0 + lane;
    const bool valid = row < h;
    const uint8_t* line = R + (size_t)(valid ? row : 0) * rb;
    const int ft = valid ? line[0] : 0;
    if (ft > 4) bad = true;
    uint32_t weight = n_raw - (uint32_t)row * rb;      // n - i of the filter byte; one less per following byte
    sum_a += ft; sum_b += (unsigned long long)weight * ft;
    const uintptr_t s = (uintptr_t)(line + 1);          // the filtered bytes start at an arbitrary address
    const uint8_t* up_row = Dst + (size_t)(row0 - 1) * pitch;
    uint8_t* out_row = Dst + (size_t)(valid ? row : 0) * pitch;
    uint32_t pack = 0;
    int a = 0, b = 0, c = 0, prev_out = 0;
    // 16 steps per chunk; the words of the next chunk are requested while this one is processed
    uint32_t nw0 = 0, nw1 = 0, nw2 = 0, nw3 = 0, nw4 = 0;
    uint4 nup = make_uint4(0, 0, 0, 0);
    auto prefetch = [&](int xn) {
      if (valid && xn + 15 >= 0 && xn < w) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>((s + (intptr_t)xn) & ~(uintptr_t)3);
        nw0 = wp[0]; nw1 = wp[1]; nw2 = wp[2]; nw3 = wp[3]; nw4 = wp[4];
      }
      if (lane == 0 && row0 > 0 && xn < w) nup = __ldcg(reinterpret_cast<const uint4*>(up_row + xn));
    };
    prefetch(-lane);
    for (int t0 = 0; t0 < w + 31; t0 += 16) {
      const int x0 = t0 - lane;
      const int o8 = (int)((s + (intptr_t)x0) & 3) * 8;
      uint32_t v[4];
      v[0] = __funnelshift_r(nw0, nw1, o8); v[1] = __funnelshift_r(nw1, nw2, o8);
      v[2] = __funnelshift_r(nw2, nw3, o8); v[3] = __funnelshift_r(nw3, nw4, o8);
      const uint32_t up[4] = {nup.x, nup.y, nup.z, nup.w};
      prefetch(x0 + 16);
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int x = x0 + j;
        // what the lane above produced one step ago is the pixel above this lane's x
        int above = __shfl_up_sync(0xffffffffu, prev_out, 1);
        if (lane == 0) above = row0 > 0 ? (int)((up[j >> 2] >> ((j & 3) * 8)) & 0xff) : 0;
        const bool active = valid && (unsigned)x < (unsigned)w;
        const int cur = (v[j >> 2] >> ((j & 3) * 8)) & 0xff;
        if (active) {
          weight--;
          sum_a += cur; sum_b += (unsigned long long)weight * cur;
          c = b;
          b = above;
          if (x == 0) { a = 0; c = 0; }
          int pred = 0;
          if (ft == 1) pred = a;
          else if (ft == 2) pred = b;
          else if (ft == 3) pred = (a + b) >> 1;
          else if (ft == 4) pred = inf_paeth(a, b, c);
          const int o = (cur + pred) & 0xff;
          a = o;
          prev_out = o;
          pack |= (uint32_t)o << ((x & 3) * 8);
          if ((x & 3) == 3 || x == w - 1) {
            // rows are 16-byte multiples with at least one spare byte: the last word may cover up to 3 pad bytes (zeros)
            *reinterpret_cast<uint32_t*>(out_row + (x & ~3)) = pack;
            pack = 0;
          }
        }
      }
    }
    __syncwarp();
  }
  for (int d = 16; d; d >>= 1) {
    sum_a += __shfl_xor_sync(0xffffffffu, sum_a, d);
    sum_b += __shfl_xor_sync(0xffffffffu, sum_b, d);
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    const uint32_t a = (uint32_t)((1 + sum_a) % 65521u), b = (uint32_t)((n_raw + sum_b) % 65521u);
    if (bad) status[f] = INF_FILTER;
    else if (adler_expect && adler_expect[f] != ((b << 16) | a)) status[f] = INF_CHECKSUM;
  }
}

// ---------------------------------------------------------------------------------------------
// CRC-32 of the IDAT chunks (PNG spec 5.3: over the chunk type and data), one thread per chunk, slicing-by-4 tables built in
// shared memory.  The host lists (offset into the frame's concatenated payload, length, stored CRC) for up to
// PNG_CRC_CAP chunks per frame while it copies the payloads; libpng writes 8 KB chunks, i.e. ~35 per KITTI frame.
constexpr int PNG_CRC_CAP = 64;
constexpr int PNG_PAYLOAD_OFS = 14;     // the zlib stream starts here in a frame's slot, so that the deflate data is 16-byte aligned

__global__ void __launch_bounds__(PNG_CRC_CAP) k_png_crc(const uint8_t* __restrict__ comp, size_t slot, const uint32_t* __restrict__ descs,
                                                        const int* __restrict__ counts, int* status) {
  __shared__ uint32_t T[4][256];
  const int t = threadIdx.x, f = blockIdx.x;
  for (int i = t; i < 256; i += PNG_CRC_CAP) {
    uint32_t c = i;
#pragma unroll
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    T[0][i] = c;
  }
  __syncthreads();
  for (int i = t; i < 256; i += PNG_CRC_CAP) {
    uint32_t c = T[0][i];
#pragma unroll
    for (int s = 1; s < 4; s++) { c = (c >> 8) ^ T[0][c & 0xff]; T[s][i] = c; }
  }
  __syncthreads();
  if (t >= counts[f]) return;
  const uint32_t* d = descs + ((size_t)f * PNG_CRC_CAP + t) * 3;
  const uint8_t* p = comp + (size_t)f * slot + PNG_PAYLOAD_OFS + d[0];
  uint32_t len = d[1];
  uint32_t c = 0xffffffffu;
  c = T[0][(c ^ 'I') & 0xff] ^ (c >> 8); c = T[0][(c ^ 'D') & 0xff] ^ (c >> 8);
  c = T[0][(c ^ 'A') & 0xff] ^ (c >> 8); c = T[0][(c ^ 'T') & 0xff] ^ (c >> 8);
  while (len && ((uintptr_t)p & 3)) { c = T[0][(c ^ *p++) & 0xff] ^ (c >> 8); len--; }
  while (len >= 4) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p) ^ c;
    c = T[3][w & 0xff] ^ T[2][(w >> 8) & 0xff] ^ T[1][(w >> 16) & 0xff] ^ T[0][w >> 24];
    p += 4; len -= 4;
  }
  while (len--) c = T[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  if (~c != d[2]) status[f] = INF_CRC;
}

}  // namespace orbk
