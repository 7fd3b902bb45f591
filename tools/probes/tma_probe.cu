// tma_probe.cu -- checks the tensor-map encoding and the TMA forms the ORB kernels use, and times them (sm_100a):
//   (A) per-keypoint windows: 40x39 u16 box-sum window + 32x31 u8 patch, one warp per keypoint, ring of DEPTH slots;
//   (B) 160x72 u8 tile with negative / out-of-range start coordinates (zero fill).
// Each result is compared with plain global loads.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeTiledFn)fn;
}

static CUtensorMap make_map3(EncodeTiledFn enc, CUtensorMapDataType dt, int esize, void* base, uint64_t w, uint64_t h, uint64_t f,
                             uint64_t pitch_bytes, uint64_t frame_bytes, uint32_t bw, uint32_t bh) {
  CUtensorMap m;
  cuuint64_t dims[3] = {w, h, f};
  cuuint64_t strides[2] = {pitch_bytes, frame_bytes};
  cuuint32_t box[3] = {bw, bh, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, dt, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (esize %d)\n", (int)r, esize); exit(1); }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load3(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

constexpr int BW = 48, BH = 39, PW = 48, PH = 31;   // box starts must be 16-byte aligned in x: aligned supersets of the 39 / 31 wide windows
constexpr int BOX_BYTES = BW * BH * 2, PAT_BYTES = PW * PH;
constexpr int BOX_SLOT = ((BOX_BYTES + 127) / 128) * 128;   // TMA destinations are 128-byte aligned
constexpr int SLOT_BYTES = BOX_SLOT + ((PAT_BYTES + 127) / 128) * 128;

struct Kp { int x, y, f; };

// one warp per keypoint stream; mode 0: TMA ring, mode 1: plain global gathers (reference + baseline timing)
template <int DEPTH, int MODE>
__global__ void __launch_bounds__(128) k_windows(const CUtensorMap* maps, const uint16_t* box, const uint8_t* img, int W, int H, int bp, int p,
                                                 const Kp* kps, int per_warp, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[4 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * 4 + warp;
  const Kp* mine = kps + (size_t)gw * per_warp;
  uint8_t* slots = smem + (size_t)warp * DEPTH * SLOT_BYTES;
  uint64_t* bar = bars + warp * 4;
  if (MODE == 0) {
    if (lane == 0) for (int d = 0; d < DEPTH; d++) mbar_init(bar + d, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (int d = 0; d < DEPTH && d < per_warp; d++) {
        mbar_expect(bar + d, BOX_BYTES + PAT_BYTES);
        tma_load3(slots + d * SLOT_BYTES, maps + 0, bar + d, (mine[d].x - 19) & ~7, mine[d].y - 19, mine[d].f);
        tma_load3(slots + d * SLOT_BYTES + BOX_SLOT, maps + 1, bar + d, (mine[d].x - 15) & ~15, mine[d].y - 15, mine[d].f);
      }
  }
  uint32_t acc = 0;
  for (int i = 0; i < per_warp; i++) {
    const Kp k = mine[i];
    const int d = i % DEPTH;
    if (MODE == 0) {
      mbar_wait(bar + d, (i / DEPTH) & 1);
      const uint16_t* bw = (const uint16_t*)(slots + d * SLOT_BYTES) + ((k.x - 19) & 7);
      const uint8_t* pw = slots + d * SLOT_BYTES + BOX_SLOT + ((k.x - 15) & 15);
#pragma unroll
      for (int t = 0; t < 16; t++) {
        const int dx = ((lane * 7 + t * 13) % 39) - 19, dy = ((lane * 11 + t * 5) % 39) - 19;
        acc = acc * 31 + bw[(dy + 19) * BW + dx + 19];
      }
      if (lane < 31)
#pragma unroll
        for (int r = 0; r < 31; r++) acc += pw[r * PW + lane] * (r + 1);
      __syncwarp();
      if (lane == 0 && i + DEPTH < per_warp) {
        const Kp n = mine[i + DEPTH];
        mbar_expect(bar + d, BOX_BYTES + PAT_BYTES);
        tma_load3(slots + d * SLOT_BYTES, maps + 0, bar + d, (n.x - 19) & ~7, n.y - 19, n.f);
        tma_load3(slots + d * SLOT_BYTES + BOX_SLOT, maps + 1, bar + d, (n.x - 15) & ~15, n.y - 15, n.f);
      }
    } else {
      const uint16_t* bf = box + (size_t)k.f * H * bp;
      const uint8_t* pf = img + (size_t)k.f * H * p;
#pragma unroll
      for (int t = 0; t < 16; t++) {
        const int dx = ((lane * 7 + t * 13) % 39) - 19, dy = ((lane * 11 + t * 5) % 39) - 19;
        const int xx = k.x + dx, yy = k.y + dy;
        const uint32_t v = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? bf[(size_t)yy * bp + xx] : 0;
        acc = acc * 31 + v;
      }
      if (lane < 31)
#pragma unroll
        for (int r = 0; r < 31; r++) {
          const int xx = k.x - 15 + lane, yy = k.y - 15 + r;
          const uint32_t v = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? pf[(size_t)yy * p + xx] : 0;
          acc += v * (r + 1);
        }
    }
  }
  out[(size_t)gw * 32 + lane] = acc;
}

// (B) tile loads: one CTA per tile, 160x72 bytes at (x0-16, y0-4); checksum per thread of its 45 bytes
template <int MODE>
__global__ void __launch_bounds__(256) k_tiles(const CUtensorMap* maps, const uint8_t* img, int W, int H, int p, int tiles_x, int tiles_y,
                                               uint32_t* out) {
  __shared__ __align__(128) uint8_t tile[160 * 72];
  __shared__ __align__(8) uint64_t bar;
  const int t = blockIdx.x, f = blockIdx.y;
  const int x0 = (t % tiles_x) * 128, y0 = (t / tiles_x) * 64;
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect(&bar, 160 * 72);
      tma_load3(tile, maps + 2, &bar, x0 - 16, y0 - 4, f);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
  } else {
    for (int i = threadIdx.x; i < 160 * 72; i += 256) {
      const int xx = x0 - 16 + i % 160, yy = y0 - 4 + i / 160;
      tile[i] = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? img[((size_t)f * H + yy) * p + xx] : 0;
    }
    __syncthreads();
  }
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < 160 * 72; i += 256) acc = acc * 131 + tile[i];
  out[((size_t)f * gridDim.x + t) * 256 + threadIdx.x] = acc;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CKC(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  CKC(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const int W = 1241, H = 376, F = 256, P = 1248, BP = 1248;
  EncodeTiledFn enc = get_encode();
  uint16_t* d_box; uint8_t* d_img;
  CKC(cudaMalloc(&d_box, (size_t)F * H * BP * 2));
  CKC(cudaMalloc(&d_img, (size_t)F * H * P));
  {
    std::vector<uint16_t> hb((size_t)F * H * BP);
    std::vector<uint8_t> hi((size_t)F * H * P);
    uint32_t s = 12345;
    for (auto& v : hb) { s = s * 1664525u + 1013904223u; v = (uint16_t)(s >> 19); }
    for (auto& v : hi) { s = s * 1664525u + 1013904223u; v = (uint8_t)(s >> 24); }
    CKC(cudaMemcpy(d_box, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(d_img, hi.data(), hi.size(), cudaMemcpyHostToDevice));
  }
  // tensor dims are the *valid* extents (w, h): everything outside reads as zero
  CUtensorMap hm[3];
  hm[0] = make_map3(enc, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d_box, W, H, F, (uint64_t)BP * 2, (uint64_t)H * BP * 2, BW, BH);
  hm[1] = make_map3(enc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_img, W, H, F, P, (uint64_t)H * P, PW, PH);
  hm[2] = make_map3(enc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_img, W, H, F, P, (uint64_t)H * P, 160, 72);
  CUtensorMap* d_maps;
  CKC(cudaMalloc(&d_maps, sizeof(hm)));
  CKC(cudaMemcpy(d_maps, hm, sizeof(hm), cudaMemcpyHostToDevice));

  // ---- (A) keypoint windows
  const int nblk = 148 * 12, per_warp = 64, nwarps = nblk * 4;
  std::vector<Kp> kps((size_t)nwarps * per_warp);
  uint32_t s = 777;
  for (auto& k : kps) {
    s = s * 1664525u + 1013904223u; k.x = (s >> 8) % W;
    s = s * 1664525u + 1013904223u; k.y = (s >> 8) % H;
    s = s * 1664525u + 1013904223u; k.f = (s >> 8) % F;
  }
  Kp* d_kps; uint32_t *d_o1, *d_o2;
  CKC(cudaMalloc(&d_kps, kps.size() * sizeof(Kp)));
  CKC(cudaMemcpy(d_kps, kps.data(), kps.size() * sizeof(Kp), cudaMemcpyHostToDevice));
  CKC(cudaMalloc(&d_o1, (size_t)nwarps * 32 * 4));
  CKC(cudaMalloc(&d_o2, (size_t)nwarps * 32 * 4));
  std::vector<uint32_t> o1((size_t)nwarps * 32), o2((size_t)nwarps * 32);
  auto check = [&](const char* what) {
    CKC(cudaMemcpy(o1.data(), d_o1, o1.size() * 4, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(o2.data(), d_o2, o2.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < o1.size(); i++) bad += o1[i] != o2[i];
    printf("%s: %zu mismatching lanes of %zu\n", what, bad, o1.size());
  };
  float ms_ref = time_ms([&] { k_windows<1, 1><<<nblk, 128, 0>>>(d_maps, d_box, d_img, W, H, BP, P, d_kps, per_warp, d_o2); }, 5);
  CKC(cudaGetLastError());
  printf("windows, global gathers     : %.3f ms for %d keypoints = %.1f ns per keypoint per SM-slot, %.2f M kp/s\n", ms_ref, nwarps * per_warp,
         0.0, nwarps * per_warp / ms_ref / 1e3);
#define RUN_DEPTH(D)                                                                                                       \
  {                                                                                                                        \
    CKC(cudaFuncSetAttribute(k_windows<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * D * SLOT_BYTES));           \
    float ms = time_ms([&] { k_windows<D, 0><<<nblk, 128, 4 * D * SLOT_BYTES>>>(d_maps, d_box, d_img, W, H, BP, P, d_kps, per_warp, d_o1); }, 5); \
    CKC(cudaGetLastError());                                                                                               \
    printf("windows, TMA ring depth %d   : %.3f ms, %.2f M kp/s, smem/CTA %d B\n", D, ms, nwarps * per_warp / ms / 1e3, 4 * D * SLOT_BYTES); \
    check("  windows vs global");                                                                                          \
  }
  RUN_DEPTH(1) RUN_DEPTH(2) RUN_DEPTH(3) RUN_DEPTH(4)

  // ---- (B) tiles
  const int tx = (W + 127) / 128, ty = (H + 63) / 64;
  uint32_t *d_t1, *d_t2;
  CKC(cudaMalloc(&d_t1, (size_t)F * tx * ty * 256 * 4));
  CKC(cudaMalloc(&d_t2, (size_t)F * tx * ty * 256 * 4));
  float ms_t0 = time_ms([&] { k_tiles<0><<<dim3(tx * ty, F), 256>>>(d_maps, d_img, W, H, P, tx, ty, d_t1); }, 5);
  CKC(cudaGetLastError());
  float ms_t1 = time_ms([&] { k_tiles<1><<<dim3(tx * ty, F), 256>>>(d_maps, d_img, W, H, P, tx, ty, d_t2); }, 5);
  CKC(cudaGetLastError());
  {
    std::vector<uint32_t> a((size_t)F * tx * ty * 256), b(a.size());
    CKC(cudaMemcpy(a.data(), d_t1, a.size() * 4, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(b.data(), d_t2, b.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < a.size(); i++) bad += a[i] != b[i];
    printf("tiles 160x72: TMA %.3f ms, byte loads %.3f ms for %d tiles (%.1f GB/s of image); %zu mismatches\n", ms_t0, ms_t1, F * tx * ty,
           (double)F * W * H / ms_t0 / 1e6, bad);
  }
  return 0;
}
