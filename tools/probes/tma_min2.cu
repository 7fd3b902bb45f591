// TMA bring-up 2: u16 windows 40x39, negative / out-of-range coordinates, one warp with lane-0 issue
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#ifndef BW_
#define BW_ 40
#endif
#ifndef BH_
#define BH_ 39
#endif
#ifndef ET
#define ET uint16_t
#define EDT CU_TENSOR_MAP_DATA_TYPE_UINT16
#endif
constexpr int BW = BW_, BH = BH_;
typedef ET elem_t;
__global__ void k(const CUtensorMap* gmap, uint32_t* out, int c0, int c1, int c2, int bytes) {
  __shared__ __align__(128) elem_t win[BW * BH];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(win)),
                 "l"(gmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra W;\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  uint32_t acc = 0;
  for (int i = lane; i < BW * BH; i += 32) acc = acc * 131 + win[i];
  out[lane] = acc;
}
int main(int argc, char** argv) {
  const int W = 1241, H = 376, F = 4, BP = 1248;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  elem_t* d_box; CKC(cudaMalloc(&d_box, (size_t)F * H * BP * sizeof(elem_t)));
  std::vector<elem_t> hb((size_t)F * H * BP);
  uint32_t s = 1; for (auto& v : hb) { s = s * 1664525u + 1013904223u; v = (elem_t)(s >> 19); }
  CKC(cudaMemcpy(d_box, hb.data(), hb.size() * sizeof(elem_t), cudaMemcpyHostToDevice));
  CUtensorMap m3;
  { cuuint64_t dims[3] = {W, H, F}; cuuint64_t st[2] = {(cuuint64_t)BP * sizeof(elem_t), (cuuint64_t)BP * sizeof(elem_t) * H}; cuuint32_t box[3] = {BW, BH, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m3, EDT, 3, d_box, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode u16 3d rc=%d\n", (int)r); }
  CUtensorMap* d_maps; CKC(cudaMalloc(&d_maps, sizeof(CUtensorMap)));
  CKC(cudaMemcpy(d_maps, &m3, sizeof(m3), cudaMemcpyHostToDevice));
  uint32_t* d_out; CKC(cudaMalloc(&d_out, 32 * 4));
  std::vector<uint32_t> o(32), ref(32);
  int cases[1][3] = {{atoi(argv[1]), atoi(argv[2]), atoi(argv[3])}};
  for (auto& c : cases) {
    for (int t = 0; t < 32; t++) { uint32_t acc = 0; for (int i = t; i < BW * BH; i += 32) { int xx = c[0] + i % BW, yy = c[1] + i / BW; acc = acc * 131 + ((xx >= 0 && xx < W && yy >= 0 && yy < H) ? hb[((size_t)c[2] * H + yy) * BP + xx] : 0); } ref[t] = acc; }
    k<<<1, 32>>>(d_maps, d_out, c[0], c[1], c[2], BW * BH * (int)sizeof(elem_t));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case (%d,%d,%d): %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
    CKC(cudaMemcpy(o.data(), d_out, 32 * 4, cudaMemcpyDeviceToHost));
    int bad = 0; for (int t = 0; t < 32; t++) bad += o[t] != ref[t];
    printf("case (%d,%d,%d): ok, %d mismatches\n", c[0], c[1], c[2], bad);
  }
  return 0;
}
