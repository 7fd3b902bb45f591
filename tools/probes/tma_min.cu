// minimal TMA bring-up: variants selected by argv[1]
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

template <int VAR>
__global__ void k(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, uint32_t* out, int c0, int c1, int c2) {
  __shared__ __align__(128) uint8_t tile[160 * 72];
  __shared__ __align__(8) uint64_t bar;
  const CUtensorMap* m = (VAR & 1) ? gmap : &pmap;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(160 * 72) : "memory");
    if (VAR & 2)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(tile)),
                   "l"(m), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)),
                   "l"(m), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra W;\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < 160 * 72; i += blockDim.x) acc = acc * 131 + tile[i];
  out[threadIdx.x] = acc;
}

int main(int argc, char** argv) {
  const int W = 1241, H = 376, F = 4, P = 1248;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  uint8_t* d_img; CKC(cudaMalloc(&d_img, (size_t)F * H * P));
  std::vector<uint8_t> hi((size_t)F * H * P);
  uint32_t s = 1; for (auto& v : hi) { s = s * 1664525u + 1013904223u; v = (uint8_t)(s >> 24); }
  CKC(cudaMemcpy(d_img, hi.data(), hi.size(), cudaMemcpyHostToDevice));
  CUtensorMap m2, m3;
  { cuuint64_t dims[2] = {W, H}; cuuint64_t st[1] = {P}; cuuint32_t box[2] = {160, 72}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 2d rc=%d\n", (int)r); }
  { cuuint64_t dims[3] = {W, H, F}; cuuint64_t st[2] = {P, (cuuint64_t)P * H}; cuuint32_t box[3] = {160, 72, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 3d rc=%d\n", (int)r); }
  CUtensorMap* d_maps; CKC(cudaMalloc(&d_maps, 2 * sizeof(CUtensorMap)));
  CKC(cudaMemcpy(d_maps, &m2, sizeof(m2), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_maps + 1, &m3, sizeof(m3), cudaMemcpyHostToDevice));
  uint32_t* d_out; CKC(cudaMalloc(&d_out, 256 * 4));
  std::vector<uint32_t> o(256), ref(256);
  const int x0 = 128, y0 = 64, f = 2;
  for (int var = 0; var < 4; var++) {
    const int ff = (var & 2) ? f : 0;
    for (int t = 0; t < 256; t++) { uint32_t acc = 0; for (int i = t; i < 160 * 72; i += 256) { int xx = x0 - 16 + i % 160, yy = y0 - 4 + i / 160; acc = acc * 131 + ((xx >= 0 && xx < W && yy >= 0 && yy < H) ? hi[((size_t)ff * H + yy) * P + xx] : 0); } ref[t] = acc; }
    if (var == 0) k<0><<<1, 256>>>(m2, d_maps, d_out, x0 - 16, y0 - 4, f);
    if (var == 1) k<1><<<1, 256>>>(m2, d_maps, d_out, x0 - 16, y0 - 4, f);
    if (var == 2) k<2><<<1, 256>>>(m3, d_maps + 1, d_out, x0 - 16, y0 - 4, f);
    if (var == 3) k<3><<<1, 256>>>(m3, d_maps + 1, d_out, x0 - 16, y0 - 4, f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("var %d (%s map, %dd): %s\n", var, (var & 1) ? "global" : "param", (var & 2) ? 3 : 2, cudaGetErrorString(e)); return 1; }
    CKC(cudaMemcpy(o.data(), d_out, 256 * 4, cudaMemcpyDeviceToHost));
    int bad = 0; for (int t = 0; t < 256; t++) bad += o[t] != ref[t];
    printf("var %d (%s map, %dd): ok, %d mismatches\n", var, (var & 1) ? "global" : "param", (var & 2) ? 3 : 2, bad);
  }
  return 0;
}
