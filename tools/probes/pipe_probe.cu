// pipe_probe.cu -- issue-rate microbenchmark for the instructions the ORB kernels are built from (sm_100a).
// Each warp runs ITER iterations of 8 independent chains of one instruction; 32 warps per SM, one CTA of 1024 threads
// per SM.  Prints thread-ops per clock per SM (128 = one warp instruction per scheduler per clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

constexpr int ITER = 2048;
typedef unsigned long long u64;

template <int OP>
__device__ __forceinline__ void step(uint32_t (&r)[8], uint32_t a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    uint32_t x = r[j];
    if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 5) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));   // native VABSDIFF4 only with add form
    if (OP == 6) x = __vmaxu2(x, a);
    if (OP == 7) x = __vimax3_s16x2(x, a, b);
    if (OP == 8) { __half2 h = __hmax2(*(__half2*)&x, *(__half2*)&a); x = *(uint32_t*)&h; }
    if (OP == 9) { __half2 h = __hadd2(*(__half2*)&x, *(__half2*)&a); x = *(uint32_t*)&h; }
    if (OP == 10) { __half2 h = __hfma2(*(__half2*)&x, *(__half2*)&a, *(__half2*)&b); x = *(uint32_t*)&h; }
    if (OP == 11) { float f = __fmul_rn(__uint_as_float(x), __uint_as_float(a)); x = __float_as_uint(f); }
    if (OP == 12) x = __dp4a(x, a, b);
    if (OP == 13) x = __popc(x) + a;
    if (OP == 14) { float f = fmaxf(__uint_as_float(x), __uint_as_float(a)); x = __float_as_uint(f); }
    if (OP == 15) x = __float2int_rz(__uint_as_float(x)) ^ a;
    if (OP == 16) x = __vabsdiffu4(x, a);
    if (OP == 17) x = __viaddmax_s16x2(x, a, b);
    if (OP == 18) { __half2 h = __hadd2_sat(*(__half2*)&x, *(__half2*)&a); x = *(uint32_t*)&h; }
    if (OP == 19) x = __hge2_mask(*(__half2*)&x, *(__half2*)&a);
    if (OP == 20) x = __dp2a_lo(x, a, b);
    if (OP == 21) { float f = __fadd_rn(__uint_as_float(x), __uint_as_float(a)); x = __float_as_uint(f); }
    r[j] = x;
  }
}
// packed f32x2 multiply: 4 chains of 64-bit pairs
__device__ __forceinline__ void step_f2(u64 (&q)[4], u64 a) {
#pragma unroll
  for (int j = 0; j < 4; j++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[j]) : "l"(a));
}
__device__ __forceinline__ void step_a2(u64 (&q)[4], u64 a) {
#pragma unroll
  for (int j = 0; j < 4; j++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[j]) : "l"(a));
}

template <int OP>
__global__ void __launch_bounds__(1024) k_probe(uint32_t a, uint32_t b, uint32_t* out, long long* cyc) {
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = threadIdx.x * 8 + j + a;
  u64 q[4];
#pragma unroll
  for (int j = 0; j < 4; j++) q[j] = ((u64)__float_as_uint(1.0f + j) << 32) | __float_as_uint(1.5f);
  const u64 aa = ((u64)__float_as_uint(1.0000001f) << 32) | __float_as_uint(0.9999999f);
  __syncthreads();
  long long t0 = clock64();
  if (OP < 100) {
#pragma unroll 4
    for (int i = 0; i < ITER; i++) step<OP>(r, a, b);
  } else if (OP == 100) {
#pragma unroll 4
    for (int i = 0; i < ITER; i++) step_f2(q, aa);
  } else if (OP == 101) {
#pragma unroll 4
    for (int i = 0; i < ITER; i++) step_a2(q, aa);
  } else if (OP == 102) {           // LOP3 + IMAD interleaved: do the two pipes dual-issue?
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[j + 1]) : "r"(a), "r"(b));
      }
    }
  } else if (OP == 103) {           // HMNMX2 (alu?) + HADD2 (fma?)
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        __half2 h = __hmax2(*(__half2*)&r[j], *(__half2*)&a); r[j] = *(uint32_t*)&h;
        __half2 g = __hadd2(*(__half2*)&r[j + 1], *(__half2*)&a); r[j + 1] = *(uint32_t*)&g;
      }
    }
  } else if (OP == 104) {           // VIMNMX.U16x2 + IMAD
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        r[j] = __vmaxu2(r[j], a);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[j + 1]) : "r"(a), "r"(b));
      }
    }
  } else if (OP == 105) {           // VIMNMX.U16x2 + LOP3 (same pipe?)
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        r[j] = __vmaxu2(r[j], a);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j + 1]) : "r"(a), "r"(b));
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s ^= r[j];
#pragma unroll
  for (int j = 0; j < 4; j++) s ^= (uint32_t)q[j] ^ (uint32_t)(q[j] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_iter, uint32_t* out, long long* cyc, int nsm) {
  k_probe<OP><<<nsm, 1024>>>(0x3c003c00u, 0x00010001u, out, cyc);
  k_probe<OP><<<nsm, 1024>>>(0x3c003c00u, 0x00010001u, out, cyc);
  cudaDeviceSynchronize();
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  double c = (double)h[nsm / 2];
  printf("%-28s %8.1f thread-ops/clk/SM  (%.2f warp-instr/clk/SM)\n", name, 1024.0 * ITER * ops_per_iter / c, 32.0 * ITER * ops_per_iter / c);
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, sizeof(uint32_t) * nsm * 1024);
  cudaMalloc(&cyc, sizeof(long long) * nsm);
  run<0>("LOP3", 8, out, cyc, nsm);
  run<1>("IADD", 8, out, cyc, nsm);
  run<2>("IMAD", 8, out, cyc, nsm);
  run<3>("PRMT", 8, out, cyc, nsm);
  run<4>("SHF", 8, out, cyc, nsm);
  run<5>("VABSDIFF4 (+acc)", 8, out, cyc, nsm);
  run<16>("__vabsdiffu4", 8, out, cyc, nsm);
  run<6>("VIMNMX.U16x2", 8, out, cyc, nsm);
  run<7>("VIMNMX3.S16x2", 8, out, cyc, nsm);
  run<17>("VIADDMNMX.S16x2", 8, out, cyc, nsm);
  run<8>("HMNMX2", 8, out, cyc, nsm);
  run<9>("HADD2", 8, out, cyc, nsm);
  run<18>("HADD2.SAT", 8, out, cyc, nsm);
  run<10>("HFMA2", 8, out, cyc, nsm);
  run<19>("HSET2 (hge2_mask)", 8, out, cyc, nsm);
  run<11>("FMUL", 8, out, cyc, nsm);
  run<21>("FADD", 8, out, cyc, nsm);
  run<100>("FMUL2 (f32x2)", 4, out, cyc, nsm);
  run<101>("FADD2 (f32x2)", 4, out, cyc, nsm);
  run<12>("IDP.4A", 8, out, cyc, nsm);
  run<20>("IDP.2A", 8, out, cyc, nsm);
  run<13>("POPC+IADD", 16, out, cyc, nsm);
  run<14>("FMNMX", 8, out, cyc, nsm);
  run<15>("F2I+LOP", 16, out, cyc, nsm);
  run<102>("LOP3 + IMAD mixed", 8, out, cyc, nsm);
  run<103>("HMNMX2 + HADD2 mixed", 8, out, cyc, nsm);
  run<104>("VIMNMX.U16x2 + IMAD mixed", 8, out, cyc, nsm);
  run<105>("VIMNMX.U16x2 + LOP3 mixed", 8, out, cyc, nsm);
  return 0;
}
