// orb_math.cuh -- bit-exact device twins of the libm calls on the reference's ORB path.
//
// The reference's CPU path calls std::atan2(float,float) (src/orb_cpu.cpp:178), std::cos / std::sin
// (float) (:217-218) and std::lround(float) (:228-232) from glibc 2.39.  CUDA's atan2f/sinf/cosf
// differ from glibc's in the last ulp for ~1-16 % of inputs, which would flip rotated BRIEF
// sample points.  The routines below perform the SAME IEEE-754 operation sequence as glibc 2.39's
// x86-64 implementations (float fdlibm atanf/atan2f; double-polynomial sinf/cosf), with every
// multiply/add issued through a round-to-nearest intrinsic so that nvcc cannot contract them into
// FMAs.  tools/libm_replica_check_*.c verify the operation sequences against glibc on the host
// (0 mismatches over 20-30 M samples); tests/test_gpu_math.py verifies the device versions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orbm {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// std::lround(float) for |v| < 2^22: round half away from zero.
// v + copysign(pred(0.5), v) truncated toward zero is exact for every float in that range
// (exhaustively checked on [0,64) in tools/libm_replica_check_atan2f.c).
__device__ __forceinline__ int lround_f(float v) {
  float h = __int_as_float(0x3effffff | (__float_as_int(v) & 0x80000000));
  return __float2int_rz(__fadd_rn(v, h));
}

// glibc 2.39 sysdeps/ieee754/flt-32/s_atanf.c (fdlibm), operation for operation
__device__ __forceinline__ float atanf_fdlibm(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  int hx = __float_as_int(x), ix = hx & 0x7fffffff, id;
  if (ix >= 0x4c000000) {   // |x| >= 2^25
    return hx > 0 ? fadd(atanhi[3], atanlo[3]) : fsub(-atanhi[3], atanlo[3]);
  }
  if (ix < 0x3ee00000) {    // |x| < 0.4375
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x)); }
      else                 { id = 1; x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x))); }
      else                 { id = 3; x = fdiv(-1.0f, x); }
    }
  }
  float z = fmul(x, x), w = fmul(z, z);
  float s1 = fmul(z, fadd(aT0, fmul(w, fadd(aT2, fmul(w, fadd(aT4, fmul(w, fadd(aT6, fmul(w, fadd(aT8, fmul(w, aT10)))))))))));
  float s2 = fmul(w, fadd(aT1, fmul(w, fadd(aT3, fmul(w, fadd(aT5, fmul(w, fadd(aT7, fmul(w, aT9)))))))));
  if (id < 0) return fsub(x, fmul(x, fadd(s1, s2)));
  float hi = id == 0 ? atanhi[0] : id == 1 ? atanhi[1] : id == 2 ? atanhi[2] : atanhi[3];
  float lo = id == 0 ? atanlo[0] : id == 1 ? atanlo[1] : id == 2 ? atanlo[2] : atanlo[3];
  z = fsub(hi, fsub(fsub(fmul(x, fadd(s1, s2)), lo), x));
  return hx < 0 ? -z : z;
}

// glibc 2.39 sysdeps/ieee754/flt-32/e_atan2f.c (fdlibm); finite inputs only (moments are integers)
__device__ __forceinline__ float atan2f_glibc(float y, float x) {
  const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f, pi_o_2 = 1.5707963705e+00f, tiny = 1.0e-30f;
  int hx = __float_as_int(x), hy = __float_as_int(y);
  int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (hx == 0x3f800000) return atanf_fdlibm(y);   // x == 1.0
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m < 2) return y;
    return m == 2 ? fadd(pi, tiny) : fsub(-pi, tiny);
  }
  if (ix == 0) return hy < 0 ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_fdlibm(fabsf(fdiv(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return fsub(pi, fsub(z, pi_lo));
    default: return fsub(fsub(z, pi_lo), pi);
  }
}

// glibc 2.39 sysdeps/ieee754/flt-32/s_sincosf.h sinf_poly + reduce_fast (non-TOINT_INTRINSICS form)
__device__ __forceinline__ float sincos_poly(double x, double x2, bool neg_cos, int n) {
  const double c0 = neg_cos ? -0x1p0 : 0x1p0;
  const double c1 = neg_cos ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
  const double c2 = neg_cos ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
  const double c3 = neg_cos ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
  const double c4 = neg_cos ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
  const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    double x3 = dmul(x, x2);
    double s1 = dadd(s2c, dmul(x2, s3c));
    double x7 = dmul(x3, x2);
    double s = dadd(x, dmul(x3, s1c));
    return __double2float_rn(dadd(s, dmul(x7, s1)));
  } else {
    double x4 = dmul(x2, x2);
    double cc2 = dadd(c3, dmul(x2, c4));
    double cc1 = dadd(c0, dmul(x2, c1));
    double x6 = dmul(x4, x2);
    double c = dadd(cc1, dmul(x4, c2));
    return __double2float_rn(dadd(c, dmul(x6, cc2)));
  }
}

// valid for |y| < 120 (angles here are in [-pi, pi])
template <bool COS>
__device__ __forceinline__ float sincosf_glibc(float y) {
  double x = (double)y;
  unsigned top = ((unsigned)__float_as_int(y) >> 20) & 0x7ff;
  const unsigned top_pio4 = (0x3f490fdbu >> 20) & 0x7ff;    // abstop12(pi/4)
  const unsigned top_tiny = (0x39800000u >> 20) & 0x7ff;    // abstop12(0x1p-12f)
  if (top < top_pio4) {
    if (top < top_tiny) return COS ? 1.0f : y;
    return sincos_poly(x, dmul(x, x), false, COS ? 1 : 0);
  }
  double r = dmul(x, 0x1.45F306DC9C883p+23);
  int n = (__double2int_rz(r) + 0x800000) >> 24;
  x = dsub(x, dmul((double)n, 0x1.921FB54442D18p0));
  double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return sincos_poly(dmul(x, sgn), dmul(x, x), (n & 2) != 0, COS ? (n ^ 1) : n);
}
__device__ __forceinline__ float cosf_glibc(float a) { return sincosf_glibc<true>(a); }
__device__ __forceinline__ float sinf_glibc(float a) { return sincosf_glibc<false>(a); }

}  // namespace orbm
