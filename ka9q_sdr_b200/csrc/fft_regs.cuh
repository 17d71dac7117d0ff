// Register-resident small DFTs with compile-time twiddles (sm_100a, fp32).
//
// Building blocks of every FFT in this library (the N-point forward transform that replaces
// fftwf_execute(fwd_plan) at reference filter.c:151 and the N_dec-point inverse that replaces
// filter.c:250). Transform definition = FFTW's: X[k] = sum_n x[n] exp(SIGN*2*pi*i*n*k/R), unnormalised;
// SIGN=-1 forward, +1 backward.
//
// dft<R,SIGN>(v) transforms float2 v[R] in place, natural order in, natural order out, fully unrolled so
// v[] stays in registers. R in {2,3,4,5,8,10,16,20,25}: radix 2/3/4/5 kernels plus Cooley-Tukey products whose
// twiddles are constexpr-evaluated (no table, no SFU).
#pragma once
#include <cuda_runtime.h>

namespace k9 {

// ---- constexpr trig (compile-time only) ----
constexpr double cx_pi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double cx_sin_taylor(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int k = 1; k < 12; k++) {
    term *= -x2 / ((2 * k) * (2 * k + 1));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double cx_cos_taylor(double x) {  // |x| <= pi/4
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int k = 1; k < 12; k++) {
    term *= -x2 / ((2 * k - 1) * (2 * k));
    sum += term;
  }
  return sum;
}
// cos/sin of 2*pi*num/den with exact octant reduction on the integers
__host__ __device__ constexpr double cx_cos2pi(long long num, long long den) {
  num %= den;
  if (num < 0) num += den;
  long long oct = (num * 8) / den;
  long long rem = num * 8 - oct * den;  // angle within octant = (pi/4)*rem/den
  double a = (oct & 1) ? (cx_pi / 4) * (double)(den - rem) / (double)den : (cx_pi / 4) * (double)rem / (double)den;
  double c = cx_cos_taylor(a), s = cx_sin_taylor(a);
  switch (oct) {
    case 0: return c;
    case 1: return s;
    case 2: return -s;
    case 3: return -c;
    case 4: return -c;
    case 5: return -s;
    case 6: return s;
    default: return c;
  }
}
__host__ __device__ constexpr double cx_sin2pi(long long num, long long den) {
  num %= den;
  if (num < 0) num += den;
  long long oct = (num * 8) / den;
  long long rem = num * 8 - oct * den;
  double a = (oct & 1) ? (cx_pi / 4) * (double)(den - rem) / (double)den : (cx_pi / 4) * (double)rem / (double)den;
  double c = cx_cos_taylor(a), s = cx_sin_taylor(a);
  switch (oct) {
    case 0: return s;
    case 1: return c;
    case 2: return c;
    case 3: return s;
    case 4: return -s;
    case 5: return -c;
    case 6: return -c;
    default: return -s;
  }
}

// ---- complex helpers on packed fp32x2 (sm_100a FADD2 / FMUL2 / FFMA2) ----
// A complex float is one 64-bit register pair; Blackwell's packed fp32 instructions add/multiply both halves in one
// issue slot, and their operand modifiers swap halves, negate one half or broadcast a scalar for free. So a complex
// add is ONE instruction (not two), a complex multiply TWO (not four), and multiplying by +-i costs nothing. The
// butterflies are ~half of what these kernels execute and the kernels are issue bound, so this is the single largest
// instruction-count lever. The pack/unpack moves below are register renaming only (ptxas allocates aligned pairs).
typedef unsigned long long k9_u64;
__device__ __forceinline__ k9_u64 pk2(float x, float y) {
  k9_u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 upk2(k9_u64 v) {
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(v));
  return d;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  k9_u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
  return upk2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  k9_u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
  return upk2(r);
}
// a * s + b with a real scalar s
__device__ __forceinline__ float2 cfma_s(float2 a, float s, float2 b) {
  k9_u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(s, s)), "l"(pk2(b.x, b.y)));
  return upk2(r);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) {
  k9_u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(s, s)));
  return upk2(r);
}
// a * (bx + i*by): (a.x bx - a.y by, a.y bx + a.x by) = a*bx + (-a.y, a.x)*by
__device__ __forceinline__ float2 cmul_xy(float2 a, float bx, float by) {
  k9_u64 t, r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk2(a.x, a.y)), "l"(pk2(bx, bx)));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(-a.y, a.x)), "l"(pk2(by, by)), "l"(t));
  return upk2(r);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return cmul_xy(a, b.x, b.y); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return cmul_xy(a, b.x, -b.y); }  // a * conj(b)
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by exp(SIGN*i*pi/2) = SIGN*i (folds into the consumer's operand modifiers)
template <int SIGN>
__device__ __forceinline__ float2 mul_i(float2 a) {
  return SIGN > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by the compile-time constant exp(SIGN*2*pi*i*K/N)
template <int SIGN, int K, int N>
__device__ __forceinline__ float2 mul_tw(float2 a) {
  constexpr int k = ((K % N) + N) % N;
  if constexpr (k == 0) {
    return a;
  } else if constexpr (4 * k == N) {
    return mul_i<SIGN>(a);
  } else if constexpr (2 * k == N) {
    return make_float2(-a.x, -a.y);
  } else if constexpr (4 * k == 3 * N) {
    return mul_i<-SIGN>(a);
  } else {
    constexpr float c = (float)cx_cos2pi(k, N);
    constexpr float s = (float)(SIGN * cx_sin2pi(k, N));
    return cmul_xy(a, c, s);
  }
}

// ---- radix kernels ----
template <int SIGN>
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <int SIGN>
__device__ __forceinline__ void dft3(float2& a, float2& b, float2& c) {
  constexpr float s60 = (float)(0.86602540378443864676);
  const float2 t1 = cadd(b, c);
  const float2 t2 = cfma_s(t1, -0.5f, a);
  const float2 t3 = mul_i<SIGN>(cscale(csub(b, c), s60));
  a = cadd(a, t1);
  b = cadd(t2, t3);
  c = csub(t2, t3);
}

template <int SIGN>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d);
  const float2 jbmd = mul_i<SIGN>(csub(b, d));
  a = cadd(apc, bpd);
  b = cadd(amc, jbmd);
  c = csub(apc, bpd);
  d = csub(amc, jbmd);
}

template <int SIGN>
__device__ __forceinline__ void dft5(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4) {
  constexpr float c1 = (float)0.30901699437494742410;   // cos(2pi/5)
  constexpr float c2 = (float)-0.80901699437494742410;  // cos(4pi/5)
  constexpr float s1 = (float)0.95105651629515357212;   // sin(2pi/5)
  constexpr float s2 = (float)0.58778525229247312917;   // sin(4pi/5)
  const float2 a14 = cadd(v1, v4), s14 = csub(v1, v4);
  const float2 a23 = cadd(v2, v3), s23 = csub(v2, v3);
  const float2 r1 = cfma_s(a23, c2, cfma_s(a14, c1, v0));
  const float2 r2 = cfma_s(a23, c1, cfma_s(a14, c2, v0));
  const float2 i1 = mul_i<SIGN>(cfma_s(s23, s2, cscale(s14, s1)));
  const float2 i2 = mul_i<SIGN>(cfma_s(s23, -s1, cscale(s14, s2)));
  v0 = cadd(v0, cadd(a14, a23));
  v1 = cadd(r1, i1);
  v4 = csub(r1, i1);
  v2 = cadd(r2, i2);
  v3 = csub(r2, i2);
}

template <int R, int SIGN>
struct Dft;

template <int SIGN>
struct Dft<1, SIGN> {
  __device__ static __forceinline__ void run(float2*) {}
};
template <int SIGN>
struct Dft<2, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { dft2<SIGN>(v[0], v[1]); }
};
template <int SIGN>
struct Dft<3, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { dft3<SIGN>(v[0], v[1], v[2]); }
};
template <int SIGN>
struct Dft<4, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { dft4<SIGN>(v[0], v[1], v[2], v[3]); }
};
template <int SIGN>
struct Dft<5, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { dft5<SIGN>(v[0], v[1], v[2], v[3], v[4]); }
};

// Cooley-Tukey product R = A*B in registers:
//   n = A*n2 + n1, k = B*k1 + k2:  X[B k1 + k2] = sum_n1 W_A^{n1 k1} * ( W_R^{n1 k2} * sum_n2 x[A n2 + n1] W_B^{n2 k2} )
template <int A, int B, int SIGN>
struct DftCT {
  template <int N1, int K2>
  __device__ static __forceinline__ void tw_row(float2 (&z)[A][B]) {
    if constexpr (K2 < B) {
      z[N1][K2] = mul_tw<SIGN, N1 * K2, A * B>(z[N1][K2]);
      tw_row<N1, K2 + 1>(z);
    }
  }
  template <int N1>
  __device__ static __forceinline__ void tw_all(float2 (&z)[A][B]) {
    if constexpr (N1 < A) {
      tw_row<N1, 0>(z);
      tw_all<N1 + 1>(z);
    }
  }
  __device__ static __forceinline__ void run(float2* v) {
    float2 z[A][B];
#pragma unroll
    for (int n1 = 0; n1 < A; n1++) {
#pragma unroll
      for (int n2 = 0; n2 < B; n2++) z[n1][n2] = v[A * n2 + n1];
      Dft<B, SIGN>::run(z[n1]);
    }
    tw_all<0>(z);
#pragma unroll
    for (int k2 = 0; k2 < B; k2++) {
      float2 c[A];
#pragma unroll
      for (int n1 = 0; n1 < A; n1++) c[n1] = z[n1][k2];
      Dft<A, SIGN>::run(c);
#pragma unroll
      for (int k1 = 0; k1 < A; k1++) v[B * k1 + k2] = c[k1];
    }
  }
};

template <int SIGN>
struct Dft<8, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<2, 4, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<16, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<4, 4, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<10, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<2, 5, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<20, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<4, 5, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<25, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<5, 5, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<6, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<2, 3, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<12, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<4, 3, SIGN>::run(v); }
};
template <int SIGN>
struct Dft<15, SIGN> {
  __device__ static __forceinline__ void run(float2* v) { DftCT<3, 5, SIGN>::run(v); }
};

}  // namespace k9
