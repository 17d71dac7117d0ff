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

// ---- complex helpers ----
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
// multiply by exp(SIGN*i*pi/2) = SIGN*i
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
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
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
  float2 t1 = cadd(b, c);
  float2 t2 = make_float2(a.x - 0.5f * t1.x, a.y - 0.5f * t1.y);
  float2 d = csub(b, c);
  float2 t3 = mul_i<SIGN>(make_float2(s60 * d.x, s60 * d.y));
  a = cadd(a, t1);
  b = cadd(t2, t3);
  c = csub(t2, t3);
}

template <int SIGN>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d);
  float2 jbmd = mul_i<SIGN>(csub(b, d));
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
  float2 a14 = cadd(v1, v4), s14 = csub(v1, v4);
  float2 a23 = cadd(v2, v3), s23 = csub(v2, v3);
  float2 r1 = make_float2(v0.x + c1 * a14.x + c2 * a23.x, v0.y + c1 * a14.y + c2 * a23.y);
  float2 r2 = make_float2(v0.x + c2 * a14.x + c1 * a23.x, v0.y + c2 * a14.y + c1 * a23.y);
  float2 i1 = mul_i<SIGN>(make_float2(s1 * s14.x + s2 * s23.x, s1 * s14.y + s2 * s23.y));
  float2 i2 = mul_i<SIGN>(make_float2(s2 * s14.x - s1 * s23.x, s2 * s14.y - s1 * s23.y));
  v0 = make_float2(v0.x + a14.x + a23.x, v0.y + a14.y + a23.y);
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
