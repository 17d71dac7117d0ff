// 2048-point complex FFT for one 128-thread CTA: three register stages (8 x 16 x 16, Stockham DIF) with exactly two
// shared-memory exchanges. This is the per-channel inverse transform that replaces fftwf_execute(slave->rev_plan)
// (reference filter.c:250) for N_dec = 2048 (the reference default geometry L=3840, M=4353, decimate 4:
// main.c:113-114, filter.c:513-515), and the forward/backward pair of the FM audio (de-emphasis) filter
// (fm.c:39-43,162-171).
//
// Data distribution (t = threadIdx.x in [0,128)):
//   input : v[8e + r] = x[t + 128e + 256r],  e in {0,1}, r in [0,8)
//   output: v[j]      = X[t + 128j],         j in [0,16)
// so the output of one call is, up to the register permutation j = e + 2r, exactly the input layout of the next —
// forward FFT -> pointwise multiply -> inverse FFT needs no exchange in between.
//
// Shared memory: one float2[2048] buffer, no padding. Both exchanges are bank-conflict free for 64-bit accesses
// (16 lanes per wavefront) through XOR swizzles of the low index bits:
//   exchange 1: a -> a ^ ((a >> 3) & 15)      exchange 2: a -> a ^ (((a >> 7) & 1) << 3)
//
// The body is deliberately instantiated ONCE per kernel (callers loop over their transforms and use
// conj(FFT(conj x)) for the opposite sign): the unrolled butterflies are ~25 KB of SASS and several copies thrash the
// instruction cache (measured: stall_no_instruction was the top stall with four inlined copies).
#pragma once
#include <math.h>
#include "fft_regs.cuh"

namespace k9 {

constexpr int FFT2048_THREADS = 128;
#ifndef FFT2048_TW_EARLY
#define FFT2048_TW_EARLY 0
#endif
#ifndef FFT2048_TW_HALF
#define FFT2048_TW_HALF 1
#endif

// Twiddle table layout (FFT2048_TW_FLOAT2 float2 entries, built by fft2048_fill_twiddles on the host), laid out so that
// the lanes of a warp read CONSECUTIVE addresses (the L1 data pipe was the kernel's limiter, and a row per thread costs
// one wavefront per lane: ncu round 1, 2560 of 7230 wavefronts per FM pair-block were twiddle rows):
//   stage 1, butterfly p in [0,256), twiddles W_2048^(p*j), j = 1..7 (j = 0 is 1 and is not stored):
//     float4 A[jj][p], jj in [0,3): (W^(p(2jj+1)), W^(p(2jj+2)))      at float2 offset 2*(256*jj + p)
//     float2 B[p]:                   W^(7p)                           at float2 offset 1536 + p
//   stage 2, butterfly p' = t >> 3 in [0,16), twiddles W_256^(p'*j), j = 0..15: 2 KB that every thread of every CTA reads
//     in full 128-byte rows; they are copied ONCE per CTA into shared memory (fft2048_stage_tw2) and read from there:
//     float4 C[jj][p']: (W^(p'*2jj), W^(p'*(2jj+1)))                  at float2 offset 1792 + 2*(16 jj + p')
//     the four p' of a warp sit side by side, so a warp's 128-bit shared load covers one 64-byte segment (one wavefront;
//     the same load from global memory returns 32 x 16 bytes through the L1 data stage = 4 wavefronts, and these rows
//     were 512 of the FM kernel's 5640 L1 wavefronts per pair-block after the stage-1 change)
// Forward sign (exp(-2*pi*i...)); conjugated on use when SIGN=+1. Same values and the same arithmetic as a row-per-
// thread table: results are bit-identical.
constexpr int FFT2048_TW_FLOAT2 = 1792 + 256;
constexpr int FFT2048_TW_B = 1536, FFT2048_TW_C = 1792;

template <int SIGN>
__device__ __forceinline__ float2 tw_mul(float2 a, float wx, float wy) {
  // a * (wx + i*SIGN'*wy) where the table holds the forward twiddle: conjugate it for the backward transform
  return cmul_xy(a, wx, SIGN > 0 ? -wy : wy);
}

constexpr int FFT2048_TW2_FLOAT4 = 128;  // shared-memory copy of the stage-2 table

// once per CTA (128 threads), before the first transform; the caller's next barrier publishes it
__device__ __forceinline__ void fft2048_stage_tw2(float4* __restrict__ tw2s, const float2* __restrict__ tw) {
  tw2s[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(tw + FFT2048_TW_C) + threadIdx.x);
}

template <int SIGN>
__device__ __forceinline__ void fft2048(float2 (&v)[16], float2* __restrict__ sb, const float2* __restrict__ tw,
                                        const float4* __restrict__ tw2s, bool war_sync = true) {
  const int t = threadIdx.x;
  // ---- stage 1: two radix-8 butterflies, p = t + 128e; y1[8p + j] = w_2048^(p j) * DFT8 ----
#if FFT2048_TW_HALF
  // Half the stage-1 twiddle loads: W_2048^((t+128) j) = W_2048^(t j) * W_16^j, so the second butterfly's outputs are first
  // turned by the compile-time constants W_16^j and then share the first butterfly's seven table entries (one more
  // rounding on those eight values; 56 fewer L1 wavefronts per transform).
  {
    Dft<8, SIGN>::run(&v[0]);
    Dft<8, SIGN>::run(&v[8]);
    const float4* A = reinterpret_cast<const float4*>(tw) + t;
    const float h = 0.70710678118654752f, c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
    // W_16^j = exp(-2 pi i j / 16) for the forward transform, conjugated for the backward one (tw_mul does that)
    v[9] = tw_mul<SIGN>(v[9], c1, -s1);
    v[10] = tw_mul<SIGN>(v[10], h, -h);
    v[11] = tw_mul<SIGN>(v[11], s1, -c1);
    v[12] = tw_mul<SIGN>(v[12], 0.f, -1.f);
    v[13] = tw_mul<SIGN>(v[13], -s1, -c1);
    v[14] = tw_mul<SIGN>(v[14], -h, -h);
    v[15] = tw_mul<SIGN>(v[15], -c1, -s1);
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
      const float4 w = __ldg(A + 256 * jj);
#pragma unroll
      for (int e = 0; e < 2; e++) {
        v[8 * e + 2 * jj + 1] = tw_mul<SIGN>(v[8 * e + 2 * jj + 1], w.x, w.y);
        v[8 * e + 2 * jj + 2] = tw_mul<SIGN>(v[8 * e + 2 * jj + 2], w.z, w.w);
      }
    }
    const float2 w7 = __ldg(tw + FFT2048_TW_B + t);
    v[7] = tw_mul<SIGN>(v[7], w7.x, w7.y);
    v[15] = tw_mul<SIGN>(v[15], w7.x, w7.y);
  }
#else
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const float4* A = reinterpret_cast<const float4*>(tw) + (t + 128 * e);
#if FFT2048_TW_EARLY
    // twiddle loads issued ahead of the butterfly so their latency hides under it (costs 14 registers)
    float4 w[3];
#pragma unroll
    for (int jj = 0; jj < 3; jj++) w[jj] = __ldg(A + 256 * jj);
    const float2 w7 = __ldg(tw + FFT2048_TW_B + t + 128 * e);
    Dft<8, SIGN>::run(&v[8 * e]);
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
      v[8 * e + 2 * jj + 1] = tw_mul<SIGN>(v[8 * e + 2 * jj + 1], w[jj].x, w[jj].y);
      v[8 * e + 2 * jj + 2] = tw_mul<SIGN>(v[8 * e + 2 * jj + 2], w[jj].z, w[jj].w);
    }
#else
    Dft<8, SIGN>::run(&v[8 * e]);
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
      const float4 w = __ldg(A + 256 * jj);
      v[8 * e + 2 * jj + 1] = tw_mul<SIGN>(v[8 * e + 2 * jj + 1], w.x, w.y);
      v[8 * e + 2 * jj + 2] = tw_mul<SIGN>(v[8 * e + 2 * jj + 2], w.z, w.w);
    }
    const float2 w7 = __ldg(tw + FFT2048_TW_B + t + 128 * e);
#endif
    v[8 * e + 7] = tw_mul<SIGN>(v[8 * e + 7], w7.x, w7.y);
  }
#endif
  if (war_sync) __syncthreads();  // WAR: previous users of the buffer are done (CTA-uniform flag)
  {
    // a = 8p + j, swizzled a ^ (t & 15): only the low nibble changes, so per j one XOR on a 3-bit value
    const int sw = t & 15;
    float2* wp = sb + ((8 * t) & ~15) + ((((t & 1) << 3)) ^ (sw & 8));
    const int s7 = sw & 7;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int off = j ^ s7;
      wp[off] = v[j];
      wp[off + 1024] = v[8 + j];
    }
  }
  __syncthreads();
  // ---- stage 2: radix 16 on n=256, s=8: thread t = q + 8p'; reads y1[t + 128 r] ----
  {
    const float2* rp = sb + (t ^ (t >> 3));
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = rp[128 * r];
  }
  Dft<16, SIGN>::run(v);
  {
    const float4* row = tw2s + (t >> 3);
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const float4 w = row[16 * jj];
      if (jj > 0) v[2 * jj] = tw_mul<SIGN>(v[2 * jj], w.x, w.y);
      v[2 * jj + 1] = tw_mul<SIGN>(v[2 * jj + 1], w.z, w.w);
    }
  }
  __syncthreads();  // WAR on the buffer
  {
    // a = q + 128 p' + 8 j (no carries), swizzled by bit 3 when p' is odd: a' = (q + 128 p') + 8*(j ^ (p'&1))
    const int pp = t >> 3;
    float2* wp = sb + (t & 7) + 128 * pp;
    const int sbit = pp & 1;
#pragma unroll
    for (int j = 0; j < 16; j++) wp[8 * (j ^ sbit)] = v[j];
  }
  __syncthreads();
  // ---- stage 3: radix 16 on n=16, s=128: thread t reads y2[t + 128 r], writes X[t + 128 j] ----
  {
    const float2* rp0 = sb + t;        // even r: no swizzle
    const float2* rp1 = sb + (t ^ 8);  // odd r: bit 3 flipped
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
      v[r] = rp0[128 * r];
      v[r + 1] = rp1[128 * (r + 1)];
    }
  }
  Dft<16, SIGN>::run(v);
}

// Host-side table builder (double precision, rounded once)
inline void fft2048_fill_twiddles(float2* tw) {
  const double pi = 3.14159265358979323846;
  auto w = [&](int num, int den) {
    const double ang = -2.0 * pi * (double)(num % den) / (double)den;
    return make_float2((float)cos(ang), (float)sin(ang));
  };
  for (int p = 0; p < 256; p++) {
    for (int jj = 0; jj < 3; jj++) {
      tw[2 * (256 * jj + p)] = w(p * (2 * jj + 1), 2048);
      tw[2 * (256 * jj + p) + 1] = w(p * (2 * jj + 2), 2048);
    }
    tw[FFT2048_TW_B + p] = w(p * 7, 2048);
  }
  for (int pp = 0; pp < 16; pp++)
    for (int jj = 0; jj < 8; jj++) {
      const int idx = FFT2048_TW_C + 2 * (16 * jj + pp);
      tw[idx] = w(pp * 2 * jj, 256);
      tw[idx + 1] = w(pp * (2 * jj + 1), 256);
    }
}

// Register permutation: output layout (index j) -> input layout (index 8e + r) with j = e + 2r.
__device__ __forceinline__ void fft2048_out_to_in(float2 (&v)[16]) {
  float2 u[16];
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int r = 0; r < 8; r++) u[8 * e + r] = v[e + 2 * r];
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = u[i];
}

}  // namespace k9
