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
#include "fft_regs.cuh"

namespace k9 {

constexpr int FFT2048_THREADS = 128;

// tw: W_2048^a = exp(-2*pi*i*a/2048), a in [0,2048) (forward sign; conjugated here when SIGN=+1)
template <int SIGN>
__device__ __forceinline__ void fft2048(float2 (&v)[16], float2* __restrict__ sb, const float2* __restrict__ tw) {
  const int t = threadIdx.x;
  // ---- stage 1: two radix-8 butterflies, p = t + 128e; y1[8p + j] = w_2048^(p j) * DFT8 ----
#pragma unroll
  for (int e = 0; e < 2; e++) {
    Dft<8, SIGN>::run(&v[8 * e]);
    const int p = t + 128 * e;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      float2 w = __ldg(tw + ((p * j) & 2047));
      if (SIGN > 0) w.y = -w.y;
      v[8 * e + j] = cmul(v[8 * e + j], w);
    }
  }
  __syncthreads();  // WAR: previous users of the buffer are done
  {
    const int sw = t & 15;
#pragma unroll
    for (int e = 0; e < 2; e++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int a = 8 * (t + 128 * e) + j;
        sb[a ^ sw] = v[8 * e + j];
      }
    }
  }
  __syncthreads();
  // ---- stage 2: radix 16 on n=256, s=8: thread t = q + 8p'; reads y1[t + 128 r] ----
  {
    const int sw = t >> 3;
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = sb[(t + 128 * r) ^ sw];
  }
  Dft<16, SIGN>::run(v);
  {
    const int pp = t >> 3;  // p'
#pragma unroll
    for (int j = 1; j < 16; j++) {
      float2 w = __ldg(tw + 8 * pp * j);
      if (SIGN > 0) w.y = -w.y;
      v[j] = cmul(v[j], w);
    }
  }
  __syncthreads();  // WAR on the buffer
  {
    const int q = t & 7, pp = t >> 3;
    const int base = (q + 128 * pp) ^ ((pp & 1) << 3);
#pragma unroll
    for (int j = 0; j < 16; j++) sb[base ^ (8 * j)] = v[j];  // a = q + 128 p' + 8 j (no carries), swizzled
  }
  __syncthreads();
  // ---- stage 3: radix 16 on n=16, s=128: thread t reads y2[t + 128 r], writes X[t + 128 j] ----
#pragma unroll
  for (int r = 0; r < 16; r++) v[r] = sb[(t + 128 * r) ^ ((r & 1) << 3)];
  Dft<16, SIGN>::run(v);
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
