// Lean 128-point Stockham passes for the large forward transforms (N = 128 * 128 * ... : cfg5's N = 2 621 440).
//
// Same pass definition as bigfft.cuh, but written like fft2048.cuh: one CTA = 128 threads = one tile of 16 columns x 128
// points, every thread carries 16 values through BOTH register stages (the generic pass kernel idles half its threads in
// the second stage and spends ~3x the instructions on index arithmetic), 32-bit indexing, and the inter-pass twiddles
// W_N^(...) of a tile are formed once per tile in shared memory instead of once per output.
//
//   first pass  (s = 1):      stage A radix 16 (one butterfly per thread), stage B radix 8 (two per thread) so that the
//                             16 lanes of a half-warp hold 16 consecutive outputs of ONE column: 128-byte stores.
//                             Reads float2 or the int16 I/Q ring (radio.c:113-114,122 scaling; if_power numerator
//                             radio.c:123 accumulated on the way).
//   middle pass (16 | s):     stage A radix 8 (two per thread), stage B radix 16 (one per thread); lanes = 16 adjacent
//                             columns in both stages.
//
// Twiddle accuracy: W_N^e is the product of two correctly rounded table entries (as in the generic kernel); the first
// pass forms W_N^(c*jt) = W_N^(c*q) * W_N^(16*c*j) from two such products (one more rounding, ~4e-8 relative).
#pragma once
#include "bigfft.cuh"
#include "fft_regs.cuh"
#include "util.cuh"

namespace k9 {

#ifndef P128_MIN_CTAS
#define P128_MIN_CTAS 6
#endif
constexpr int P128_THREADS = 128;
constexpr int P128_S = 129;  // padded column stride of the exchange buffer (float2): lanes = columns -> distinct banks

struct P128Shared {
  float2 u[16 * P128_S];
  float2 tw[256];  // per-tile inter-pass twiddles
  float red[4];
};

__device__ __forceinline__ float2 p128_twN(const PassArgs& a, unsigned e) {
  return cmul(__ldg(a.tw_lo + (e & 1023u)), __ldg(a.tw_hi + (e >> 10)));
}

// ---- first pass: s == 1, n_cur == N. grid = (ncols/16, batch) ----
template <bool RING_S16>
__global__ void __launch_bounds__(P128_THREADS, P128_MIN_CTAS) fft_pass128_first_kernel(const PassArgs a) {
  __shared__ P128Shared sh;
  const int t = threadIdx.x;
  const int batch = blockIdx.y;
  const int col0 = blockIdx.x * 16;
  const int ncols = a.ncols;
  float2 v[16];
  float esum = 0.f;
  {
    // stage A: thread (col, p), p < 8: x[c + ncols*(p + 8r)], r < 16
    const int col = t & 15, p = t >> 4;
    const int idx0 = col0 + col + ncols * p;
    const int step = 8 * ncols;
    if (RING_S16) {
      const int cap = (int)a.ring_cap;
      int pos = (int)((a.ring_off + (long long)batch * a.ring_step) % a.ring_cap) + idx0;  // < 2*cap
      if (pos >= cap) pos -= cap;
      const short2* in = reinterpret_cast<const short2*>(a.in);
      short2 raw[16];
#pragma unroll
      for (int r = 0; r < 16; r++) {
        raw[r] = __ldg(in + pos);
        pos += step;
        if (pos >= cap) pos -= cap;
      }
      const bool stats = a.energy != nullptr;
#pragma unroll
      for (int r = 0; r < 16; r++) {
        // reference radio.c:113-114,122: (float)int16 * SCALE16, then * gain_factor
        v[r] = make_float2(((float)raw[r].x * a.scale) * a.gain, ((float)raw[r].y * a.scale) * a.gain);
        if (stats && idx0 + r * step >= a.stat_from) esum += v[r].x * v[r].x + v[r].y * v[r].y;
      }
    } else {
      const float2* in = reinterpret_cast<const float2*>(a.in) + (long long)batch * a.in_batch_stride + idx0;
#pragma unroll
      for (int r = 0; r < 16; r++) v[r] = __ldg(in + (long long)r * step);
    }
    Dft<16, -1>::run(v);
    const float2* twr = a.tw_r + 0;  // W_128^(p*j)
    float2* up = sh.u + col * P128_S + 16 * p;
    up[0] = v[0];
#pragma unroll
    for (int j = 1; j < 16; j++) up[j] = cmul(v[j], __ldg(twr + p * j));
    // per-tile twiddles: Q[col][j] = W_N^(16*c*j), j < 8 (thread t -> col = t >> 3, j = t & 7)
    {
      const int c = col0 + (t >> 3);
      sh.tw[t] = p128_twN(a, 16u * (unsigned)c * (unsigned)(t & 7));
    }
    if (RING_S16 && a.energy != nullptr) {
      esum = warp_sum(esum);
      if ((t & 31) == 0) sh.red[t >> 5] = esum;
    }
  }
  __syncthreads();
  if (RING_S16 && a.energy != nullptr && t == 0) {
    // one atomic per tile (status only: numerator of demod->sig.if_power, reference radio.c:123,143-144)
    const float e = (sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]);
    if (e != 0.f) atomicAdd(a.energy + batch, e);
  }
  {
    // stage B: thread (q, col), q = t & 15, col = (t >> 4) + 8b: u[col][q + 16r], r < 8 -> jt = q + 16j
    const int q = t & 15;
    float2* outb = a.out + (long long)batch * a.out_batch_stride;
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int col = (t >> 4) + 8 * b;
      const int c = col0 + col;
      const float2* up = sh.u + col * P128_S + q;
      float2* w = v + 8 * b;
#pragma unroll
      for (int r = 0; r < 8; r++) w[r] = up[16 * r];
      Dft<8, -1>::run(w);
      const float2 P = p128_twN(a, (unsigned)c * (unsigned)q);  // W_N^(c*q)
      float2* o = outb + (long long)c * 128 + q;
      o[0] = cmul(w[0], P);
#pragma unroll
      for (int j = 1; j < 8; j++) o[16 * j] = cmul(w[j], cmul(P, sh.tw[8 * col + j]));
    }
  }
}

// ---- middle pass: 1 < s, 16 | s, n_cur != 128. grid = (ncols/16, batch) ----
__global__ void __launch_bounds__(P128_THREADS, P128_MIN_CTAS) fft_pass128_mid_kernel(const PassArgs a) {
  __shared__ P128Shared sh;
  const int t = threadIdx.x;
  const int batch = blockIdx.y;
  const int col0 = blockIdx.x * 16;
  const int ncols = a.ncols;
  const int s = a.N / a.n_cur;
  const int pg = col0 / s, qg0 = col0 - pg * s;  // the 16 columns of a tile share pg (16 | s)
  const int col = t & 15;
  float2 v[16];
  {
    // stage A: thread (col, p), p = p0 + 8b: x[c + ncols*(p + 16r)], r < 8
    const int p0 = t >> 4;
    const float2* in = reinterpret_cast<const float2*>(a.in) + (long long)batch * a.in_batch_stride + col0 + col;
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int r = 0; r < 8; r++) v[8 * b + r] = __ldg(in + (long long)ncols * (p0 + 8 * b + 16 * r));
    // per-tile twiddles: W_N^(pg*s*jt), jt = t < 128
    sh.tw[t] = p128_twN(a, (unsigned)pg * (unsigned)s * (unsigned)t);
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int p = p0 + 8 * b;
      float2* w = v + 8 * b;
      Dft<8, -1>::run(w);
      float2* up = sh.u + col * P128_S + 8 * p;
      up[0] = w[0];
#pragma unroll
      for (int j = 1; j < 8; j++) up[j] = cmul(w[j], __ldg(a.tw_r + p * j));
    }
  }
  __syncthreads();
  {
    // stage B: thread (col, q), q = t >> 4 < 8: u[col][q + 8r], r < 16 -> jt = q + 8j
    const int q = t >> 4;
    const float2* up = sh.u + col * P128_S + q;
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = up[8 * r];
    Dft<16, -1>::run(v);
    float2* o = a.out + (long long)batch * a.out_batch_stride + qg0 + col + (long long)s * (128 * pg + q);
    const long long ostep = 8ll * s;
#pragma unroll
    for (int j = 0; j < 16; j++) o[ostep * j] = cmul(v[j], sh.tw[q + 8 * j]);
  }
}

// ---- last pass, R = 160 = 16 x 10 (no inter-pass twiddle, s = N/160, pg = 0): 160 threads, tile = 16 columns ----
// stage A radix 16 (one butterfly per thread), stage B radix 10 (256 butterflies over 160 threads: two rounds, the
// second 60 % full); lanes = 16 adjacent columns in both stages, so loads and stores are 128-byte rows.
#ifndef P160_MIN_CTAS
#define P160_MIN_CTAS 4
#endif
constexpr int P160_THREADS = 160;
constexpr int P160_S = 161;
__global__ void __launch_bounds__(P160_THREADS, P160_MIN_CTAS) fft_pass160_last_kernel(const PassArgs a) {
  __shared__ float2 u[16 * P160_S];
  const int t = threadIdx.x;
  const int batch = blockIdx.y;
  const int col0 = blockIdx.x * 16;
  const int ncols = a.ncols;  // == s
  const int col = t & 15;
  float2 v[16];
  {
    const int p = t >> 4;  // < 10: x[c + ncols*(p + 10r)], r < 16
    const float2* in = reinterpret_cast<const float2*>(a.in) + (long long)batch * a.in_batch_stride + col0 + col;
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = __ldg(in + (long long)ncols * (p + 10 * r));
    Dft<16, -1>::run(v);
    float2* up = u + col * P160_S + 16 * p;
    up[0] = v[0];
#pragma unroll
    for (int j = 1; j < 16; j++) up[j] = cmul(v[j], __ldg(a.tw_r + p * j));
  }
  __syncthreads();
  float2* outb = a.out + (long long)batch * a.out_batch_stride + col0 + col;
#pragma unroll
  for (int round = 0; round < 2; round++) {
    const int q = (t >> 4) + 10 * round;  // < 16: u[col][q + 16r], r < 10 -> jt = q + 16j
    if (q < 16) {
      const float2* up = u + col * P160_S + q;
#pragma unroll
      for (int r = 0; r < 10; r++) v[r] = up[16 * r];
      Dft<10, -1>::run(v);
      float2* o = outb + (long long)ncols * q;
      if (a.route_mask == nullptr) {
#pragma unroll
        for (int j = 0; j < 10; j++) o[(long long)ncols * 16 * j] = v[j];
      } else {
        // fused exchange: the 16 lanes of a half-warp hold one 128-byte row; it is stored into the spectrum buffer of
        // every rank that reads it (peer memory over NVLink, own memory if this rank reads it) and nowhere else
#pragma unroll
        for (int j = 0; j < 10; j++) {
          unsigned m = a.route_mask[(col0 + ncols * (q + 16 * j)) >> 4];
          float2* dst = o + (long long)ncols * 16 * j;
          while (m) {
            const int r = __ffs(m) - 1;
            m &= m - 1;
            *reinterpret_cast<float2*>(reinterpret_cast<char*>(dst) + a.route_delta[r]) = v[j];
          }
        }
      }
    }
  }
  if (a.route_mask != nullptr) __threadfence_system();  // the peer stores are performed before the kernel ends
}

// Shared-memory carve-out of the lean pass kernels. CTAs of kernels with different carve-outs cannot be resident on one
// SM at the same time, so a forward FFT that is meant to run BESIDE the channel kernels (the FFT of batch k+1 under the
// channel kernels of batch k; at 1024 channels per GPU the FM kernel leaves more than half of every SM free) has to ask for
// the same carve-out they use, or it waits until they have drained (measured at 8 GPUs: no overlap at all without this).
static void set_pass128_carveout(int pct) {
  static int configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (configured[dev] == pct + 1000) return;
  cudaFuncSetAttribute(fft_pass128_first_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(fft_pass128_first_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(fft_pass128_mid_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(fft_pass160_last_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  configured[dev] = pct + 1000;
}

// returns cudaErrorNotSupported when the generic kernels should run this pass
static cudaError_t launch_pass128(int R1, int R2, const PassArgs& a, int batch, int sign, cudaStream_t st) {
  if (R1 == 10 && R2 == 16 && sign < 0 && a.n_cur == 160 && a.in_mode == IN_C32 && a.ncols % 16 == 0 &&
      (long long)a.N < (1ll << 30)) {
    fft_pass160_last_kernel<<<dim3(a.ncols / 16, batch), P160_THREADS, 0, st>>>(a);
    return cudaGetLastError();
  }
  if (!(R1 == 8 && R2 == 16) || sign >= 0 || a.n_cur == 128) return cudaErrorNotSupported;
  if (a.ncols % 16 != 0 || a.ring_cap >= (1ll << 30) || (long long)a.N >= (1ll << 30)) return cudaErrorNotSupported;
  const int s = a.N / a.n_cur;
  dim3 grid(a.ncols / 16, batch);
  if (s == 1) {
    if (a.in_mode == IN_RING_S16)
      fft_pass128_first_kernel<true><<<grid, P128_THREADS, 0, st>>>(a);
    else if (a.in_mode == IN_C32)
      fft_pass128_first_kernel<false><<<grid, P128_THREADS, 0, st>>>(a);
    else
      return cudaErrorNotSupported;
  } else {
    if (s % 16 != 0 || a.in_mode != IN_C32) return cudaErrorNotSupported;
    fft_pass128_mid_kernel<<<grid, P128_THREADS, 0, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace k9
