// PL-tone analyser (reference fm.c:189-285, pltask): a second slave of the FM demodulator's audio master filter —
// REAL -> REAL, decimate 32 (48 kHz -> 1500 Hz), 300 Hz low-pass, Kaiser beta 2 — feeding a 16384-sample ring (10.9 s);
// every 512 new samples (0.34 s) the ring's 16384-point real transform is searched for the strongest bin, which is
// reported as demod->sig.plfreq if it holds more than 1 % of the energy and lies in 67..255 Hz. Status only.
//
// Included by chan_kernels.cu. Optional (ka9q_stream_enable_pl): fm_kernel then also exports bins 0..32 and 2016..2047 of
// the pair's audio transform Z = FFT(audA + j audB), which it has in registers anyway; pl_kernel (one CTA per FM channel,
// blocks in order) un-mixes the channel's real spectrum A[k] = (Z[k] + conj(Z[N-k]))/2 (B likewise), applies the slave
// response, runs the 64-point c2r by direct summation (30 kept outputs x 33 bins), appends to the ring and, when due,
// evaluates the 16384-point transform as X[k] = sum_{r<8} W^(rk) F_r[k mod 2048] with the CTA's own fft2048.
#pragma once

namespace k9 {

constexpr int PL_FFT = 1 << 14;   // fm.c:225: (1 << 19) / PL_decimate
constexpr int PL_DEC = 32;        // fm.c:201
constexpr int PL_BINS = NDEC / PL_DEC / 2 + 1;  // 33 bins of the master spectrum feed the slave
constexpr int PL_SPEC = 65;       // exported per pair-block: Z[0..32], Z[2047], Z[2046] ... Z[2016]

struct PlShared {
  float2 buf[NDEC];
  float4 tw2[FFT2048_TW2_FLOAT4];
  float2 acc[PL_FFT / 2];   // X[0..8191]
  float2 y[PL_BINS];
  float2 cs[64];            // exp(+j 2 pi n / 64)
  float red[16];
  float scal[4];
  int iscal[4];
};

__global__ void __launch_bounds__(FFT2048_THREADS, 2) pl_kernel(const ChanLaunch a) {
  extern __shared__ __align__(16) unsigned char pl_raw[];
  PlShared& sh = *reinterpret_cast<PlShared*>(pl_raw);
  const int t = threadIdx.x;
  const PlWork wk = a.pl_work[blockIdx.x];
  PlState S = a.pl_state[blockIdx.x];
  float* ring = a.pl_ring + (long long)blockIdx.x * PL_FFT;
  const int olen = a.olen;
  const int pl_l = olen / PL_DEC;  // PL_L (fm.c:204): 30
  fft2048_stage_tw2(sh.tw2, a.tw2048);
  if (t < 64) {
    float sn, cs;
    sincospif(2.0f * (float)t / 64.0f, &sn, &cs);
    sh.cs[t] = make_float2(cs, sn);
  }
  __syncthreads();
  float2 v[16];
#pragma unroll 1
  for (int b = 0; b < a.nblocks; b++) {
    const float2* Z = a.pl_spec + ((long long)b * a.pl_npairs + wk.pair) * PL_SPEC;
    // ---- slave filter: Y[k] = R[k] * A[k], k = 0..32 (filter.c:206-208), A / B un-mixed from the pair's transform
    if (t < PL_BINS) {
      const float2 zp = Z[t];
      const float2 zm = t == 0 ? Z[0] : Z[PL_BINS + t - 1];  // Z[N - k]
      float2 s;
      if (wk.half == 0)
        s = make_float2(0.5f * (zp.x + zm.x), 0.5f * (zp.y - zm.y));  // (Z[k] + conj(Z[N-k])) / 2
      else
        s = make_float2(0.5f * (zp.y + zm.y), 0.5f * (zm.x - zp.x));  // (Z[k] - conj(Z[N-k])) / (2j)
      sh.y[t] = cmul(__ldg(a.pl_resp + t), s);
    }
    __syncthreads();
    // ---- 64-point c2r (unnormalised, imaginary parts of DC and Nyquist ignored), last pl_l outputs kept (filter.c:140)
    if (t < pl_l) {
      const int n = 64 - pl_l + t;
      float acc = sh.y[0].x + ((n & 1) ? -sh.y[32].x : sh.y[32].x);
      for (int k = 1; k < 32; k++) {
        const float2 w = sh.cs[(k * n) & 63];
        acc += 2.0f * (sh.y[k].x * w.x - sh.y[k].y * w.y);
      }
      ring[(S.fft_ptr + t) & (PL_FFT - 1)] = acc;  // fm.c:239-249
    }
    S.fft_ptr = (S.fft_ptr + pl_l) & (PL_FFT - 1);
    S.last_fft += pl_l;
    // ---- every 512 samples: strongest bin of the 16384-point transform of the ring (fm.c:251-277)
    if (S.last_fft >= 512) {  // CTA-uniform
      S.last_fft = 0;
      __threadfence_block();
      __syncthreads();
#pragma unroll 1
      for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
          for (int q = 0; q < 8; q++) v[8 * e + q] = make_float2(__ldcg(ring + 8 * (t + 128 * e + 256 * q) + r), 0.f);
        fft2048<-1>(v, sh.buf, a.tw2048, sh.tw2);
        // v[j] = F_r[t + 128 j]; X[k' + 2048 q] += W_16384^(r k') W_8^(r q) F_r[k']
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int kp = t + 128 * j;
          float sn, cs;
          sincospif(-(float)(r * kp) * (2.0f / PL_FFT), &sn, &cs);
          const float2 f = cmul(v[j], make_float2(cs, sn));
#pragma unroll
          for (int q = 0; q < 4; q++) {
            // W_8^(r q) = exp(-j 2 pi r q / 8)
            float s8, c8;
            sincospif(-0.25f * (float)((r * q) & 7), &s8, &c8);
            const float2 g = cmul(f, make_float2(c8, s8));
            float2* dst = sh.acc + kp + 2048 * q;
            if (r == 0)
              *dst = g;
            else {
              const float2 o = *dst;
              *dst = make_float2(o.x + g.x, o.y + g.y);
            }
          }
        }
      }
      __syncthreads();
      float tot = 0.f, best = 0.f;
      int bestk = -1;
      for (int k = 1 + t; k < PL_FFT / 2; k += FFT2048_THREADS) {  // skip DC (fm.c:260)
        const float2 x = sh.acc[k];
        const float e = x.x * x.x + x.y * x.y;
        tot += e;
        if (e > best) {  // ascending k within the thread: the first maximum wins, as in the reference's scan
          best = e;
          bestk = k;
        }
      }
      tot = warp_sum(tot);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
        if (ob > best || (ob == best && ok >= 0 && (bestk < 0 || ok < bestk))) {
          best = ob;
          bestk = ok;
        }
      }
      if ((t & 31) == 0) {
        sh.red[t >> 5] = tot;
        sh.red[4 + (t >> 5)] = best;
        sh.iscal[t >> 5] = bestk;
      }
      __syncthreads();
      if (t == 0) {
        float T = (sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]);
        for (int w = 1; w < 4; w++)
          if (sh.red[4 + w] > best || (sh.red[4 + w] == best && sh.iscal[w] >= 0 && (bestk < 0 || sh.iscal[w] < bestk))) {
            best = sh.red[4 + w];
            bestk = sh.iscal[w];
          }
        float pf = S.plfreq;
        if (bestk > 0 && best > 0.01f * T) {  // fm.c:271-276
          const float f = (float)bestk * (a.dsamprate / PL_DEC) / PL_FFT;
          if (f > 67 && f < 255) pf = f;
        } else {
          pf = NAN;
        }
        sh.scal[0] = pf;
      }
      __syncthreads();
      S.plfreq = sh.scal[0];
      __syncthreads();
    }
    if (t == 0) a.status[(long long)b * a.nchan_total + wk.chan].reserved[1] = S.plfreq;  // demod->sig.plfreq
    __syncthreads();
  }
  if (t == 0) a.pl_state[blockIdx.x] = S;
}

int launch_pl(const ChanLaunch& a, int nchan_pl, cudaStream_t st) {
  if (nchan_pl <= 0) return 0;
  static bool configured_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    cudaFuncSetAttribute(pl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PlShared));
    configured_dev[dev & 63] = true;
  }
  pl_kernel<<<nchan_pl, FFT2048_THREADS, sizeof(PlShared), st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace k9
