// K3: fused per-channel kernels (sm_100a, fp32, no tensor cores — this is streaming FFT/demod work, HBM/issue bound).
//
// Each CTA (128 threads) owns one work item (an FM channel pair, or one AM / linear channel) for all blocks of the
// launch, so the carried per-channel state (discriminator state, AGC gain/hang, DC estimate, oscillator phase)
// lives in registers between consecutive 20 ms blocks and touches HBM once per launch.
//
// Per block and channel (reference file:line in brackets):
//   1. bin rotation: read the 2048-bin window of the shared N-point spectrum centred on the channel's carrier bin
//      (replaces the per-sample second-LO multiply, radio.c:132, by the identity of SURVEY Appendix C),
//      multiply by the channel response H                                   [filter.c:206-227, CROSS_CONJ :239-249]
//   2. 2048-point inverse FFT, keep the last olen samples (overlap discard)  [filter.c:250, :131]
//   3. per-block LO phase factor exp(j*2*pi*((-k*(m*L-(M-1))) mod N)/N)       [Appendix C]
//   4. demodulate + quantise                                               [fm.c / am.c / linear.c, audio.c:22-28]
#include <math.h>
#include "chan.cuh"
#include "fft2048.cuh"
#include "util.cuh"

namespace k9 {

// ---------------------------------------------------------------- common device helpers

struct CtaShared {
  float re[FFT2048_PLANE];
  float im[FFT2048_PLANE];
  float aux0[1024];   // FM: audio of channel A / AM,linear: amplitude
  float aux1[1024];   // FM: audio of channel B / AM: output / linear: per-sample gain
  float red[8];
  unsigned good[32];
  float scal[4];
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();  // protect red[] from the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}
__device__ __forceinline__ float block_min(float v, float* red) {
  v = warp_min(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fminf(fminf(red[0], red[1]), fminf(red[2], red[3]));
}

// float -> int16 as audio.c:22-28: clip, then C truncation toward zero of 32767*x
__device__ __forceinline__ int16_t scaleclip(float x) {
  if (x >= 1.0f) return 32767;
  if (x <= -1.0f) return -32768;
  return (int16_t)(int)(32767.0f * x);
}

// Step 1: v[8e+r] = Y[p], p = t + 128e + 256r, Y = H .* (rotated window of X)   [filter.c:206-227]
// ISB (CROSS_CONJ) folds the mirror bin in as filter.c:239-249 does.
__device__ __forceinline__ float2 load_bin(const float2* __restrict__ X, int N, long long bin, int p) {
  const int s = (p <= NDEC / 2) ? p : p - NDEC;  // signed bin offset: DC..+Nyquist, then negative frequencies
  long long idx = bin + s;
  if (idx < 0) idx += N;
  if (idx >= N) idx -= N;
  return __ldg(X + idx);
}

__device__ __forceinline__ void load_filtered(const float2* __restrict__ X, int N, long long bin,
                                              const float2* __restrict__ H, bool isb, float2 (&v)[16]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int e = 0; e < 2; e++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int p = t + 128 * e + 256 * r;
      float2 y = cmul(__ldg(H + p), load_bin(X, N, bin, p));
      if (isb && p != 0 && p != NDEC / 2) {
        const int pm = NDEC - p;
        float2 ym = cmul(__ldg(H + pm), load_bin(X, N, bin, pm));
        if (p < NDEC / 2)
          y = make_float2(y.x + ym.x, y.y - ym.y);  // pos + conj(neg)
        else
          y = make_float2(y.x - ym.x, y.y + ym.y);  // neg - conj(pos)
      }
      v[8 * e + r] = y;
    }
  }
}

// exp(j*theta_m) for block m of a channel at bin k (Appendix C): theta = 2*pi*((-k*(m*L-(M-1))) mod N)/N
__device__ __forceinline__ float2 block_phase(long long bin, long long m, int L, int M, int N) {
  long long start = (m * (long long)L - (long long)(M - 1)) % N;  // may be negative
  if (start < 0) start += N;
  long long e = (bin % N) * start % N;  // bin in [0,N), start in [0,N): product < 2^62
  e = (N - e) % N;                      // -k*start mod N
  double s, c;
  sincospi(2.0 * (double)e / (double)N, &s, &c);
  return make_float2((float)c, (float)s);
}

// Steps 1-3 for one channel-block. On return the olen kept samples are in sh.re/sh.im[0..olen) (synchronised),
// and *sumsq / *sumamp hold this thread's partial sums of |y|^2 and |y|.
__device__ __forceinline__ void channel_filter(const ChanLaunch& a, CtaShared& sh, int chan, const ChanParams& P, int b,
                                               float* sumsq, float* sumamp) {
  const int t = threadIdx.x;
  const float2* X = a.spec + (long long)b * a.spec_stride;
  float2 v[16];
  load_filtered(X, a.N, P.bin, a.resp + (long long)chan * NDEC, (P.flags & CH_ISB) != 0, v);
  fft2048<+1>(v, sh.re, sh.im, a.tw2048);
  const float2 ph = block_phase(P.bin, a.block0 + b, a.L, a.M, a.N);
  const int first = NDEC - a.olen;
  __syncthreads();  // everyone has read its stage-3 inputs; planes can be reused for y
  float ssq = 0.f, samp = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int n = t + 128 * j;
    if (n >= first) {
      const int o = n - first;
      const float2 y = cmul(v[j], ph);
      sh.re[o] = y.x;
      sh.im[o] = y.y;
      const float q = y.x * y.x + y.y * y.y;
      ssq += q;
      samp += sqrtf(q);
      if (a.filt_dbg) a.filt_dbg[((long long)b * a.nchan_total + chan) * a.olen + o] = y;
    }
  }
  *sumsq = ssq;
  *sumamp = samp;
  __syncthreads();
}

// ---------------------------------------------------------------- FM (pairs)

// index of the last good sample strictly below o, or -1
__device__ __forceinline__ int prev_good(const unsigned* good, int o) {
  int w = o >> 5;
  unsigned bits = (o & 31) ? (good[w] & ((1u << (o & 31)) - 1u)) : 0u;
  while (true) {
    if (bits) return (w << 5) + 31 - __clz(bits);
    if (--w < 0) return -1;
    bits = good[w];
  }
}

__device__ __forceinline__ float fm_arg(float2 y, float2 st) {
  // cargf(samp * state) (fm.c:131)
  const float re = y.x * st.x - y.y * st.y;
  const float im = y.x * st.y + y.y * st.x;
  return atan2f(im, re);
}

__global__ void __launch_bounds__(FFT2048_THREADS) fm_kernel(const ChanLaunch a) {
  __shared__ CtaShared sh;
  const int t = threadIdx.x;
  const int2 wk = a.work[blockIdx.x];
  const int chans[2] = {wk.x, wk.y};
  ChanParams P[2];
  ChanState S[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (chans[h] >= 0) {
      P[h] = a.params[chans[h]];
      S[h] = a.state[chans[h]];
    }
  }
  const int olen = a.olen;
  const int first = NDEC - olen;
  const bool filtered = P[0].audio_slot >= 0;

  for (int b = 0; b < a.nblocks; b++) {
    const long long m = a.block0 + b;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      float* aud = h ? sh.aux1 : sh.aux0;
      const int c = chans[h];
      if (c < 0) {
        for (int o = t; o < olen; o += FFT2048_THREADS) aud[o] = 0.f;
        continue;
      }
      float ssq, samp;
      channel_filter(a, sh, c, P[h], b, &ssq, &samp);
      // squelch statistics (fm.c:91-103)
      const float tot_sq = block_sum(ssq, sh.red);
      const float tot_amp = block_sum(samp, sh.red);
      const float bb_power = tot_sq / (2 * olen);
      const float avg_amp = tot_amp / ((float)M_SQRT2 * olen);
      const float fm_variance = bb_power - avg_amp * avg_amp;
      float snr = avg_amp * avg_amp / (2 * fm_variance) - 1;
      snr = fmaxf(0.0f, snr);
      if (snr > 2) {
        S[h].fm_below = 0;
      } else {
        if (++S[h].fm_below > 1000) S[h].fm_below = 1000;
      }
      const bool open = S[h].fm_below < 2;
      if (open) {
        const float min_ampl = 0.55f * 0.55f * avg_amp * avg_amp;  // fm.c:121
        // good-sample bitmap, one ballot per 32 samples
        for (int i = 0; i < 8; i++) {
          const int o = t + 128 * i;
          bool g = false;
          if (o < olen) {
            const float q = sh.re[o] * sh.re[o] + sh.im[o] * sh.im[o];
            g = q > min_ampl;
          }
          const unsigned mask = __ballot_sync(0xffffffffu, g);
          if ((t & 31) == 0) sh.good[(t >> 5) + 4 * i] = mask;
        }
        __syncthreads();
        float fsum = 0.f, pos = -INFINITY, neg = INFINITY;
        for (int i = 0; i < 8; i++) {
          const int o = t + 128 * i;
          if (o < olen) {
            const bool g = (sh.good[o >> 5] >> (o & 31)) & 1u;
            const int src = g ? o : prev_good(sh.good, o);  // sample whose discriminator value is emitted
            float audio;
            if (src < 0) {
              audio = S[h].fm_lastaudio;  // no good sample yet in this block: repeat the carried one (fm.c:141)
            } else {
              const int pg = prev_good(sh.good, src);
              const float2 st = (pg >= 0) ? make_float2(sh.re[pg], -sh.im[pg]) : S[h].fm_state;
              audio = fm_arg(make_float2(sh.re[src], sh.im[src]), st);
            }
            aud[o] = audio;
            fsum += audio;
            if (g && o > 0) {
              pos = fmaxf(pos, audio);
              neg = fminf(neg, audio);
            }
            if (o == 0) sh.scal[0] = g ? audio : 0.f;
            if (o == olen - 1) sh.scal[1] = audio;
          }
        }
        const float tot_f = block_sum(fsum, sh.red);
        const float init = sh.scal[0];
        const float pdev_pos = fmaxf(block_max(pos, sh.red), init);
        const float pdev_neg = fminf(block_min(neg, sh.red), init);
        const float avg_f = tot_f / olen;
        // carried discriminator state
        const int lg = prev_good(sh.good, olen);
        if (lg >= 0) S[h].fm_state = make_float2(sh.re[lg], -sh.im[lg]);
        S[h].fm_lastaudio = sh.scal[1];
        // frequency offset and peak deviation, only while the squelch is fully open (fm.c:145-154)
        if (S[h].fm_below < 1) {
          const float dsr = a.dsamprate;
          S[h].fm_foffset = (float)(dsr * avg_f * (0.5 * M_1_PI));
          S[h].fm_pdeviation = (float)(dsr * fmaxf(pdev_pos - avg_f, -(pdev_neg - avg_f)) * (0.5 * M_1_PI));
        }
      } else {
        // squelch closed (fm.c:155-160)
        S[h].fm_state = make_float2(0.f, 0.f);
        S[h].fm_lastaudio = 0.f;
        for (int o = t; o < olen; o += FFT2048_THREADS) aud[o] = 0.f;
      }
      if (t == 0) {
        ChanStatus st;
        st.bb_power = bb_power;
        st.snr = snr;
        st.foffset = S[h].fm_foffset;
        st.pdeviation = S[h].fm_pdeviation;
        st.agc_gain = P[h].fm_gain;
        st.squelch_open = open ? 1 : 0;
        st.reserved[0] = st.reserved[1] = 0.f;
        a.status[(long long)b * a.nchan_total + c] = st;
      }
      __syncthreads();
    }
    // ---- post-detection audio filter for the pair: REAL overlap-save, L=olen, M=NDEC-olen+1 (fm.c:39-66,162-171).
    // Two real channels ride one complex transform: z = audA + j audB, filtered by the real impulse response.
    int16_t* pcm_row = a.pcm + (long long)b * a.pcm_stride;
    if (filtered) {
      float2 v[16];
      const int ringbase = (int)(((m + 1) * (long long)olen) & (NDEC - 1));
      float* hA = a.audio_hist + (long long)chans[0] * NDEC;
      float* hB = chans[1] >= 0 ? a.audio_hist + (long long)chans[1] * NDEC : nullptr;
#pragma unroll
      for (int e = 0; e < 2; e++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int p = t + 128 * e + 256 * r;
          float2 z;
          if (p < first) {
            const int ri = (ringbase + p) & (NDEC - 1);
            z.x = hA[ri];
            z.y = hB ? hB[ri] : 0.f;
          } else {
            z.x = sh.aux0[p - first];
            z.y = sh.aux1[p - first];
            // the new samples become history for the next blocks
            const int ri = (ringbase + p) & (NDEC - 1);
            hA[ri] = z.x;
            if (hB) hB[ri] = z.y;
          }
          v[8 * e + r] = z;
        }
      }
      fft2048<-1>(v, sh.re, sh.im, a.tw2048);
      const float2* R = a.audio_resp + (long long)P[0].audio_slot * NDEC;
#pragma unroll
      for (int j = 0; j < 16; j++) v[j] = cmul(v[j], __ldg(R + t + 128 * j));
      fft2048_out_to_in(v);
      fft2048<+1>(v, sh.re, sh.im, a.tw2048);
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int n = t + 128 * j;
        if (n >= first) {
          const int o = n - first;
          pcm_row[P[0].pcm_off + o] = scaleclip(v[j].x * P[0].fm_gain);  // fm.c:169-170
          if (chans[1] >= 0) pcm_row[P[1].pcm_off + o] = scaleclip(v[j].y * P[1].fm_gain);
        }
      }
    } else {
      // FLAT: raw discriminator output goes out unfiltered and unscaled (fm.c:55,164-172)
      for (int o = t; o < olen; o += FFT2048_THREADS) {
        pcm_row[P[0].pcm_off + o] = scaleclip(sh.aux0[o]);
        if (chans[1] >= 0) pcm_row[P[1].pcm_off + o] = scaleclip(sh.aux1[o]);
      }
    }
    __syncthreads();
  }
  if (t == 0) {
#pragma unroll
    for (int h = 0; h < 2; h++)
      if (chans[h] >= 0) a.state[chans[h]] = S[h];
  }
}

// ---------------------------------------------------------------- AM (envelope)

__global__ void __launch_bounds__(FFT2048_THREADS) am_kernel(const ChanLaunch a) {
  __shared__ CtaShared sh;
  const int t = threadIdx.x;
  const int c = a.work[blockIdx.x].x;
  const ChanParams P = a.params[c];
  ChanState S = a.state[c];
  const int olen = a.olen;
  for (int b = 0; b < a.nblocks; b++) {
    float ssq, samp;
    channel_filter(a, sh, c, P, b, &ssq, &samp);
    for (int o = t; o < olen; o += FFT2048_THREADS)
      sh.aux0[o] = sqrtf(sh.re[o] * sh.re[o] + sh.im[o] * sh.im[o]);  // am.c:56-58
    const float signal = block_sum(ssq, sh.red);                      // includes the barrier publishing aux0
    if (t == 0) {
      // strictly serial recurrences: carrier-DC tracker and hang AGC (am.c:60-74); one lane, original operation order
      float gain = S.agc_gain, dc = S.am_dc;
      int hang = S.hang;
      const float headroom = P.headroom, rf = P.recovery_factor;
      const int hangmax = P.hangmax;
      for (int n = 0; n < olen; n++) {
        const float s = sh.aux0[n];
        dc += 0.0001f * (s - dc);
        if (isnan(gain)) {
          gain = headroom / dc;
        } else if (gain * dc > headroom) {
          gain = headroom / dc;
          hang = hangmax;
        } else if (hang != 0) {
          hang--;
        } else {
          gain *= rf;
        }
        sh.aux1[n] = (s - dc) * gain;
      }
      sh.scal[0] = gain;
      sh.scal[1] = dc;
      sh.scal[2] = __int_as_float(hang);
    }
    __syncthreads();
    S.agc_gain = sh.scal[0];
    S.am_dc = sh.scal[1];
    S.hang = __float_as_int(sh.scal[2]);
    int16_t* pcm_row = a.pcm + (long long)b * a.pcm_stride + P.pcm_off;
    for (int o = t; o < olen; o += FFT2048_THREADS) pcm_row[o] = scaleclip(sh.aux1[o]);
    if (t == 0) {
      ChanStatus st;
      st.bb_power = signal / (2 * olen);  // am.c:78 (noise term is identically 0 there)
      st.snr = NAN;
      st.foffset = 0.f;
      st.pdeviation = 0.f;
      st.agc_gain = S.agc_gain;
      st.squelch_open = 1;
      st.reserved[0] = S.am_dc;
      st.reserved[1] = 0.f;
      a.status[(long long)b * a.nchan_total + c] = st;
    }
    __syncthreads();
  }
  if (t == 0) a.state[c] = S;
}

// ---------------------------------------------------------------- linear (SSB / CW / IQ / ISB), no PLL

__global__ void __launch_bounds__(FFT2048_THREADS) linear_kernel(const ChanLaunch a) {
  __shared__ CtaShared sh;
  const int t = threadIdx.x;
  const int c = a.work[blockIdx.x].x;
  const ChanParams P = a.params[c];
  ChanState S = a.state[c];
  const int olen = a.olen;
  for (int b = 0; b < a.nblocks; b++) {
    float ssq, samp;
    channel_filter(a, sh, c, P, b, &ssq, &samp);
    float sig = 0.f, noi = 0.f;
    for (int o = t; o < olen; o += FFT2048_THREADS) {
      const float rp = sh.re[o] * sh.re[o], ip = sh.im[o] * sh.im[o];  // linear.c:256-259
      sig += rp;
      noi += ip;
      sh.aux0[o] = sqrtf(rp + ip);
    }
    const float signal = block_sum(sig, sh.red);
    const float noise = block_sum(noi, sh.red);
    if (t == 0) {
      // hang AGC (linear.c:269-280): serial, one lane, original operation order
      float gain = S.agc_gain;
      int hang = S.hang;
      const float headroom = P.headroom, rf = P.recovery_factor;
      const int hangmax = P.hangmax;
      for (int n = 0; n < olen; n++) {
        const float amplitude = sh.aux0[n];
        if (isnan(gain)) {
          gain = headroom / amplitude;
        } else if (amplitude * gain > headroom) {
          gain = headroom / amplitude;
          hang = hangmax;
        } else if (hang != 0) {
          hang--;
        } else {
          gain *= rf;
        }
        sh.aux1[n] = gain;
      }
      sh.scal[0] = gain;
      sh.scal[2] = __int_as_float(hang);
    }
    __syncthreads();
    S.agc_gain = sh.scal[0];
    S.hang = __float_as_int(sh.scal[2]);
    int16_t* pcm_row = a.pcm + (long long)b * a.pcm_stride + P.pcm_off;
    const bool shifted = P.shift_cycles != 0.0;
    for (int o = t; o < olen; o += FFT2048_THREADS) {
      const float g = sh.aux1[o];
      float2 z = make_float2(sh.re[o] * g, sh.im[o] * g);  // linear.c:280
      if (shifted) {
        // post-detection shift oscillator (linear.c:283-289, osc.c:39-51): phasor(n) = exp(j*2*pi*f*n), n counted
        // from the first sample the oscillator was stepped on
        double ph = S.shift_phase + P.shift_cycles * (double)o;
        ph -= floor(ph);
        double sn, cs;
        sincospi(2.0 * ph, &sn, &cs);
        z = cmul(z, make_float2((float)cs, (float)sn));
      }
      if (P.channels == 1) {
        pcm_row[o] = scaleclip(z.x);  // linear.c:291-296
      } else {
        pcm_row[2 * o] = scaleclip(z.x);  // I left, Q right (linear.c:299)
        pcm_row[2 * o + 1] = scaleclip(z.y);
      }
    }
    if (shifted) {
      double ph = S.shift_phase + P.shift_cycles * (double)olen;
      S.shift_phase = ph - floor(ph);
    }
    if (t == 0) {
      ChanStatus st;
      st.bb_power = (signal + noise) / (2 * olen);  // linear.c:302
      st.snr = NAN;                                 // linear.c:309 (no PLL)
      st.foffset = 0.f;
      st.pdeviation = 0.f;
      st.agc_gain = S.agc_gain;
      st.squelch_open = 1;
      st.reserved[0] = signal;
      st.reserved[1] = noise;
      a.status[(long long)b * a.nchan_total + c] = st;
    }
    __syncthreads();
  }
  if (t == 0) a.state[c] = S;
}

// ---------------------------------------------------------------- launchers

int launch_fm(const ChanLaunch& a, cudaStream_t st) {
  if (a.nwork <= 0) return 0;
  fm_kernel<<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_am(const ChanLaunch& a, cudaStream_t st) {
  if (a.nwork <= 0) return 0;
  am_kernel<<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_linear(const ChanLaunch& a, cudaStream_t st) {
  if (a.nwork <= 0) return 0;
  linear_kernel<<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace k9
