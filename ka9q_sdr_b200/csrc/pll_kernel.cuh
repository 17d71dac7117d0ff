// Coherent linear modes: carrier-tracking PLL, squaring loop, FFT acquisition (reference linear.c:129-246) in front of the
// ordinary linear demodulator (linear.c:247-311). Included by chan_kernels.cu (uses its helpers). Mode rows with `pll` /
// `square` in modes.txt: CAM, AME, DSB, CISB.
//
// One CTA (128 threads) per channel, all blocks of the launch in order (everything here is carried state). Per block:
//   1. window x response -> 2048-point inverse FFT -> kept samples y[n] (with the block's LO phase, Appendix C)
//   2. y (or y^2 in squaring mode) appended to the channel's 65536-sample carrier-search ring            (linear.c:131-153)
//   3. lock detector with hysteresis on the PREVIOUS block's loop SNR                                     (linear.c:157-170)
//   4. unlocked and more than half a ring of new samples: 65536-point forward FFT of the ring, strongest bin inside
//      +-300 Hz (x2 when squaring) -> coarse oscillator; integrator reset when the bin moved              (linear.c:173-201)
//      The transform is evaluated only where it is needed: X[k] = sum_{r<32} W_65536^(r k) F_r[k mod 2048], F_r the
//      2048-point transform of the stride-32 subsequence r — 32 calls of the CTA's own fft2048, the +-410 (820) wanted bins
//      accumulated in registers. No host round trip, no library FFT.
//   5. y[n] *= coarse[n] * fine[n] (two NCOs, complex double), accum += y[n] (y[n]^2) -> carrier phase     (linear.c:207-223)
//      The NCO recurrences (osc.c:39-51, renormalised every 16384 steps) are evaluated in closed form in double:
//      phasor_0 * exp(j 2 pi f n); both agree with the recurrence to ~1e-15, nine orders below the fp32 samples they multiply.
//   6. lag-lead loop filter once per block -> fine NCO frequency; smoothed frequency offset                (linear.c:226-245)
//   7. the linear demodulator: I / Q powers, hang AGC (serial, one lane, the reference's operation order), post-detection
//      shift, scaleclip; loop SNR = signal / noise - 1                                                    (linear.c:247-310)
#pragma once

namespace k9 {

constexpr int PLL_FFT = 1 << 16;  // linear.c:43
constexpr int PLL_MAXBINS = 7;    // wanted bins per thread and side: ceil(820 / 128)

struct PllShared {
  float2 buf[NDEC];               // FFT exchange buffer
  float4 tw2[FFT2048_TW2_FLOAT4];
  float2 y[OLEN_MAX];             // kept samples of the block (LO phase applied), then PLL-rotated
  float amp[OLEN_MAX];
  float qg[OLEN_MAX];             // headroom / amp, then gain[n]
  float red[16];
  float scal[8];
  int iscal[4];
};

__device__ __forceinline__ double2 cmul_d(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cexp2pi_d(double cycles) {
  double s, c;
  sincospi(2.0 * (cycles - floor(cycles)), &s, &c);
  return make_double2(c, s);
}
__device__ __forceinline__ double2 cnormalize_d(double2 a) {
  const double r = rsqrt(a.x * a.x + a.y * a.y);
  return make_double2(a.x * r, a.y * r);
}

// strongest bin of the 65536-point forward transform of the ring inside [lowlimit, highlimit] (linear.c:178-189):
// returns the bin (first one in ascending order among equals) and its energy through sh.iscal[0] / sh.scal[0]
__device__ __noinline__ void pll_acquire(PllShared& sh, const float2* ring, const float2* __restrict__ tw2048,
                                         int lowlimit, int highlimit) {
  const int t = threadIdx.x;
  float2 accp[PLL_MAXBINS], accn[PLL_MAXBINS];  // bins k = t + 128 j (k >= 0) and k = t + 128 (16 - PLL_MAXBINS + j) - 2048 (k < 0)
#pragma unroll
  for (int j = 0; j < PLL_MAXBINS; j++) accp[j] = accn[j] = make_float2(0.f, 0.f);
  float2 v[16];
#pragma unroll 1
  for (int r = 0; r < 32; r++) {
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int q = 0; q < 8; q++)  // (L2-coherent loads: part of the ring was written by this CTA a moment ago)
        v[8 * e + q] = __ldcg(ring + 32 * (t + 128 * e + 256 * q) + r);
    fft2048<-1>(v, sh.buf, tw2048, sh.tw2);
    // v[j] = F_r[t + 128 j]
#pragma unroll
    for (int j = 0; j < PLL_MAXBINS; j++) {
      {
        const int k = t + 128 * j;
        float sn, cs;
        sincospif(-(float)((r * k) & (PLL_FFT - 1)) * (2.0f / PLL_FFT), &sn, &cs);
        const float2 w = make_float2(cs, sn), f = v[j];
        accp[j].x += f.x * w.x - f.y * w.y;
        accp[j].y += f.x * w.y + f.y * w.x;
      }
      {
        const int jj = 16 - PLL_MAXBINS + j;
        const int k = t + 128 * jj - NDEC;  // negative
        float sn, cs;
        sincospif(-(float)((r * k) & (PLL_FFT - 1)) * (2.0f / PLL_FFT), &sn, &cs);
        const float2 w = make_float2(cs, sn), f = v[jj];
        accn[j].x += f.x * w.x - f.y * w.y;
        accn[j].y += f.x * w.y + f.y * w.x;
      }
    }
  }
  // arg max over [lowlimit, highlimit], ties to the lower bin (ascending scan with a strict compare, linear.c:183-189)
  float best = 0.f;
  int bestk = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < PLL_MAXBINS; j++) {
    const int kp = t + 128 * j, kn = t + 128 * (16 - PLL_MAXBINS + j) - NDEC;
    const float ep = accp[j].x * accp[j].x + accp[j].y * accp[j].y;  // cnrmf
    const float en = accn[j].x * accn[j].x + accn[j].y * accn[j].y;
    if (kn >= lowlimit && kn <= highlimit && (en > best || (en == best && en > 0.f && kn < bestk))) {
      best = en;
      bestk = kn;
    }
    if (kp >= lowlimit && kp <= highlimit && (ep > best || (ep == best && ep > 0.f && kp < bestk))) {
      best = ep;
      bestk = kp;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
    if (ob > best || (ob == best && ok < bestk)) {
      best = ob;
      bestk = ok;
    }
  }
  __syncthreads();
  if ((t & 31) == 0) {
    sh.red[t >> 5] = best;
    sh.red[4 + (t >> 5)] = __int_as_float(bestk);
  }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < 4; w++) {
      const float ob = sh.red[w];
      const int ok = __float_as_int(sh.red[4 + w]);
      if (ob > best || (ob == best && ok < bestk)) {
        best = ob;
        bestk = ok;
      }
    }
    sh.scal[0] = best;
    sh.iscal[0] = best > 0.f ? bestk : 0;
  }
  __syncthreads();
}

template <int OLEN_T>
__global__ void __launch_bounds__(FFT2048_THREADS, 2) pll_kernel(const ChanLaunch a) {
  extern __shared__ __align__(16) unsigned char pll_raw[];
  PllShared& sh = *reinterpret_cast<PllShared*>(pll_raw);
  const int t = threadIdx.x;
  const int c = a.work[blockIdx.x].x, slot = a.work[blockIdx.x].y;
  const int olen = OLEN_T ? OLEN_T : a.olen;
  const int first = NDEC - olen;
  const int jb = first >> 7, rem = first & 127;
  const ChanParams P = a.params[c];
  const PllParams Q = a.pll_params[slot];
  PllState S = a.pll_state[slot];
  ChanState CS = a.state[c];
  float2* ring = a.pll_ring + (long long)slot * PLL_FFT;
  const bool isb = P.flags & CH_ISB, square = P.flags & CH_SQUARE;
  const float2* H = a.resp + (long long)P.resp_slot * NDEC;
  int eph = phase_index0(P.bin, a.start0, a.N);
  fft2048_stage_tw2(sh.tw2, a.tw2048);
  __syncthreads();
  float2 v[16];
#pragma unroll 1
  for (int b = 0; b < a.nblocks; b++) {
    const float2* X = a.spec + (long long)b * a.spec_stride;
    // ---- 1. predetection filter
    if (isb) {
      stage_filtered_isb(X, a.N, (int)P.bin, H, sh.buf);
      __syncthreads();
      load16(v, sh.buf);
    } else {
      load_filtered16(v, X, a.N, (int)P.bin, H);
    }
    fft2048<+1>(v, sh.buf, a.tw2048, sh.tw2);
    const float2 ph = phase_from_index(a, eph);
    const int ko = t - first;
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (j >= jb && ((j > jb) || (t >= rem))) sh.y[ko + 128 * j] = cmul(v[j], ph);
    __syncthreads();
    if (a.filt_dbg) dump_filter_output(a.filt_dbg + ((long long)b * a.nchan_total + c) * olen, sh.y, olen, make_float2(1.f, 0.f));
    // ---- 2. carrier-search ring
    for (int n = t; n < olen; n += FFT2048_THREADS) {
      const float2 y = sh.y[n];
      ring[(S.fft_ptr + n) & (PLL_FFT - 1)] = square ? cmul(y, y) : y;
    }
    S.fft_ptr = (S.fft_ptr + olen) & (PLL_FFT - 1);
    S.fft_samples = min(S.fft_samples + olen, PLL_FFT);
    // ---- 3. lock detector on the previous block's SNR (CTA-uniform: every thread carries the state)
    if (S.snr < Q.snrthresh)
      S.lock_count -= olen;
    else
      S.lock_count += olen;
    if (S.lock_count >= Q.lock_limit) {
      S.lock_count = Q.lock_limit;
      S.pll_lock = 1;
    }
    if (S.lock_count <= -Q.lock_limit) {
      S.lock_count = -Q.lock_limit;
      S.pll_lock = 0;
    }
    // ---- 4. acquisition
    if (!S.pll_lock && S.fft_samples > PLL_FFT / 2) {
      S.fft_samples = 0;
      __threadfence_block();
      __syncthreads();  // the ring rows just written are visible to the whole CTA
      pll_acquire(sh, ring, a.tw2048, Q.lowlimit, Q.highlimit);
      const float maxenergy = sh.scal[0];
      const int maxbin = sh.iscal[0];
      if (maxenergy > 0) {
        double new_delta_f = Q.binsize * maxbin;  // float product, as the reference's `binsize * maxbin`
        if (square) new_delta_f /= 2;
        if (new_delta_f != S.delta_f) {
          S.delta_f = new_delta_f;
          S.integrator = 0;
          S.coarse_freq = -Q.samptime * S.delta_f;  // set_osc(&coarse, -samptime * delta_f, 0.0): a float product
        }
      }
      __syncthreads();
    }
    // ---- 5. both NCOs applied, carrier phase
    {
      const double f = S.coarse_freq + S.fine_freq;
      const double2 psi0 = cmul_d(S.coarse_ph, S.fine_ph);
      double2 w = cmul_d(psi0, cexp2pi_d(f * (double)t));
      const double2 step = cexp2pi_d(f * 128.0);
      float2 acc = make_float2(0.f, 0.f);
      for (int n = t; n < olen; n += FFT2048_THREADS) {
        const float2 y = sh.y[n];
        const float2 z = make_float2((float)((double)y.x * w.x - (double)y.y * w.y), (float)((double)y.x * w.y + (double)y.y * w.x));
        sh.y[n] = z;
        const float2 ss = square ? cmul(z, z) : z;
        acc.x += ss.x;
        acc.y += ss.y;
        w = cmul_d(w, step);
      }
      float dummy = 0.f;
      block_reduce3<0>(acc.x, acc.y, dummy, sh.red);
      float cphase = atan2f(acc.y, acc.x);  // cargf(accum)
      if (isnan(cphase)) cphase = 0;
      if (square) cphase /= 2;
      S.cphase = cphase;
      // the oscillators after olen steps (phasor kept on the unit circle: osc.c:53-59 does that every 16384 steps)
      S.coarse_ph = cnormalize_d(cmul_d(S.coarse_ph, cexp2pi_d(S.coarse_freq * (double)olen)));
      S.fine_ph = cnormalize_d(cmul_d(S.fine_ph, cexp2pi_d(S.fine_freq * (double)olen)));
      // ---- 6. loop filter (linear.c:228-245)
      S.integrator += cphase * Q.blocktime;  // + ramp, which the reference has switched off (linear.c:67)
      const float feedback = Q.integrator_gain * S.integrator + Q.prop_gain * cphase;
      S.fine_freq = -feedback * Q.samptime;  // set_osc(&fine, -feedback * samptime, 0.0): a float product
      S.foffset = isnan(S.foffset) ? feedback + S.delta_f : (float)((double)S.foffset + 0.001 * (double)(feedback + S.delta_f - S.foffset));
    }
    // ---- 7. linear demodulator on the rotated samples
    float sig = 0.f, noi = 0.f, dummy = 0.f;
    for (int n = t; n < olen; n += FFT2048_THREADS) {
      const float2 s = sh.y[n];
      const float rp = s.x * s.x, ip = s.y * s.y;
      sig += rp;
      noi += ip;
      const float amp = sqrtf(rp + ip);
      sh.amp[n] = amp;
      sh.qg[n] = P.headroom / amp;
    }
    block_reduce3<0>(sig, noi, dummy, sh.red);  // (its barriers publish amp / qg)
    if (t == 0) {
      float gain = CS.agc_gain;
      int hang = CS.hang;
      const float headroom = P.headroom, rf = P.recovery_factor;
      const int hangmax = P.hangmax;
      for (int n = 0; n < olen; n++) {  // linear.c:269-279
        const float x = sh.amp[n], q = sh.qg[n];
        if (isnan(gain)) {
          gain = q;
        } else if (x * gain > headroom) {
          gain = q;
          hang = hangmax;
        } else if (hang != 0) {
          hang--;
        } else {
          gain *= rf;
        }
        sh.qg[n] = gain;
      }
      sh.scal[1] = gain;
      sh.iscal[1] = hang;
    }
    __syncthreads();
    CS.agc_gain = sh.scal[1];
    CS.hang = sh.iscal[1];
    {
      int16_t* pcm_row = a.pcm + (long long)b * a.pcm_stride + P.pcm_off;
      const bool shifted = P.shift_cycles != 0.0;
      const double phase0 = shifted ? CS.shift_phase + P.shift_cycles * (double)olen * b : 0.0;
      for (int n = t; n < olen; n += FFT2048_THREADS) {
        const float gn = sh.qg[n];
        float2 z = make_float2(sh.y[n].x * gn, sh.y[n].y * gn);  // linear.c:280
        if (shifted) {
          const double2 sp = cexp2pi_d(phase0 + P.shift_cycles * (double)n);
          z = cmul(z, make_float2((float)sp.x, (float)sp.y));
        }
        if (P.channels == 1) {
          pcm_row[n] = scaleclip(z.x);
        } else {
          pcm_row[2 * n] = scaleclip(z.x);
          pcm_row[2 * n + 1] = scaleclip(z.y);
        }
      }
    }
    // loop SNR (linear.c:304-309)
    if (noi != 0) {
      S.snr = sig / noi - 1;
      if (S.snr < 0) S.snr = 0;
    } else {
      S.snr = NAN;
    }
    if (t == 0) {
      ChanStatus st;
      st.bb_power = (sig + noi) / (2 * olen);
      st.snr = S.snr;
      st.foffset = S.foffset;
      st.pdeviation = 0.f;
      st.agc_gain = CS.agc_gain;
      st.squelch_open = S.pll_lock;   // PLL channels: the lock flag (demod->sig.pll_lock)
      st.reserved[0] = S.cphase;      // demod->sig.cphase
      st.reserved[1] = (float)S.lock_count;  // demod->sig.lock_timer
      a.status[(long long)b * a.nchan_total + c] = st;
    }
    eph = phase_advance(eph, P.phase_step, a.N);
    __syncthreads();
  }
  if (t == 0) {
    if (P.shift_cycles != 0.0) {
      const double p = CS.shift_phase + P.shift_cycles * (double)olen * a.nblocks;
      CS.shift_phase = p - floor(p);
    }
    a.state[c] = CS;
    a.pll_state[slot] = S;
  }
}

int launch_pll(const ChanLaunch& a, cudaStream_t st) {
  if (a.nwork <= 0) return 0;
  static bool configured_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    cudaFuncSetAttribute(pll_kernel<960>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PllShared));
    cudaFuncSetAttribute(pll_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PllShared));
    configured_dev[dev & 63] = true;
  }
  if (a.olen == 960)
    pll_kernel<960><<<a.nwork, FFT2048_THREADS, sizeof(PllShared), st>>>(a);
  else
    pll_kernel<0><<<a.nwork, FFT2048_THREADS, sizeof(PllShared), st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace k9
