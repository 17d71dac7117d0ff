// K3: fused per-channel kernels (sm_100a, fp32, no tensor cores — this is streaming FFT/demod work, issue/HBM bound).
//
// Each CTA (128 threads) owns one work item (an FM channel pair, or one AM / linear channel) for all blocks of the
// launch, so the carried per-channel state (discriminator state, AGC gain/hang, DC estimate, oscillator phase)
// stays on chip between consecutive 20 ms blocks and touches HBM once per launch.
//
// Per block and channel (reference file:line in brackets):
//   1. bin rotation: read the 2048-bin window of the shared N-point spectrum centred on the channel's carrier bin
//      (replaces the per-sample second-LO multiply, radio.c:132, by the identity of SURVEY Appendix C),
//      multiply by the channel response H                                   [filter.c:206-227, CROSS_CONJ :239-249]
//   2. 2048-point inverse FFT, keep the last olen samples (overlap discard)  [filter.c:250, :131]
//   3. per-block LO phase factor exp(j*2*pi*((-k*(m*L-(M-1))) mod N)/N)       [Appendix C]
//   4. demodulate + quantise                                               [fm.c / am.c / linear.c, audio.c:22-28]
#include <math.h>
#include <stdlib.h>
#include "chan.cuh"
#include "fft2048.cuh"
#include "util.cuh"

namespace k9 {

// ---------------------------------------------------------------- common device helpers

#ifndef FM_FEWER_BARRIERS
#define FM_FEWER_BARRIERS 1
#endif
#ifndef FM_CARVEOUT_PCT
#define FM_CARVEOUT_PCT 70
#endif

// FM pairs keep no audio on chip: the discriminator appends its output straight to the channel's audio-history ring in
// global memory (it has to land there anyway, fm.c:162 / filter.c:164), and the audio transform reads the whole 2048
// sample window back from the ring (L1/L2 hits). 16.8 KB per CTA: 8 CTAs fit the 164 KB carve-out, which leaves
// ~90 KB of L1 for the spectrum windows that neighbouring channels share and the twiddle / de-emphasis tables.
// FM_TMA: spectrum windows staged by the TMA unit (1-D bulk copy, cp.async.bulk + mbarrier: UBLKCP in SASS) instead of
// sixteen 64-bit loads per thread. 0 = register-direct loads (LDG); 1 = bulk copy into the FFT exchange buffer itself
// (no extra shared memory, the copy's latency is exposed to the CTA and covered by the other CTAs of the SM);
// 2 = bulk copy into a dedicated landing buffer, issued one channel-job ahead so it lands under the previous job's
// transform (33 KB per CTA: 6 CTAs per SM). Measured A/B: DESIGN.md section 3.
#ifndef FM_TMA
#define FM_TMA 0
#endif
constexpr int TMA_WIN = NDEC + 2;  // the copy starts on an even bin (16-byte aligned source): 2050 bins cover any window

struct FmShared {
  alignas(16) float2 buf[FM_TMA == 1 ? TMA_WIN : NDEC];  // FFT exchange buffer; afterwards the olen kept samples y[0..olen)
#if FM_TMA == 2
  alignas(16) float2 land[TMA_WIN];  // landing buffer of the prefetched spectrum window
#endif
#if FM_TMA
  alignas(8) unsigned long long tma_bar;  // mbarrier the bulk copy completes on
#endif
  float4 tw2[FFT2048_TW2_FLOAT4];  // stage-2 twiddles of the 2048-point transform (fft2048.cuh)
  float red[16], red2[16];  // one scratch row per reduction of a channel-block: no barrier needed to recycle them
  unsigned good[32];
  float scal[8];
  int ephase[2];  // (k * block_start) mod N per channel of the work item
  ChanParams P[2];
  ChanState S[2];
};

// three block-wide reductions in one round trip. MODE 0: sum,sum,sum  1: sum,max,min  2: sum,sum,min
// PROTECT=false: the caller guarantees a barrier between the previous readers of red[] and this call.
template <int MODE, bool PROTECT = true>
__device__ __forceinline__ void block_reduce3(float& a, float& b, float& c, float* red) {
  a = warp_sum(a);
  b = MODE == 1 ? warp_max(b) : warp_sum(b);
  c = MODE == 0 ? warp_sum(c) : warp_min(c);
  if (PROTECT) __syncthreads();  // protect red[] from the previous use
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[w] = a;
    red[4 + w] = b;
    red[8 + w] = c;
  }
  __syncthreads();
  a = (red[0] + red[1]) + (red[2] + red[3]);
  b = MODE == 1 ? fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])) : (red[4] + red[5]) + (red[6] + red[7]);
  c = MODE == 0 ? (red[8] + red[9]) + (red[10] + red[11]) : fminf(fminf(red[8], red[9]), fminf(red[10], red[11]));
}

// float -> int16 as audio.c:22-28: clip, then C truncation toward zero of 32767*x
__device__ __forceinline__ int16_t scaleclip(float x) {
  int v = __float2int_rz(32767.0f * x);  // (short)(SHRT_MAX * x): truncation toward zero
  v = (x <= -1.0f) ? -32768 : v;
  v = (x >= 1.0f) ? 32767 : v;
  return (int16_t)v;
}

// Step 1: v[8e+r] = Y[p], p = t + 128e + 256r, Y = H .* (rotated window of X)   [filter.c:206-227]
// ISB (CROSS_CONJ) folds the mirror bin in as filter.c:239-249 does. 32-bit index math (N < 2^31).
__device__ __forceinline__ float2 load_bin(const float2* __restrict__ X, int N, int bin, int p) {
  const int s = (p <= NDEC / 2) ? p : p - NDEC;  // signed bin offset: DC..+Nyquist, then negative frequencies
  int idx = bin + s;                             // bin in [0,N), |s| <= 1024 < N
  if (idx < 0) idx += N;
  if (idx >= N) idx -= N;
  return __ldg(X + idx);
}

// ISB (CROSS_CONJ, filter.c:239-249): Y[p] = pos + conj(neg), Y[N_dec - p] = neg - conj(pos), staged into the shared buffer
// in natural order, p = t + 128k (rolled: every element needs its mirror bin as well).
__device__ __forceinline__ void stage_filtered_isb(const float2* __restrict__ X, int N, int bin,
                                                   const float2* __restrict__ H, float2* __restrict__ buf) {
  const int t = threadIdx.x;
#pragma unroll 2
  for (int k = 0; k < 16; k++) {
    const int p = t + 128 * k;
    float2 y = cmul(__ldg(H + p), load_bin(X, N, bin, p));
    if (p != 0 && p != NDEC / 2) {
      const int pm = NDEC - p;
      const float2 ym = cmul(__ldg(H + pm), load_bin(X, N, bin, pm));
      if (p < NDEC / 2)
        y = make_float2(y.x + ym.x, y.y - ym.y);  // pos + conj(neg)
      else
        y = make_float2(y.x - ym.x, y.y + ym.y);  // neg - conj(pos)
    }
    buf[p] = y;
  }
}

// registers <- staged input: v[8e + r] = buf[t + 128e + 256r] (conflict-free, consecutive lanes)
__device__ __forceinline__ void load16(float2 (&v)[16], const float2* __restrict__ buf) {
  const float2* bp = buf + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int r = 0; r < 8; r++) v[8 * e + r] = bp[128 * e + 256 * r];
}
// Filtered spectrum window straight into the transform's input registers (no shared-memory staging): row j = e + 2r of
// thread t is bin p = t + 128j, v[8e + r] = H[p] * X[(bin + s(p)) mod N], s(p) = p for p <= N_dec/2, else p - N_dec
// (filter.c:206-227). Consecutive lanes read consecutive bins, so both streams are coalesced.
__device__ __forceinline__ void load_filtered16(float2 (&v)[16], const float2* __restrict__ X, int N, int bin,
                                                const float2* __restrict__ H) {
  const int t = threadIdx.x;
  if (bin >= 1024 && bin + 1024 < N) {
    // the whole window lies inside [0, N) (every channel but the few at the band edges; CTA-uniform): one base
    // pointer, constant offsets. Row 8 of thread 0 is the Nyquist bin, which belongs to the positive side
    // (filter.c:206: p <= N_dec/2).
    const float2* p0 = X + bin + t;
    const float2* pn = t ? p0 - 1024 : p0 + 1024;
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int j = e + 2 * r;
        v[8 * e + r] = __ldg(j < 8 ? p0 + 128 * j : (j == 8 ? pn : p0 + 128 * (j - 16)));
      }
  } else {
    int i0 = bin + t;  // rows 0..7: positive frequencies
    if (i0 < 0) i0 += N;
    if (i0 >= N) i0 -= N;
    int i1 = bin + t - 1024;  // rows 8..15: negative frequencies
    if (i1 < 0) i1 += N;
    if (i1 >= N) i1 -= N;
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int j = e + 2 * r;
        int i = (j < 8 ? i0 : i1) + 128 * (j & 7);
        if (i >= N) i -= N;
        if (j == 8 && t == 0) {  // the Nyquist bin
          i = bin + 1024;
          if (i >= N) i -= N;
        }
        v[8 * e + r] = __ldg(X + i);
      }
  }
  const float2* Hp = H + t;
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int r = 0; r < 8; r++) v[8 * e + r] = cmul(__ldg(Hp + 128 * (e + 2 * r)), v[8 * e + r]);
}
#if FM_TMA
// ---- TMA-staged window (cp.async.bulk: global -> shared, completion on an mbarrier) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// the whole 2048-bin window of a channel lies inside [0, N) together with the alignment pad (all but the band-edge channels)
__device__ __forceinline__ bool tma_window_inside(int bin, int N) { return bin >= 1024 && bin + 1026 <= N; }
// one elected thread: order the CTA's earlier generic-proxy accesses to dst before the async-proxy write, arm the
// barrier with the byte count and start the copy of bins [start, start + 2050), start = (bin - 1023) rounded down to even
__device__ __forceinline__ void tma_issue_window(float2* dst, const float2* __restrict__ X, int bin, unsigned long long* bar) {
  const int start = (bin - 1023) & ~1;
  const unsigned bytes = TMA_WIN * sizeof(float2);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(X + start), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// staged window -> transform input registers, times the response: row j = e + 2r of thread t is FFT index p = t + 128 j;
// p <= 1024 is bin + p (G[p + 1023 + off]), p > 1024 is bin + p - 2048 (G[p - 1025 + off]); off = 0 / 1 from the alignment
__device__ __forceinline__ void load_filtered16_staged(float2 (&v)[16], const float2* __restrict__ G, int bin,
                                                       const float2* __restrict__ H) {
  const int t = threadIdx.x;
  const int off = (bin - 1023) & 1;
  const float2* gp = G + t + 1023 + off;  // rows 0..7
  const float2* gn = G + t - 1025 + off;  // rows 9..15 (+128 j); row 8: Nyquist for t == 0
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int j = e + 2 * r;
      v[8 * e + r] = j < 8 ? gp[128 * j] : (j == 8 ? (t ? gn[1024] : gp[1024]) : gn[128 * j]);
    }
  const float2* Hp = H + t;
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int r = 0; r < 8; r++) v[8 * e + r] = cmul(__ldg(Hp + 128 * (e + 2 * r)), v[8 * e + r]);
}
#endif

// transform output -> buffer in natural order: buf[t + 128j] = v[j], only the rows that hold kept samples (j >= jb)
__device__ __forceinline__ void store16(const float2 (&v)[16], float2* __restrict__ buf, int jb) {
  float2* bp = buf + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 16; j++)
    if (j >= jb) bp[128 * j] = v[j];
}

// Per-block LO phase (Appendix C): exp(j*2*pi*((-k*start_m) mod N)/N) = W_N^(e), e = (k*start_m) mod N with the forward
// table W_N^a = exp(-2*pi*i*a/N) split as lo[a & 1023] * hi[a >> 10] (the forward FFT's own twiddle tables).
__device__ __forceinline__ float2 phase_from_index(const ChanLaunch& a, int e) {
  return cmul(__ldg(a.twN_lo + (e & 1023)), __ldg(a.twN_hi + (e >> 10)));
}
// e for the first block of the launch: one 64-bit modular product per channel per launch
__device__ __forceinline__ int phase_index0(long long bin, int start0, int N) { return (int)((bin * (long long)start0) % N); }
__device__ __forceinline__ int phase_advance(int e, int step, int N) {
  e += step;
  return e >= N ? e - N : e;
}

// After the inverse transform: store the kept rows (j >= jb; the row j == jb is partly history) for the discriminator and,
// in the same pass over the registers, this thread's partial sums of |y|^2 and |y| and its minimum |y|^2 over the kept
// samples n in [first, NDEC). The per-block LO phase (Appendix C) is NOT applied here: |y| does not depend on it; the
// consumers that do (discriminator state across blocks, linear output, the debug capture) apply it themselves, once per
// block or fused into a multiply they already do.
__device__ __forceinline__ void store16_stats(const float2 (&v)[16], float2* __restrict__ buf, int first, float* sumsq,
                                              float* sumamp, float* minsq) {
  const int t = threadIdx.x;
  const int jb = first >> 7, rem = first & 127;
  float2* bp = buf + t;
  float ssq = 0.f, samp = 0.f, mn = INFINITY;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    if (j >= jb) {  // warp-uniform
      bp[128 * j] = v[j];
      float q = v[j].x * v[j].x + v[j].y * v[j].y;
      const bool kept = (j > jb) || (t >= rem);
      ssq += kept ? q : 0.f;
      mn = fminf(mn, kept ? q : INFINITY);
      // |y| for the squelch statistics only (fm.c:95): MUFU.RSQ based, ~1 ulp; it only feeds threshold decisions
      q = fmaxf(q, 1e-37f);
      samp += kept ? q * rsqrtf(q) : 0.f;
    }
  }
  *sumsq = ssq;
  *sumamp = samp;
  *minsq = mn;
}

// optional raw filter-output capture for the parity tests (off the hot path)
__device__ __noinline__ void dump_filter_output(float2* dst, const float2* src, int olen, float2 ph) {
#pragma unroll 1
  for (int o = threadIdx.x; o < olen; o += FFT2048_THREADS) dst[o] = cmul(src[o], ph);
}

// ---------------------------------------------------------------- FM (pairs)

#ifndef FM_CTAS_PER_SM
#define FM_CTAS_PER_SM 8  // 64 registers/thread, 8 x 25 KB shared memory (6 CTAs x 80 registers measured the same)
#endif

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

// atan2 for the discriminator: octant reduction + degree-8 minimax polynomial in (min/max)^2, max abs error 9e-8 rad
// before the final pi/2, pi reflections (fitted against float64 atan over [0,1]; the discriminator output is scaled by
// ~5.6e3 LSB/rad, so this is < 1e-3 LSB). Zero / non-finite arguments take libm's atan2f so the special values
// (e.g. state == 0 right after the squelch opens, fm.c:156) follow IEEE exactly as the reference's cargf does.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  if (!(mx > 0.f && mx < 1e30f)) return atan2f(y, x);
  const float a = __fdividef(mn, mx);
  const float s = a * a;
  float p = 0.002456682501360774f;
  p = fmaf(p, s, -0.014401180669665337f);
  p = fmaf(p, s, 0.03978091850876808f);
  p = fmaf(p, s, -0.07234828919172287f);
  p = fmaf(p, s, 0.10498931258916855f);
  p = fmaf(p, s, -0.141612246632576f);
  p = fmaf(p, s, 0.19985906779766083f);
  p = fmaf(p, s, -0.33332598209381104f);
  p = fmaf(p, s, 0.9999998807907104f);
  float r = p * a;
  if (ay > ax) r = 1.57079637f - r;
  if (x < 0.f) r = 3.14159274f - r;
  return copysignf(r, y);
}

__device__ __forceinline__ float fm_arg(float2 y, float2 st) {
  // cargf(samp * state) (fm.c:131)
  const float re = y.x * st.x - y.y * st.y;
  const float im = y.x * st.y + y.y * st.x;
  return fast_atan2f(im, re);
}

// Cold path of the discriminator, kept out of line so it does not sit in the hot loop's instruction-cache footprint:
// some sample of the block is below the blanking threshold (fm.c:121,130,141). Builds the good-sample bitmap (one
// ballot per 32 samples), then audio[n] = arg(y[src] * conj(y[prev good before src])), src = last good sample <= n.
__device__ __noinline__ void fm_discriminate_blanked(const float2* __restrict__ ybuf, int olen, float min_ampl,
                                                     float2 old_state, float old_last, float* __restrict__ aud, int rb,
                                                     unsigned* good, float* scal, float* fsum_out, float* pos_out,
                                                     float* neg_out) {
  const int t = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    const int o = t + 128 * i;
    bool g = false;
    if (o < olen) {
      const float2 y = ybuf[o];
      g = (y.x * y.x + y.y * y.y) > min_ampl;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, g);
    if ((t & 31) == 0) good[(t >> 5) + 4 * i] = mask;
  }
  __syncthreads();
  float fsum = 0.f, pos = -INFINITY, neg = INFINITY;
  const int lane = t & 31;
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    const int o = t + 128 * i;
    if (o < olen) {
      const unsigned wbits = good[o >> 5];
      const bool g = (wbits >> lane) & 1u;
      const bool gp = lane ? ((wbits >> (lane - 1)) & 1u) : (o ? (good[(o >> 5) - 1] >> 31) : 1u);
      float2 ys = ybuf[o], st = old_state;  // old_state is already conj(previous block's last good sample)
      bool have = true, cj = false;
      if (g && gp) {
        if (o > 0) {
          st = ybuf[o - 1];
          cj = true;
        }
      } else {
        const int src = g ? o : prev_good(good, o);
        have = src >= 0;  // no good sample yet in this block: repeat the carried audio value
        if (have) {
          ys = ybuf[src];
          const int pg = prev_good(good, src);
          if (pg >= 0) {
            st = ybuf[pg];
            cj = true;
          }
        }
      }
      if (cj) st.y = -st.y;
      const float audio = have ? fm_arg(ys, st) : old_last;
      aud[(rb + o) & (NDEC - 1)] = audio;
      fsum += audio;
      if (g && o > 0) {
        pos = fmaxf(pos, audio);
        neg = fminf(neg, audio);
      }
      if (o == 0) scal[0] = g ? audio : 0.f;
      if (o == olen - 1) scal[1] = audio;
    }
  }
  *fsum_out = fsum;
  *pos_out = pos;
  *neg_out = neg;
}

// Off-grid carrier (ka9q_stream_set_fine_lo): rotate the kept samples by the fine part of the second LO,
// exp(j 2 pi cyc n'), n' = output samples since stream start. The within-block part is applied here, in place; the phase
// at the block's first sample is returned and joins the per-block LO phase. Cold, out of line.
__device__ __noinline__ float2 fm_fine_rotate(float2* ybuf, int olen, double cyc, long long m) {
  const float fc = (float)cyc;
#pragma unroll 1
  for (int o = threadIdx.x; o < olen; o += FFT2048_THREADS) {
    float sn, cs;
    sincospif(2.f * fc * (float)o, &sn, &cs);
    ybuf[o] = cmul(ybuf[o], make_float2(cs, sn));
  }
  double base = cyc * (double)(m * olen);
  base -= floor(base);
  double sn, cs;
  sincospi(2.0 * base, &sn, &cs);
  return make_float2((float)cs, (float)sn);
}

// Squelch + discriminator for one channel-block whose kept samples (without the block's LO phase ph) are in sh.buf
// (fm.c:86-160). Appends olen audio samples to the ring aud[] at rb, updates sh.S[h] and the status row. The discriminator only sees
// phase differences, so ph enters once: the carried state conj(last good sample) is kept in the true (rotated) domain
// and moved into / out of this block's unrotated domain with one complex multiply each way.
template <class SH>
__device__ __forceinline__ void fm_discriminate(const ChanLaunch& a, SH& sh, const int olen, int h, int c, int b,
                                                float ssq, float samp, float minsq, float2 ph,
                                                float* __restrict__ aud, int rb) {
  const int t = threadIdx.x;
  const float2* ybuf = sh.buf + (NDEC - olen);  // kept samples y[0..olen)
  // (its barrier also publishes the kept samples store16_stats just wrote; red[] was last read a barrier ago)
  block_reduce3<2, !FM_FEWER_BARRIERS>(ssq, samp, minsq, sh.red);
  if (sh.P[h].shift_cycles != 0.0) {  // FM: the field holds the fine LO of an off-grid carrier
    ph = cmul(ph, fm_fine_rotate(sh.buf + (NDEC - olen), olen, sh.P[h].shift_cycles, a.block0 + b));
    __syncthreads();
  }
  if (a.filt_dbg) dump_filter_output(a.filt_dbg + ((long long)b * a.nchan_total + c) * olen, ybuf, olen, ph);
  const float bb_power = ssq / (2 * olen);
  const float avg_amp = samp / ((float)M_SQRT2 * olen);
  const float fm_variance = bb_power - avg_amp * avg_amp;
  float snr = avg_amp * avg_amp / (2 * fm_variance) - 1;
  snr = fmaxf(0.0f, snr);
  int below = sh.S[h].fm_below;
  if (snr > 2) {  // fm.c:108-113
    below = 0;
  } else {
    if (++below > 1000) below = 1000;
  }
  const bool open = below < 2;
  float2 new_state = make_float2(0.f, 0.f);
  float dbg_allgood = -1.f;
  float new_last = 0.f, foffset = sh.S[h].fm_foffset, pdeviation = sh.S[h].fm_pdeviation;
  if (open) {
    const float min_ampl = 0.55f * 0.55f * avg_amp * avg_amp;  // fm.c:121
    // into this block's unrotated domain. A carried state of exactly (0,0) (the squelch was shut, fm.c:156) stays
    // (+0,+0): the rotation would hand atan2 other zero signs than the reference's cargf(samp * 0) sees, and the first
    // audio sample after the squelch re-opens is 0 or +-pi depending on exactly those signs.
    const float2 st0 = sh.S[h].fm_state;
    const float2 old_state = (st0.x == 0.f && st0.y == 0.f) ? make_float2(0.f, 0.f) : cmul(st0, ph);
    const float old_last = sh.S[h].fm_lastaudio;
    const bool all_good = minsq > min_ampl;  // every sample passes the blanking threshold (the usual case)
    dbg_allgood = all_good ? 1.f : 0.f;
    float fsum = 0.f, pos = -INFINITY, neg = INFINITY;
    if (all_good) {
      // Uniform fast path (a separate loop on purpose: merged with the general one the compiler predicates the
      // blanking logic and every sample pays for it). audio[n] = arg(y[n] * conj(y[n-1])) (fm.c:130-132).
      {
        const float a0 = fm_arg(ybuf[t], t ? make_float2(ybuf[t - 1].x, -ybuf[t - 1].y) : old_state);
        aud[(rb + t) & (NDEC - 1)] = a0;
        fsum = a0;
        if (t) {
          pos = a0;
          neg = a0;
        } else {
          sh.scal[0] = a0;
        }
      }
#pragma unroll 2  // (4 and 8 measured slower: code size)
      for (int o = t + 128; o < olen; o += FFT2048_THREADS) {
        const float2 yp = ybuf[o - 1];
        const float audio = fm_arg(ybuf[o], make_float2(yp.x, -yp.y));
        aud[(rb + o) & (NDEC - 1)] = audio;
        fsum += audio;
        pos = fmaxf(pos, audio);
        neg = fminf(neg, audio);
      }
      if (t == ((olen - 1) & 127)) sh.scal[1] = aud[(rb + olen - 1) & (NDEC - 1)];  // this thread wrote it
    } else {
      fm_discriminate_blanked(ybuf, olen, min_ampl, old_state, old_last, aud, rb, sh.good, sh.scal, &fsum, &pos, &neg);
    }
    block_reduce3<1, !FM_FEWER_BARRIERS>(fsum, pos, neg, FM_FEWER_BARRIERS ? sh.red2 : sh.red);
    const float init = sh.scal[0];
    const float pdev_pos = fmaxf(pos, init);
    const float pdev_neg = fminf(neg, init);
    const float avg_f = fsum / olen;
    const int lg = all_good ? olen - 1 : prev_good(sh.good, olen);
    new_state = sh.S[h].fm_state;
    if (lg >= 0) {
      const float2 yl = ybuf[lg];
      new_state = cmul(make_float2(yl.x, -yl.y), make_float2(ph.x, -ph.y));  // conj(y * ph): back to the true domain
    }
    new_last = sh.scal[1];
    if (below < 1) {  // frequency offset and peak deviation only while fully open (fm.c:145-154)
      const float dsr = a.dsamprate;
      foffset = (float)(dsr * avg_f * (0.5 * M_1_PI));
      pdeviation = (float)(dsr * fmaxf(pdev_pos - avg_f, -(pdev_neg - avg_f)) * (0.5 * M_1_PI));
    }
  } else {
    // squelch closed (fm.c:155-160)
#pragma unroll 1
    for (int o = t; o < olen; o += FFT2048_THREADS) aud[(rb + o) & (NDEC - 1)] = 0.f;
  }
  __syncthreads();  // all reads of sh.S[h], sh.scal and y are done
  if (t == 0) {
    sh.S[h].fm_below = below;
    sh.S[h].fm_state = new_state;
    sh.S[h].fm_lastaudio = new_last;
    sh.S[h].fm_foffset = foffset;
    sh.S[h].fm_pdeviation = pdeviation;
    ChanStatus st;
    st.bb_power = bb_power;
    st.snr = snr;
    st.foffset = foffset;
    st.pdeviation = pdeviation;
    st.agc_gain = sh.P[h].fm_gain;
    st.squelch_open = open ? 1 : 0;
    st.reserved[0] = dbg_allgood;  // 1: no sample was blanked in this block, 0: some were, -1: squelch shut
    st.reserved[1] = 0.f;
    a.status[(long long)b * a.nchan_total + c] = st;
  }
}

// FLAT mode: the raw discriminator output goes out unfiltered and unscaled (fm.c:55,164-172). Cold, out of line.
__device__ __noinline__ void fm_flat_output(const float* audA, const float* audB, int rb, int16_t* pa, int16_t* pb, int olen) {
  __syncthreads();
#pragma unroll 1
  for (int o = threadIdx.x; o < olen; o += FFT2048_THREADS) {
    const int ri = (rb + o) & (NDEC - 1);
    pa[o] = scaleclip(audA[ri]);
    if (pb) pb[o] = scaleclip(audB[ri]);
  }
}

// ---- block-split form (few pairs per GPU) ----
// A cluster of `fm_split` CTAs shares one pair: CTA e takes blocks b = e (mod split). The transforms of different blocks
// are independent; what chains the blocks of a channel is small: the discriminator / squelch state (ChanState) and the
// audio ring. So block b's CTA runs its first predetection transform at once, then waits until block b-1's CTA has (1)
// finished both discriminators and (2) read its own 2048-sample audio window out of the ring and used it (block b's
// audio overwrites the oldest part of that window). State travels through a.state[], the order through one sequence number
// per pair (absolute block count, release / acquire at GPU scope). The cluster guarantees that the CTAs are co-resident.
// want: blocks of this pair discriminated since stream start, as the block about to be discriminated needs it
__device__ __noinline__ void fm_split_wait(const long long* f, const ChanState* state, long long want, FmShared& sh, int2 wk) {
  const int t = threadIdx.x;
  if (t == 0) {
    const long long t0 = clock64();
    while (true) {
      long long v;
      asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= want) break;
      __nanosleep(40);
      if (clock64() - t0 > 8000000000LL) __trap();  // ~4 s: the partner CTA is gone; fail loudly instead of hanging
    }
  }
  __syncthreads();
  if (t < 2) {
    const int c = t ? wk.y : wk.x;
    if (c >= 0) {
      const int* src = reinterpret_cast<const int*>(state + c);
      int* dst = reinterpret_cast<int*>(&sh.S[t]);
#pragma unroll
      for (int i = 0; i < (int)(sizeof(ChanState) / sizeof(int)); i++) dst[i] = __ldcg(src + i);
    }
  }
  __syncthreads();
}
__device__ __noinline__ void fm_split_signal(long long* f, ChanState* state, long long done, FmShared& sh, int2 wk) {
  const int t = threadIdx.x;
  if (t < 2) {
    const int c = t ? wk.y : wk.x;
    if (c >= 0) state[c] = sh.S[t];
  }
  __syncthreads();  // every thread's ring loads and the two state rows are ordered before the release below
  if (t == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(f), "l"(done) : "memory");
  }
}

// OLEN_T: output samples per block known at compile time (960 = the reference geometry at every input rate, SURVEY
// Appendix B) so the kept-row tests fold away; 0 = take it from the launch arguments.
template <int OLEN_T, bool SPLIT>
__global__ void __launch_bounds__(FFT2048_THREADS, FM_CTAS_PER_SM) fm_kernel(const ChanLaunch a) {
  __shared__ FmShared sh;
  const int t = threadIdx.x;
  const int split = SPLIT ? a.fm_split : 1;          // CTAs per pair (= cluster size)
  const int pairi = SPLIT ? blockIdx.x / split : blockIdx.x;
  const int e0 = SPLIT ? blockIdx.x % split : 0;     // this CTA's first block
  const int2 wk = a.work[pairi];
  fft2048_stage_tw2(sh.tw2, a.tw2048);
  if (t < 2) {
    const int c = t ? wk.y : wk.x;
    if (c >= 0) {
      sh.P[t] = a.params[c];
      sh.S[t] = a.state[c];
      int ep = phase_index0(sh.P[t].bin, a.start0, a.N);
      for (int i = 0; i < e0; i++) ep = phase_advance(ep, sh.P[t].phase_step, a.N);
      sh.ephase[t] = ep;
    }
  }
#if FM_TMA
  if (t == 0) mbar_init(&sh.tma_bar, 1);
  unsigned tma_phase = 0;
#endif
  __syncthreads();
  const int olen = OLEN_T ? OLEN_T : a.olen;
  const int first = NDEC - olen;
  const int jb = first >> 7, rem = first & 127;
  const bool filtered = sh.P[0].audio_slot >= 0;
#if FM_TMA
  // CTA-uniform: every window of the work item lies inside [0, N) (band-edge channels take the LDG path, which wraps)
  const bool use_tma = tma_window_inside((int)sh.P[0].bin, a.N) && (wk.y < 0 || tma_window_inside((int)sh.P[1].bin, a.N));
#endif
#if FM_TMA == 2
  // prefetch pipeline: the window of the next channel-job is copied while the current one is transformed
  if (t == 0 && e0 < a.nblocks && use_tma)
    tma_issue_window(sh.land, a.spec + (long long)e0 * a.spec_stride, (int)sh.P[0].bin, &sh.tma_bar);
#endif
  // Audio-history rings of the pair (2048 floats each). An absent channel B reads the spare all-zero ring that follows
  // the last channel's; nothing is ever written to it.
  float* const histA = a.audio_hist + (long long)wk.x * NDEC;
  float* const histB = a.audio_hist + (long long)(wk.y >= 0 ? wk.y : a.nchan_total) * NDEC;
  float2 v[16];

  for (int b = e0; b < a.nblocks; b += split) {
    const long long m = a.block0 + b;
    const float2* X = a.spec + (long long)b * a.spec_stride;
    int16_t* pcm_row = a.pcm + (long long)b * a.pcm_stride;
    // REAL overlap-save input of the post-detection filter, L=olen, M=NDEC-olen+1 (fm.c:39-43): the 2048-sample window
    // of block m is ring[(ringbase + p) & 2047], p = 0..2047; its last olen entries are this block's new audio.
    const int ringbase = (int)(((m + 1) * (long long)olen) & (NDEC - 1));
    // Four transforms per pair-block share ONE copy of the FFT code: job 0/1 = predetection filter of channel A/B,
    // job 2 = forward transform of the audio pair (as conj(IFFT(conj z))), job 3 = its inverse.
#pragma unroll 1
    for (int job = 0; job < 4; job++) {
      const int h = job & 1;
      const int c = h ? wk.y : wk.x;
      if (job < 2) {
        if (c < 0) continue;
#if FM_TMA == 1
        if (use_tma) {
          __syncthreads();  // every earlier reader / writer of the exchange buffer is done
          if (t == 0) tma_issue_window(sh.buf, X, (int)sh.P[h].bin, &sh.tma_bar);
          mbar_wait(&sh.tma_bar, tma_phase);
          tma_phase ^= 1;
          load_filtered16_staged(v, sh.buf, (int)sh.P[h].bin, a.resp + (long long)sh.P[h].resp_slot * NDEC);
        } else
#elif FM_TMA == 2
        if (use_tma) {
          mbar_wait(&sh.tma_bar, tma_phase);
          tma_phase ^= 1;
          load_filtered16_staged(v, sh.land, (int)sh.P[h].bin, a.resp + (long long)sh.P[h].resp_slot * NDEC);
          __syncthreads();  // the landing buffer has been read by everybody: start the next window's copy
          if (t == 0) {
            if (h == 0 && wk.y >= 0)
              tma_issue_window(sh.land, X, (int)sh.P[1].bin, &sh.tma_bar);
            else if (b + split < a.nblocks)  // this CTA's next block
              tma_issue_window(sh.land, X + (long long)split * a.spec_stride, (int)sh.P[0].bin, &sh.tma_bar);
          }
        } else
#endif
          load_filtered16(v, X, a.N, (int)sh.P[h].bin, a.resp + (long long)sh.P[h].resp_slot * NDEC);
      } else if (job == 2) {
        if (!filtered) {
          if (SPLIT) fm_split_signal(a.fm_seq + pairi, a.state, m + 1, sh, wk);
          break;
        }
        // Two real channels ride one complex transform, z = audA + j audB (the filter's impulse response is real),
        // straight into the transform's input registers: row j = e + 2r of thread t is sample p = t + 128j. The
        // discriminators appended this block's samples to the rings (global memory, made visible to the CTA by the
        // barrier that ends fm_discriminate). conj on the way in: forward transform via the backward code.
        const int r0 = (ringbase + t) & (NDEC - 1);
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int ri = (r0 + 128 * (e + 2 * r)) & (NDEC - 1);
            // (block-split form: the older part of the window was written by another CTA, so no L1)
            v[8 * e + r] = SPLIT ? make_float2(__ldcg(histA + ri), -__ldcg(histB + ri)) : make_float2(histA[ri], -histB[ri]);
          }
      }
      // the buffer's previous readers are a barrier behind us except after job 2 (its last stage just read it)
      fft2048<+1>(v, sh.buf, a.tw2048, sh.tw2, !FM_FEWER_BARRIERS || job == 3 || (FM_TMA == 1 && job < 2));
      if (job < 2) {
        float ssq, samp, minsq;
        const int e = sh.ephase[h];
        __syncthreads();  // every thread has read its stage-3 inputs; the buffer can take the output
        store16_stats(v, sh.buf, first, &ssq, &samp, &minsq);
        if (SPLIT && job == 0 && b > 0) fm_split_wait(a.fm_seq + pairi, a.state, m, sh, wk);  // block b-1's state and ring are final
        fm_discriminate(a, sh, olen, h, c, b, ssq, samp, minsq, phase_from_index(a, e), h ? histB : histA,
                        (ringbase + first) & (NDEC - 1));
        if (t == 0) {
          int ep = e;
          for (int i = 0; i < split; i++) ep = phase_advance(ep, sh.P[h].phase_step, a.N);
          sh.ephase[h] = ep;
        }
      } else if (job == 2) {
        // Block-split form: the next block's discriminators may overwrite the ring from here on. Signalled only now, after
        // the forward transform has consumed the window: a barrier does not wait for loads still in flight, the
        // transform's first exchange through shared memory does (every value of the window has been used).
        if (SPLIT) fm_split_signal(a.fm_seq + pairi, a.state, m + 1, sh, wk);
        if (a.pl_spec) {  // PL-tone analyser enabled: the low bins of Z = FFT(audA + j audB) for pl_kernel (Z = conj(v))
          float2* o = a.pl_spec + ((long long)b * a.pl_npairs + pairi) * 65;
          if (t <= 32) o[t] = make_float2(v[0].x, -v[0].y);
          if (t >= 96) o[33 + 127 - t] = make_float2(v[15].x, -v[15].y);  // Z[2048 - k], k = 128 - t
        }
        const float2* R = a.audio_resp + (long long)sh.P[0].audio_slot * NDEC + t;
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float2 Z = make_float2(v[j].x, -v[j].y);  // undo the conj: Z = FFT-(z)
          v[j] = cmul(Z, __ldg(R + 128 * j));
        }
        fft2048_out_to_in(v);
      } else {
        const float gA = sh.P[0].fm_gain, gB = sh.P[1].fm_gain;
        int16_t* pa = pcm_row + sh.P[0].pcm_off - first;
        int16_t* pb = pcm_row + sh.P[1].pcm_off - first;
        const bool haveB = wk.y >= 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
          if (j >= jb) {  // warp-uniform; straight from the registers, no shared-memory round trip
            if ((j > jb) || (t >= rem)) {
              pa[t + 128 * j] = scaleclip(v[j].x * gA);  // fm.c:169-170
              if (haveB) pb[t + 128 * j] = scaleclip(v[j].y * gB);
            }
          }
        }
      }
    }
    if (!filtered)
      fm_flat_output(histA, histB, (ringbase + first) & (NDEC - 1), pcm_row + sh.P[0].pcm_off,
                     wk.y >= 0 ? pcm_row + sh.P[1].pcm_off : nullptr, olen);
    __syncthreads();
  }
  if (!SPLIT && t < 2) {  // (block-split form: the CTA of the last block stored the state when it signalled)
    const int c = t ? wk.y : wk.x;
    if (c >= 0) a.state[c] = sh.S[t];
  }
}

// ---------------------------------------------------------------- AM (envelope) and linear (SSB / CW / IQ / ISB)
//
// Both end in a strictly serial per-sample recurrence (hang AGC, and AM's carrier-DC tracker: am.c:60-74,
// linear.c:269-280) that must keep the reference's operation order. One lane can only retire ~1 sample per 25 cycles, so
// a CTA carries AGC_G channels: the parallel parts (response multiply, inverse FFT, amplitudes, quantisation) are done
// channel after channel by all 128 threads, then lane g of warp 0 runs channel g's recurrence — AGC_G serial loops in
// lockstep in ONE warp (the recurrences are written with selects, so the lanes never diverge). One lane per warp, as
// in the first version, costs a full issue slot per channel and instruction: ncu showed 610 M warp-instructions per
// launch at 8192 AM channels, 70 % of them single-lane (profiles/r01_am_kernel_a.txt).

// AGC_G channels per CTA (<= 4 = warps per CTA). Few channels: 1 per CTA minimises latency on a mostly idle GPU; many
// channels: 4 per CTA amortises the serial phases and maximises throughput. The launcher picks.

constexpr int AGC_ROW = 1024 + 1;  // floats per channel row: consecutive channels start one bank apart

template <int AGC_G>
struct AgcShared {
  float2 buf[NDEC];            // FFT exchange buffer
  float4 tw2[FFT2048_TW2_FLOAT4];  // stage-2 twiddles of the 2048-point transform
  float amp[AGC_G][AGC_ROW];   // amplitude (AM: envelope s[n]); rows padded: lane g reads row g in the serial loops
  float qg[AGC_G][AGC_ROW];    // attack value headroom/x[n] precomputed in parallel; overwritten by gain[n]
  float red[16];
  float scal[4][4];
  // dynamic tail: AM: dc[g][AGC_ROW] floats; linear: kept samples y[g][1024] float2
};

template <bool LINEAR, int AGC_G, int OLEN_T>
__global__ void __launch_bounds__(FFT2048_THREADS, AGC_G == 4 ? (LINEAR ? 2 : 3) : 4) agc_kernel(const ChanLaunch a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  AgcShared<AGC_G>& sh = *reinterpret_cast<AgcShared<AGC_G>*>(smraw);
  float2* ykeep = reinterpret_cast<float2*>(smraw + sizeof(AgcShared<AGC_G>));  // [AGC_G][1024], LINEAR only
  float* dcv = reinterpret_cast<float*>(smraw + sizeof(AgcShared<AGC_G>));      // [AGC_G][AGC_ROW], AM only
  const int t = threadIdx.x;
  fft2048_stage_tw2(sh.tw2, a.tw2048);
  __syncthreads();
  const int olen = OLEN_T ? OLEN_T : a.olen;
  const int first = NDEC - olen;
  const int jb = first >> 7, rem = first & 127;
  int chan[AGC_G];
#pragma unroll
  for (int g = 0; g < AGC_G; g++) {
    const int w = blockIdx.x * AGC_G + g;
    chan[g] = w < a.nwork ? a.work[w].x : -1;
  }
  // thread g < AGC_G owns channel g: its parameters, state and LO phase index stay in that thread's registers
  const int myw = blockIdx.x * AGC_G + t;
  const int myc = (t < AGC_G && myw < a.nwork) ? a.work[myw].x : -1;
  ChanParams P;
  ChanState S;
  int eph = 0;
  if (myc >= 0) {
    P = a.params[myc];
    S = a.state[myc];
    eph = phase_index0(P.bin, a.start0, a.N);
  }
  float2 v[16];
#pragma unroll 1
  for (int b = 0; b < a.nblocks; b++) {
    const float2* X = a.spec + (long long)b * a.spec_stride;
    // ---- parallel part, one channel at a time ----
#pragma unroll 1
    for (int g = 0; g < AGC_G; g++) {
      const int c = chan[g];
      if (c < 0) continue;
      const int bin = (int)a.params[c].bin;
      const bool isb = LINEAR && (a.params[c].flags & CH_ISB);
      if (isb) {  // CTA-uniform; the mirror-bin fold is staged through shared memory
        stage_filtered_isb(X, a.N, bin, a.resp + (long long)a.params[c].resp_slot * NDEC, sh.buf);
        __syncthreads();
        load16(v, sh.buf);
      } else {
        load_filtered16(v, X, a.N, bin, a.resp + (long long)a.params[c].resp_slot * NDEC);  // straight into the transform's registers
      }
      fft2048<+1>(v, sh.buf, a.tw2048, sh.tw2);
      // amplitudes (am.c:56-58, linear.c:256-261) and block power straight from the registers
      float sig = 0.f, noi = 0.f, dummy = 0.f;
      float* ampg = sh.amp[g];
      float2* yk = ykeep + g * 1024;
      const int ko = t - first;  // kept-sample index of row j is ko + 128 j
#pragma unroll
      for (int j = 0; j < 16; j++) {
        if (j >= jb) {  // warp-uniform
          if ((j > jb) || (t >= rem)) {
            const float rp = v[j].x * v[j].x, ip = v[j].y * v[j].y;
            sig += rp;
            noi += ip;
            ampg[ko + 128 * j] = sqrtf(rp + ip);
            if (LINEAR) yk[ko + 128 * j] = v[j];
          }
        }
      }
      if (a.filt_dbg) {
        __syncthreads();
        store16(v, sh.buf, jb);
        __syncthreads();
        const int e0 = phase_index0(a.params[c].bin, a.start0, a.N);
        int e = e0;
        for (int i = 0; i < b; i++) e = phase_advance(e, a.params[c].phase_step, a.N);
        dump_filter_output(a.filt_dbg + ((long long)b * a.nchan_total + c) * olen, sh.buf + first, olen,
                           phase_from_index(a, e));
      }
      block_reduce3<0>(sig, noi, dummy, sh.red);  // also orders the exchange buffer for the next channel
      if (t == 0) {
        sh.scal[g][0] = sig;
        sh.scal[g][1] = noi;
      }
    }
    __syncthreads();
    // ---- serial part, split so that only true dependences stay on the critical path ----
    // (1) AM: the carrier-DC tracker is its own recurrence, independent of the gain (am.c:60): one lane per channel,
    //     an FADD+FFMA chain. (2) The attack value headroom/x[n] (x = dc for AM, amplitude for linear) does not depend on
    //     the AGC state, so all threads compute those IEEE divisions in parallel. (3) The gain/hang recurrence itself
    //     (am.c:62-73, linear.c:269-279) is then compare + select per sample, same operands and order as the reference.
    if (!LINEAR) {
      if (myc >= 0) {
        const float* am = sh.amp[t];
        float* dco = dcv + t * AGC_ROW;
        float dc = S.am_dc;
        // explicit batches of 8: all loads first, the FADD+FFMA chain, then the stores (the arrays could alias for
        // the compiler, which otherwise serialises every load behind the previous store)
        int n = 0;
        for (; n + 8 <= olen; n += 8) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; i++) x[i] = am[n + i];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            dc += 0.0001f * (x[i] - dc);
            x[i] = dc;
          }
#pragma unroll
          for (int i = 0; i < 8; i++) dco[n + i] = x[i];
        }
        for (; n < olen; n++) {
          dc += 0.0001f * (am[n] - dc);
          dco[n] = dc;
        }
        S.am_dc = dc;
      }
      __syncthreads();
    }
#pragma unroll 1
    for (int g = 0; g < AGC_G; g++) {
      if (chan[g] < 0) continue;
      const float headroom = a.params[chan[g]].headroom;
      const float* xs = LINEAR ? sh.amp[g] : dcv + g * AGC_ROW;
      for (int o = t; o < olen; o += FFT2048_THREADS) sh.qg[g][o] = headroom / xs[o];
    }
    __syncthreads();
    if (myc >= 0) {
      const float* xs = LINEAR ? sh.amp[t] : dcv + t * AGC_ROW;
      float* qg = sh.qg[t];
      float gain = S.agc_gain;
      int hang = S.hang;
      const float headroom = P.headroom, rf = P.recovery_factor;
      const int hangmax = P.hangmax;
      auto agc_step = [&](float x, float q) {
        const bool startup = isnan(gain);                     // am.c:64 / linear.c:269: gain = headroom/x, hang untouched
        const bool over = !startup && (x * gain > headroom);  // attack: gain = headroom/x, hang = hangmax
        const bool hold = !startup && !over && hang != 0;
        const float grown = gain * rf;
        gain = (startup || over) ? q : (hold ? gain : grown);
        hang = over ? hangmax : (hold ? hang - 1 : hang);
        return gain;
      };
      int n = 0;
      for (; n + 8 <= olen; n += 8) {  // batches of 8: loads, chain, stores (see the DC tracker above)
        float x[8], q[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          x[i] = xs[n + i];
          q[i] = qg[n + i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = agc_step(x[i], q[i]);
#pragma unroll
        for (int i = 0; i < 8; i++) qg[n + i] = q[i];
      }
      for (; n < olen; n++) qg[n] = agc_step(xs[n], qg[n]);
      S.agc_gain = gain;
      S.hang = hang;
      ChanStatus st;
      const float sig = sh.scal[t][0], noi = sh.scal[t][1];
      st.bb_power = (sig + noi) / (2 * olen);  // am.c:78, linear.c:302
      st.snr = NAN;                            // linear.c:309 (no PLL)
      st.foffset = 0.f;
      st.pdeviation = 0.f;
      st.agc_gain = gain;
      st.squelch_open = 1;
      st.reserved[0] = LINEAR ? sig : S.am_dc;
      st.reserved[1] = LINEAR ? noi : 0.f;
      a.status[(long long)b * a.nchan_total + myc] = st;
    }
    __syncthreads();
    // ---- parallel output: gain / shift / quantise (linear.c:280-299, am.c:74, audio.c:22-28) ----
    // the owner threads hold the channels' phase indices; publish the per-channel scalars the output loop needs
    if (myc >= 0 && LINEAR) {
      const float2 ph = phase_from_index(a, eph);
      sh.scal[t][2] = ph.x;
      sh.scal[t][3] = ph.y;
    }
    if (LINEAR) __syncthreads();
#pragma unroll 1
    for (int g = 0; g < AGC_G; g++) {
      const int c = chan[g];
      if (c < 0) continue;
      int16_t* pcm_row = a.pcm + (long long)b * a.pcm_stride + a.params[c].pcm_off;
      if (!LINEAR) {
        const float* dcg = dcv + g * AGC_ROW;
        for (int o = t; o < olen; o += FFT2048_THREADS)
          pcm_row[o] = scaleclip((sh.amp[g][o] - dcg[o]) * sh.qg[g][o]);  // am.c:74
      } else {
        const float2 ph = make_float2(sh.scal[g][2], sh.scal[g][3]);
        const double shift_cycles = a.params[c].shift_cycles;
        const bool shifted = shift_cycles != 0.0;
        const double phase0 = shifted ? a.state[c].shift_phase + shift_cycles * (double)olen * b : 0.0;
        const int nch = a.params[c].channels;
        for (int o = t; o < olen; o += FFT2048_THREADS) {
          const float gn = sh.qg[g][o];
          const float2 y = cmul(ykeep[g * 1024 + o], ph);  // the block's LO phase rides the gain multiply
          float2 z = make_float2(y.x * gn, y.y * gn);      // linear.c:280
          if (shifted) {
            // post-detection shift oscillator (linear.c:283-289, osc.c:39-51): phasor(n) = exp(j*2*pi*f*n), n counted
            // from the first sample the oscillator was stepped on
            double p = phase0 + shift_cycles * (double)o;
            p -= floor(p);
            double sn, cs;
            sincospi(2.0 * p, &sn, &cs);
            z = cmul(z, make_float2((float)cs, (float)sn));
          }
          if (nch == 1) {
            pcm_row[o] = scaleclip(z.x);  // linear.c:291-296
          } else {
            pcm_row[2 * o] = scaleclip(z.x);  // I left, Q right (linear.c:299)
            pcm_row[2 * o + 1] = scaleclip(z.y);
          }
        }
      }
    }
    if (myc >= 0) eph = phase_advance(eph, P.phase_step, a.N);
    __syncthreads();
  }
  if (myc >= 0) {
    if (LINEAR && P.shift_cycles != 0.0) {
      double ph = S.shift_phase + P.shift_cycles * (double)olen * a.nblocks;
      S.shift_phase = ph - floor(ph);
    }
    a.state[myc] = S;
  }
}


// ---------------------------------------------------------------- AM / linear, split form (the default)
//
// The fused agc_kernel above keeps a whole CTA (and its 49-65 KB of shared memory) waiting while one warp walks the
// serial recurrences: at 8192 channels ncu shows 25 % issue utilisation, 12 resident warps per SM and the barrier as the
// top stall (profiles/r01_am_kernel_b.txt). The split form gives each part the shape it wants:
//   agc_front_kernel   one CTA per channel, like the FM predetection job: window x response -> 2048-point inverse FFT ->
//                      amplitudes (and, for linear, the kept samples) straight from registers to a global scratch row
//   agc_serial_kernel  one LANE per channel, 32 channels per warp in lockstep (the recurrences are select-based): tiles of
//                      32 samples are transposed through shared memory so every global access is a 128-byte row; the
//                      next tile is prefetched into registers while the current one is walked
//                      the store stage is also the output stage: AM (s - DC) * gain, linear gain x LO phase x shift
//                      oscillator, -> scaleclip -> PCM rows (the gains never go back to memory; the linear path's kept
//                      samples are prefetched into the store warp's registers one tile ahead)
// Same operations in the same order per channel as the fused kernel (bit-identical PCM), which stays selectable with
// KA9Q_B200_AGC_FUSED=1.

struct FrontShared {
  float2 buf[NDEC];
  float4 tw2[FFT2048_TW2_FLOAT4];  // stage-2 twiddles of the 2048-point transform
  float red[16];
};

template <bool LINEAR, int OLEN_T>
__global__ void __launch_bounds__(FFT2048_THREADS, 8) agc_front_kernel(const ChanLaunch a) {
  __shared__ FrontShared sh;
  const int t = threadIdx.x;
  const int w = blockIdx.x;
  const int c = a.work[w].x;
  const int olen = OLEN_T ? OLEN_T : a.olen;
  const int first = NDEC - olen;
  const int jb = first >> 7, rem = first & 127;
  const int bin = (int)a.params[c].bin;
  const bool isb = LINEAR && (a.params[c].flags & CH_ISB);
  const float2* H = a.resp + (long long)a.params[c].resp_slot * NDEC;
  float2 v[16];
  fft2048_stage_tw2(sh.tw2, a.tw2048);
  __syncthreads();
#pragma unroll 1
  for (int b = 0; b < a.nblocks; b++) {
    const float2* X = a.spec + (long long)b * a.spec_stride;
    if (isb) {  // CTA-uniform; the mirror-bin fold is staged through shared memory
      stage_filtered_isb(X, a.N, bin, H, sh.buf);
      __syncthreads();
      load16(v, sh.buf);
    } else {
      load_filtered16(v, X, a.N, bin, H);
    }
    fft2048<+1>(v, sh.buf, a.tw2048, sh.tw2);
    // amplitudes (am.c:56-58, linear.c:256-261) and block power straight from the registers
    const long long row = (long long)b * a.nwork + w;
    float* xr = a.agc_x + row * olen + (t - first);  // kept-sample index of row j is t - first + 128 j
    float2* yr = LINEAR ? a.agc_y + row * olen + (t - first) : nullptr;
    float sig = 0.f, noi = 0.f, dummy = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (j >= jb) {  // warp-uniform
        if ((j > jb) || (t >= rem)) {
          const float rp = v[j].x * v[j].x, ip = v[j].y * v[j].y;
          sig += rp;
          noi += ip;
          xr[128 * j] = sqrtf(rp + ip);
          if (LINEAR) yr[128 * j] = v[j];
        }
      }
    }
    if (a.filt_dbg) {
      __syncthreads();
      store16(v, sh.buf, jb);
      __syncthreads();
      const int e0 = phase_index0(a.params[c].bin, a.start0, a.N);
      int e = e0;
      for (int i = 0; i < b; i++) e = phase_advance(e, a.params[c].phase_step, a.N);
      dump_filter_output(a.filt_dbg + ((long long)b * a.nchan_total + c) * olen, sh.buf + first, olen,
                         phase_from_index(a, e));
    }
    block_reduce3<0>(sig, noi, dummy, sh.red);  // also orders the exchange buffer for the next block
    if (t == 0) {
      a.agc_pow[2 * row] = sig;
      a.agc_pow[2 * row + 1] = noi;
    }
  }
}

// One LANE per channel, 32 channels per CTA, and the per-sample work software-pipelined over eight warps so that no
// warp carries more than one short dependent chain or ~400 instructions per tile (a single warp doing everything needs
// ~41 instructions and 131 cycles per sample: measured 0.27 ms per 4 blocks however few channels there are). Tiles of
// 32 samples x 32 channels move through five stages (eight warps), one barrier per step, four buffers deep:
//   warp 0 (load)   tile s  : 32 coalesced 128-byte rows -> shared, transposed, by cp.async issued three steps ahead (the
//                             rows come from DRAM: one step is shorter than that latency)
//   warp 1 (dc)     tile s-1: AM carrier-DC tracker (am.c:60): FADD + FFMA chain                       [AM only]
//   warps 2-5 (div) tile s-2: attack values headroom / x, which do not depend on the AGC state. IEEE division compiles to
//                             a ~70-cycle sequence with a guarded slow path that does not interleave with its
//                             neighbours: 32 per lane in one warp made this the slowest stage (2800 cycles per step),
//                             so four warps take 8 samples of the tile each
//   warp 6 (agc)    tile s-3: the gain / hang recurrence (am.c:62-73, linear.c:269-279), compare + select only
//   warp 7 (store)  tile s-4: AM: (s - DC) * gain -> scaleclip -> int16 PCM rows; linear: gain rows back to the scratch
// grid = ceil(nwork / 32).
constexpr int SER_TP = 36;       // tile row pitch (floats): 16-byte aligned rows, 128-bit row accesses of the 32 lanes
                                 // (one row per lane, 144 bytes apart) are bank-conflict free per quarter-warp
constexpr int SER_THREADS = 256;  // warps: 0 load, 1 dc, 2-5 div (8 samples of the tile each), 6 agc, 7 store
constexpr int SER_THREADS_LIN = 320;  // linear: warps 1, 7, 8, 9 are the output stage (8 samples of the tile each)
struct SerialShared {
  float x[8][32 * SER_TP];  // amplitude tiles [channel][sample], filled by cp.async three steps ahead
  float q[4][32 * SER_TP];  // headroom / x, overwritten by the result (linear: gain; AM: (s - DC) * gain)
  float d[4][32 * SER_TP];  // AM: carrier level DC[n] (what the AGC follows)
  float o[4][32 * SER_TP];  // AM: s - DC
  int pcm_off[32];
};

__device__ __forceinline__ void ld8(float (&x)[8], const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}

template <bool LINEAR>
__global__ void __launch_bounds__(LINEAR ? SER_THREADS_LIN : SER_THREADS, 2) agc_serial_kernel(const ChanLaunch a) {
  extern __shared__ __align__(16) unsigned char ser_raw[];
  SerialShared& sh = *reinterpret_cast<SerialShared*>(ser_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // pipeline stage of this warp: 0 load, 1 dc, 2 div (warps 2..5, quarter `sub` of the tile each), 3 agc, 4 store
  // linear: the output stage is ~25 instructions per sample, far too much for one warp inside a pipeline step (measured:
  // 3.5 us per step with one output warp, 2 us with two, against 1.2 us for the recurrence). Four output warps — the idle
  // DC warp, the store warp and two extra warps (SER_THREADS_LIN = 320) — take eight samples of the tile each.
  const int role = (LINEAR && (warp == 1 || warp >= 7)) ? 4 : (warp < 2 ? warp : (warp < 6 ? 2 : warp - 3));
  const int ow = warp == 1 ? 0 : warp - 6;  // output warp 0..3 (warps 1, 7, 8, 9)
  const int sub = warp - 2;
  const int w0 = blockIdx.x * 32;
  const int nrows = min(32, a.nwork - w0);
  const int w = w0 + lane;
  const int c = lane < nrows ? a.work[w].x : -1;
  const int olen = a.olen;
  const int tpb = (olen + 31) >> 5;  // tiles per block
  const int ntiles = tpb * a.nblocks;
  // per-lane channel state, each field advanced (and finally stored) by one warp only
  float gain = 1.f, dc = 0.f, headroom = 1.f, rf = 1.f;
  int hang = 0, hangmax = 0;
  if (c >= 0) {
    const ChanParams& P = a.params[c];
    const ChanState& S = a.state[c];
    gain = S.agc_gain;
    hang = S.hang;
    dc = S.am_dc;
    headroom = P.headroom;
    rf = P.recovery_factor;
    hangmax = P.hangmax;
    if (role == 0) sh.pcm_off[lane] = P.pcm_off;
  } else if (role == 0) {
    sh.pcm_off[lane] = 0;
  }
  // linear output stage = the store warp, one lane per channel like every other stage: the channel's constants stay in the
  // lane's registers, its kept filter-output samples of the NEXT tile are prefetched into registers (32 x float2)
  int o_eph = 0, o_pstep = 0, o_nch = 1, o_off = 0;
  double o_shc = 0.0, o_shp0 = 0.0, o_shp = 0.0;
  float2 o_ph = make_float2(1.f, 0.f);
  float2 yv[8];                // this warp's eight samples of the next tile
  int pb = 0, pk = 0, pT = 0;  // block / tile-in-block / index of the next tile to prefetch
  if (LINEAR && role == 4 && c >= 0) {
    const ChanParams& P = a.params[c];
    o_eph = phase_index0(P.bin, a.start0, a.N);
    o_pstep = P.phase_step;
    o_shc = P.shift_cycles;
    o_shp0 = a.state[c].shift_phase;
    o_nch = P.channels;
    o_off = P.pcm_off;
  }
  // rows of tile (block b, tile k) in the scratch: this lane's column of row r is tile_ptr(b, k) + r*olen
  auto tile_ptr = [&](int b, int k) -> float* {
    return a.agc_x + ((long long)b * a.nwork + w0) * olen + 32 * k + lane;
  };
  // asynchronous transposed copy of a full tile (partial tiles are filled synchronously in their own step); always one
  // commit group per call so the group count stays uniform. (ib, ik) walk the tiles in order: no division per step.
  int ib = 0, ik = 0, iT = 0;
  auto issue_next_tile = [&]() {
    if (iT < ntiles) {
      const int cnt = min(32, olen - 32 * ik);
      if (nrows == 32 && cnt == 32) {
        const float* g = tile_ptr(ib, ik);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sh.x[iT & 7] + lane);
#pragma unroll
        for (int r = 0; r < 32; r++) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + 4u * (unsigned)(r * SER_TP)), "l"(g) : "memory");
          g += olen;
        }
      }
      iT++;
      if (++ik == tpb) {
        ik = 0;
        ib++;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  if (role == 0) {
    issue_next_tile();
    issue_next_tile();
    issue_next_tile();
  }
  __syncthreads();
  int b = 0, k = 0;  // block and tile-in-block of the tile this warp works on (advanced after every live step)
#pragma unroll 1
  for (int s = 0; s < ntiles + 4; s++) {
    const int T = s - role;  // the tile this warp works on in this step
    const bool live = T >= 0 && T < ntiles;
    const int cnt = min(32, olen - 32 * k);
    const int bi = T & 3;
    if (role == 0) {
      issue_next_tile();
      asm volatile("cp.async.wait_group 3;\n" ::: "memory");  // tile T has landed
      if (live && !(nrows == 32 && cnt == 32)) {  // edge tile: plain loads, the unused slots get a harmless 1.0
        const float* g = tile_ptr(b, k);
        float* xt = sh.x[T & 7];
#pragma unroll 4
        for (int r = 0; r < 32; r++) xt[r * SER_TP + lane] = (r < nrows && lane < cnt) ? g[(long long)r * olen] : 1.f;
      }
    } else if (role == 1) {
      if (live && !LINEAR) {
        const float* xt = sh.x[T & 7] + lane * SER_TP;
        float* dt = sh.d[bi] + lane * SER_TP;
        float* ot = sh.o[bi] + lane * SER_TP;
        int i0 = 0;
#pragma unroll 1
        for (; i0 + 8 <= cnt; i0 += 8) {  // batches of 8: loads, chain, stores
          float x[8], dd[8];
          ld8(x, xt + i0);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            dc += 0.0001f * (x[i] - dc);  // am.c:60
            dd[i] = dc;                   // the AGC follows the carrier level (am.c:62)
            x[i] = x[i] - dc;             // s - DC (am.c:74)
          }
          st8(dt + i0, dd);
          st8(ot + i0, x);
        }
#pragma unroll 1
        for (; i0 < cnt; i0++) {  // olen not a multiple of 8
          const float x = xt[i0];
          dc += 0.0001f * (x - dc);
          dt[i0] = dc;
          ot[i0] = x - dc;
        }
      }
    } else if (role == 2) {
      if (live) {
        const float* xt = (LINEAR ? sh.x[T & 7] : sh.d[bi]) + lane * SER_TP + 8 * sub;
        float* qt = sh.q[bi] + lane * SER_TP + 8 * sub;
        const int m = min(8, cnt - 8 * sub);  // samples of this quarter inside the tile
        if (m == 8) {
          float x[8];
          ld8(x, xt);
#pragma unroll
          for (int i = 0; i < 8; i++) x[i] = headroom / x[i];
          st8(qt, x);
        } else {
#pragma unroll 1
          for (int i = 0; i < m; i++) qt[i] = headroom / xt[i];
        }
      }
    } else if (role == 3) {
      if (live) {
        const float* xt = (LINEAR ? sh.x[T & 7] : sh.d[bi]) + lane * SER_TP;
        const float* ot = sh.o[bi] + lane * SER_TP;
        float* qt = sh.q[bi] + lane * SER_TP;
        // am.c:62-73 / linear.c:269-279, same operands and order as the reference, written with selects. The start-up
        // test (gain still NaN: am.c:64, linear.c:269) is hoisted out: a gain that is a number stays one.
        const bool any_startup = __any_sync(0xffffffffu, isnan(gain));
        auto agc_step_general = [&](float x, float q) {
          const bool startup = isnan(gain);                     // gain = headroom/x, hang untouched
          const bool over = !startup && (x * gain > headroom);  // attack: gain = headroom/x, hang = hangmax
          const bool hold = !startup && !over && hang != 0;
          const float grown = gain * rf;
          gain = (startup || over) ? q : (hold ? gain : grown);
          hang = over ? hangmax : (hold ? hang - 1 : hang);
          return gain;
        };
        auto agc_step = [&](float x, float q) {
          const bool over = x * gain > headroom;
          const float kept = hang != 0 ? gain : gain * rf;  // hold, or recover
          gain = over ? q : kept;
          hang = over ? hangmax : (hang != 0 ? hang - 1 : 0);  // am.c:68-69 / linear.c:275-276, also for hangmax < 0
          return gain;
        };
        int i0 = 0;
        if (!any_startup) {
#pragma unroll 1
          for (; i0 + 8 <= cnt; i0 += 8) {
            float x[8], q[8], o[8];
            ld8(x, xt + i0);
            ld8(q, qt + i0);
            if (!LINEAR) ld8(o, ot + i0);
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = agc_step(x[i], q[i]);
            if (!LINEAR) {
#pragma unroll
              for (int i = 0; i < 8; i++) q[i] = o[i] * q[i];
            }
            st8(qt + i0, q);
          }
        }
#pragma unroll 1
        for (; i0 < cnt; i0++) {  // the stream's first tile, or olen not a multiple of 8
          const float gn = agc_step_general(xt[i0], qt[i0]);
          qt[i0] = LINEAR ? gn : ot[i0] * gn;
        }
        if (k == tpb - 1 && c >= 0) {  // end of block b: status row (the DC of the block's last sample is in the tile)
          const long long row = (long long)b * a.nwork + w;
          const float sig = a.agc_pow[2 * row], noi = a.agc_pow[2 * row + 1];
          ChanStatus st;
          st.bb_power = (sig + noi) / (2 * olen);  // am.c:78, linear.c:302
          st.snr = NAN;                            // linear.c:309 (no PLL)
          st.foffset = 0.f;
          st.pdeviation = 0.f;
          st.agc_gain = gain;
          st.squelch_open = 1;
          st.reserved[0] = LINEAR ? sig : xt[cnt - 1];
          st.reserved[1] = LINEAR ? noi : 0.f;
          a.status[(long long)b * a.nchan_total + c] = st;
        }
      }
    } else {
      if (live) {
        const float* qt = sh.q[bi] + lane;
        if (LINEAR) {
          // linear.c:280-299 + audio.c:22-28 straight from the gains just computed (same operations and order as the
          // fused kernel): the gain rows never go back to memory and no third kernel runs. Lane = channel.
          if (k == 0) {  // first tile of block b: the block's LO phase and shift-oscillator phase
            o_ph = phase_from_index(a, o_eph);
            o_eph = phase_advance(o_eph, o_pstep, a.N);
            o_shp = o_shc != 0.0 ? o_shp0 + o_shc * (double)olen * b : 0.0;
          }
          if (c >= 0) {
            const float* qr = sh.q[bi] + lane * SER_TP;
            int16_t* row = a.pcm + (long long)b * a.pcm_stride + o_off + (long long)o_nch * 32 * k;
            {
              const int h0 = 0;
              const int i0 = 8 * ow;  // first sample of this warp's group inside the tile
              if (i0 < cnt) {  // warp-uniform
                float gn[8];
                ld8(gn, qr + i0);
                int16_t out[16];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                  const float2 y = cmul(yv[h0 + i], o_ph);             // the block's LO phase rides the gain multiply
                  float2 z = make_float2(y.x * gn[i], y.y * gn[i]);  // linear.c:280
                  if (o_shc != 0.0) {  // post-detection shift oscillator (linear.c:283-289, osc.c:39-51)
                    double ps = o_shp + o_shc * (double)(32 * k + i0 + i);
                    ps -= floor(ps);
                    double sn, cs;
                    sincospi(2.0 * ps, &sn, &cs);
                    z = cmul(z, make_float2((float)cs, (float)sn));
                  }
                  out[2 * i] = scaleclip(z.x);
                  out[2 * i + 1] = scaleclip(z.y);
                }
                if (i0 + 8 <= cnt && (olen & 7) == 0) {  // whole group inside the block, rows 16-byte aligned: 128-bit stores
                  if (o_nch == 1) {   // mono: I only (linear.c:291-296)
                    uint4 w;
                    w.x = (unsigned short)out[0] | ((unsigned)(unsigned short)out[2] << 16);
                    w.y = (unsigned short)out[4] | ((unsigned)(unsigned short)out[6] << 16);
                    w.z = (unsigned short)out[8] | ((unsigned)(unsigned short)out[10] << 16);
                    w.w = (unsigned short)out[12] | ((unsigned)(unsigned short)out[14] << 16);
                    *reinterpret_cast<uint4*>(row + i0) = w;
                  } else {  // I left, Q right (linear.c:299)
                    uint4 w0, w1;
                    w0.x = (unsigned short)out[0] | ((unsigned)(unsigned short)out[1] << 16);
                    w0.y = (unsigned short)out[2] | ((unsigned)(unsigned short)out[3] << 16);
                    w0.z = (unsigned short)out[4] | ((unsigned)(unsigned short)out[5] << 16);
                    w0.w = (unsigned short)out[6] | ((unsigned)(unsigned short)out[7] << 16);
                    w1.x = (unsigned short)out[8] | ((unsigned)(unsigned short)out[9] << 16);
                    w1.y = (unsigned short)out[10] | ((unsigned)(unsigned short)out[11] << 16);
                    w1.z = (unsigned short)out[12] | ((unsigned)(unsigned short)out[13] << 16);
                    w1.w = (unsigned short)out[14] | ((unsigned)(unsigned short)out[15] << 16);
                    *reinterpret_cast<uint4*>(row + 2 * i0) = w0;
                    *reinterpret_cast<uint4*>(row + 2 * i0 + 8) = w1;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; i++)
                    if (i0 + i < cnt) {
                      if (o_nch == 1) {
                        row[i0 + i] = out[2 * i];
                      } else {
                        row[2 * (i0 + i)] = out[2 * i];
                        row[2 * (i0 + i) + 1] = out[2 * i + 1];
                      }
                    }
                }
              }
            }
          }
        } else {
          // am.c:74 + audio.c:22-28: row r is channel w0 + r
          int16_t* pcm_blk = a.pcm + (long long)b * a.pcm_stride + 32 * k + lane;
#pragma unroll 8
          for (int r = 0; r < nrows; r++)
            if (lane < cnt) pcm_blk[sh.pcm_off[r]] = scaleclip(qt[r * SER_TP]);
        }
      }
    }
    if (LINEAR && role == 4 && T + 1 >= 0 && T + 1 < ntiles && pT == T + 1) {
      // this lane's channel: this warp's eight kept filter-output samples of the tile handled NEXT step, as four 128-bit
      // loads (64 contiguous bytes per lane; the guard keeps the last tile of a block inside its row)
      if (c >= 0) {
        const float4* g = reinterpret_cast<const float4*>(a.agc_y + ((long long)pb * a.nwork + w) * olen + 32 * pk + 8 * ow);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float4 v = (32 * pk + 8 * ow + 2 * j + 1 < olen) ? __ldg(g + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          yv[2 * j] = make_float2(v.x, v.y);
          yv[2 * j + 1] = make_float2(v.z, v.w);
        }
      }
      pT++;
      if (++pk == tpb) {
        pk = 0;
        pb++;
      }
    }
    if (live && ++k == tpb) {
      k = 0;
      b++;
    }
    __syncthreads();
  }
  if (c >= 0) {
    // every field is written by the warp that advanced it (agc_shift_advance_kernel advances the shift oscillator)
    if (role == 1 && !LINEAR) a.state[c].am_dc = dc;
    if (role == 3) {
      a.state[c].agc_gain = gain;
      a.state[c].hang = hang;
    }
  }
}

// after the recurrence kernel: the shift oscillators have run olen * nblocks samples further
__global__ void agc_shift_advance_kernel(const ChanLaunch a) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.nwork) return;
  const int c = a.work[w].x;
  const double shift_cycles = a.params[c].shift_cycles;
  if (shift_cycles != 0.0) {
    const double ph = a.state[c].shift_phase + shift_cycles * (double)a.olen * a.nblocks;
    a.state[c].shift_phase = ph - floor(ph);
  }
}

}  // namespace k9
#include "pll_kernel.cuh"
#include "pl_kernel.cuh"
namespace k9 {

// ---------------------------------------------------------------- launchers

int fm_carveout(bool mixed) { return mixed ? (int)cudaSharedmemCarveoutMaxShared : FM_CARVEOUT_PCT; }

int launch_fm(const ChanLaunch& a, cudaStream_t st, bool mixed) {
  if (a.nwork <= 0) return 0;
  // 8 CTAs x (16.8 KB + 1 KB reserved) = 143 KB: alone, ask for the 164 KB carve-out (percent of 228 KB, rounded up by
  // the driver to the next supported size) and keep ~90 KB of L1 (measured: 228 -> 196 -> 164 KB = 0.528 -> 0.495 ->
  // 0.491 ms per launch at cfg5); beside AM / linear kernels use their (maximum) carve-out so the CTAs can share SMs.
  // (function attributes are per device: one slot per device, so several GPUs driven from one process all get set up)
  static int configured[64];
  const int pct = fm_carveout(mixed);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (configured[dev] != pct + 1000) {
    cudaFuncSetAttribute(fm_kernel<960, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fm_kernel<0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fm_kernel<960, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fm_kernel<0, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    configured[dev] = pct + 1000;
  }
  // Few pairs: split every pair's blocks over a cluster of 2 (or 4) CTAs (fm_split_wait).
  static int sms[64];
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  const char* ev = getenv("KA9Q_B200_FM_SPLIT");  // 1 / 2 / 4 force the form (tests, A/B); read per launch
  const int forced = ev ? atoi(ev) : -1;
  // Measured (scripts/gpu_fm_split_probe.py, cfg5 stream, 8 blocks): 32 pairs 0.0625 -> 0.0534 ms with 2 CTAs per pair,
  // 0.0529 with 4; 256 pairs 0.1065 -> 0.1043 / 0.1137; 512 pairs 0.1352 -> 0.1330 / 0.1689. The blocks of a pair stay
  // chained through both discriminators and the ring loads (more than half of a pair-block at single-CTA latency), so
  // the form pays only while most SMs would otherwise hold one CTA or none: 2 CTAs per pair up to two pairs per SM.
  int split = 1;
  if (a.fm_seq && a.nblocks >= 2) {
    if (a.nwork <= 2 * sms[dev]) split = 2;
    if (forced == 1 || forced == 2 || (forced == 4 && a.nblocks >= 4)) split = forced;
  }
  if (split == 1) {
    if (a.olen == 960)
      fm_kernel<960, false><<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
    else
      fm_kernel<0, false><<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  ChanLaunch as = a;
  as.fm_split = split;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.nwork * split);
  cfg.blockDim = dim3(FFT2048_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = split;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const cudaError_t e = a.olen == 960 ? cudaLaunchKernelEx(&cfg, fm_kernel<960, true>, as)
                                      : cudaLaunchKernelEx(&cfg, fm_kernel<0, true>, as);
  return e == cudaSuccess ? 0 : -1;
}
template <bool LINEAR, int G>
static int launch_agc_g(const ChanLaunch& a, cudaStream_t st) {
  const size_t smem = sizeof(AgcShared<G>) + (LINEAR ? sizeof(float2) * 1024 : sizeof(float) * AGC_ROW) * G;
  static bool configured_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  bool& configured = configured_dev[dev & 63];
  if (!configured) {
    cudaFuncSetAttribute(agc_kernel<LINEAR, G, 960>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(agc_kernel<LINEAR, G, 960>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(agc_kernel<LINEAR, G, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(agc_kernel<LINEAR, G, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  if (a.olen == 960)
    agc_kernel<LINEAR, G, 960><<<(a.nwork + G - 1) / G, FFT2048_THREADS, smem, st>>>(a);
  else
    agc_kernel<LINEAR, G, 0><<<(a.nwork + G - 1) / G, FFT2048_THREADS, smem, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
template <bool LINEAR>
static int launch_agc_fused(const ChanLaunch& a, cudaStream_t st) {
  // one channel per CTA until the GPU (148 SMs x ~3 CTAs) is full, then pack to amortise the serial phases
  if (a.nwork >= 148 * 3 * 2) return launch_agc_g<LINEAR, 4>(a, st);
  if (a.nwork >= 148 * 3) return launch_agc_g<LINEAR, 2>(a, st);
  return launch_agc_g<LINEAR, 1>(a, st);
}
template <bool LINEAR>
static int launch_agc(const ChanLaunch& a, cudaStream_t st) {
  if (a.nwork <= 0) return 0;
  const char* env = getenv("KA9Q_B200_AGC_FUSED");  // read per launch so a test can compare the two forms in one process
  const bool fused = env && atoi(env) != 0;
  // the recurrence kernel takes ~0.15 ms per 4 blocks however few channels there are (its pipeline is latency-bound);
  // below ~1000 channels the fused kernel with one channel per CTA is as fast or faster
  if (fused || !a.agc_x || (a.nwork < 1024 && !(env && atoi(env) == 0))) return launch_agc_fused<LINEAR>(a, st);
  // (Tried: block by block on two streams, the front kernel of block b+1 beside the recurrence of block b. Measured
  // slower, 0.374 vs 0.351 ms at 8192 AM channels: the front kernel fills every SM, so the recurrence CTAs only get in
  // once it drains, and the per-block launches add their own fill / drain.)
  {
    static bool configured_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured_dev[dev & 63]) {
      cudaFuncSetAttribute(agc_serial_kernel<LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SerialShared));
      configured_dev[dev & 63] = true;
    }
  }
  if (a.olen == 960)
    agc_front_kernel<LINEAR, 960><<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
  else
    agc_front_kernel<LINEAR, 0><<<a.nwork, FFT2048_THREADS, 0, st>>>(a);
  agc_serial_kernel<LINEAR><<<(a.nwork + 31) / 32, LINEAR ? SER_THREADS_LIN : SER_THREADS, sizeof(SerialShared), st>>>(a);
  // the recurrence kernel wrote the PCM (AM: (s - DC) * gain; linear: gain x LO phase x shift oscillator); only the
  // shift oscillators' phase is left to advance
  if (LINEAR) agc_shift_advance_kernel<<<(a.nwork + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_am(const ChanLaunch& a, cudaStream_t st) { return launch_agc<false>(a, st); }
int launch_linear(const ChanLaunch& a, cudaStream_t st) { return launch_agc<true>(a, st); }

}  // namespace k9
