// Front-end decimator service on the device: what the reference's `hackrf` daemon does between the A/D
// and the I/Q multicast, device resident, so that the decimated int16 stream can be written straight into the
// channelizer's device ring.
//
//   rx_callback (hackrf.c:129-196)  per USB transfer ("callback block"): int8 I/Q (-128 counted as a clip and taken as
//       -127) x 1/127, minus the DC estimate, I / Q gain balance, phase correction; at the END of every callback block
//       the DC, imbalance and sin(phi) estimates are advanced from that block's sums and new gains derived
//   process (hackrf.c:198-345)      + Fs/4 rotation by the running sample count (hackrf.c:271-291), cascade of log2(D)
//       15-tap half-band decimators on each plane (decimate.c:44-147, stage D-1 first ... stage 0 last), x 0.5^stages,
//       (short)round(32767 s)  (hackrf.c:297-328, 469)
//
// Two passes over the int8 input (2 bytes per complex sample each), nothing else touches HBM but the int16 output:
//   fe_stats_kernel   one launch per callback block: the block's sums with the coefficients in force, reduced in a fixed
//                     order by the last CTA to finish, which also advances the estimates (the coefficient chain is
//                     sequential by construction: block k+1 is corrected with what block k produced)
//   fe_cascade_kernel the whole batch, both planes: correction + rotation fused into the load of the cascade's first level,
//                     all levels in shared memory with the halo recomputed per CTA (as hb15_cascade_kernel, decimate.cu),
//                     attenuation and rounding fused into the store
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/ka9q_b200.h"
#include "util.cuh"

using namespace k9;

namespace {

struct FeCoef {  // what rx_callback applies to one callback block
  float dc_re, dc_im, gain_i, gain_q, secphi, tanphi;
};
struct FeEst {   // HackCD.* estimates carried between callback blocks (hackrf.c:182-194)
  float dc_re, dc_im, imbalance, sinphi, in_power;
  unsigned long long clips;
};

constexpr int FE_STATS_CTAS = 64;
constexpr int FE_THREADS = 256;
constexpr int FE_MAXS = 6;
constexpr int FE_C = 64;  // final outputs per CTA
constexpr int FE_BUF = (FE_C << FE_MAXS) + 13 * ((1 << FE_MAXS) - 1) + 8;

__device__ __forceinline__ float2 fe_correct(char2 raw, const FeCoef& k, int* clips) {
  int qi = raw.x, qq = raw.y;
  if (qq == -128) {
    (*clips)++;
    qq = -127;
  }
  if (qi == -128) {
    (*clips)++;
    qi = -127;
  }
  float re = (float)qi * (float)(1. / 127), im = (float)qq * (float)(1. / 127);  // SCALE8 (hackrf.c:155)
  re -= k.dc_re;
  im -= k.dc_im;
  re *= k.gain_i;
  im *= k.gain_q;
  im = k.secphi * im - k.tanphi * re;  // hackrf.c:174
  return make_float2(re, im);
}

struct StatsPartial {
  float sum_re, sum_im, i_energy, q_energy, dotprod;
  int clips;
};

__global__ void __launch_bounds__(FE_THREADS) fe_stats_kernel(const char2* __restrict__ raw, int n, int blk, FeCoef* coef /* [nblk+1] */,
                                                              FeEst* est, StatsPartial* partial, unsigned* counter,
                                                              float dc_alpha, float rate_factor) {
  const FeCoef k = coef[blk];
  float sre = 0, sim = 0, ie = 0, qe = 0, dp = 0;
  int clips = 0;
  for (int i = blockIdx.x * FE_THREADS + threadIdx.x; i < n; i += gridDim.x * FE_THREADS) {
    const char2 r = raw[i];
    int qi = r.x, qq = r.y;
    if (qq == -128) {
      clips++;
      qq = -127;
    }
    if (qi == -128) {
      clips++;
      qi = -127;
    }
    float re = (float)qi * (float)(1. / 127), im = (float)qq * (float)(1. / 127);
    sre += re;  // samp_sum is taken before the DC is removed (hackrf.c:157)
    sim += im;
    re -= k.dc_re;
    im -= k.dc_im;
    ie += re * re;  // energies before the gain correction (hackrf.c:164-165)
    qe += im * im;
    re *= k.gain_i;
    im *= k.gain_q;
    dp += re * im;  // phase error before the phase correction (hackrf.c:172)
  }
  __shared__ float red[5][FE_THREADS / 32];
  __shared__ int redc[FE_THREADS / 32];
  __shared__ bool last;
  sre = warp_sum(sre);
  sim = warp_sum(sim);
  ie = warp_sum(ie);
  qe = warp_sum(qe);
  dp = warp_sum(dp);
  for (int o = 16; o > 0; o >>= 1) clips += __shfl_xor_sync(0xffffffffu, clips, o);
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    red[0][w] = sre;
    red[1][w] = sim;
    red[2][w] = ie;
    red[3][w] = qe;
    red[4][w] = dp;
    redc[w] = clips;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    StatsPartial p = {0, 0, 0, 0, 0, 0};
    for (int w = 0; w < FE_THREADS / 32; w++) {
      p.sum_re += red[0][w];
      p.sum_im += red[1][w];
      p.i_energy += red[2][w];
      p.q_energy += red[3][w];
      p.dotprod += red[4][w];
      p.clips += redc[w];
    }
    partial[blockIdx.x] = p;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    // the last CTA to finish reduces the per-CTA partials in a fixed order (lane l takes CTAs l, l + 32, ...; then a
    // shuffle tree): deterministic, and one warp instead of one thread walking 64 dependent L2 loads
    __threadfence();
    StatsPartial s = {0, 0, 0, 0, 0, 0};
    for (unsigned c = threadIdx.x; c < gridDim.x; c += 32) {
      const StatsPartial* q = partial + c;
      s.sum_re += __ldcg(&q->sum_re);
      s.sum_im += __ldcg(&q->sum_im);
      s.i_energy += __ldcg(&q->i_energy);
      s.q_energy += __ldcg(&q->q_energy);
      s.dotprod += __ldcg(&q->dotprod);
      s.clips += __ldcg(&q->clips);
    }
    s.sum_re = warp_sum(s.sum_re);
    s.sum_im = warp_sum(s.sum_im);
    s.i_energy = warp_sum(s.i_energy);
    s.q_energy = warp_sum(s.q_energy);
    s.dotprod = warp_sum(s.dotprod);
    for (int o = 16; o > 0; o >>= 1) s.clips += __shfl_xor_sync(0xffffffffu, s.clips, o);
    if (threadIdx.x != 0) return;
    *counter = 0;
    FeEst e = *est;
    FeCoef nk = k;
    // hackrf.c:182-194, in its order
    e.dc_re += dc_alpha * (s.sum_re - n * e.dc_re);
    e.dc_im += dc_alpha * (s.sum_im - n * e.dc_im);
    const float block_energy = 0.5f * (s.i_energy + s.q_energy);
    if (block_energy > 0) {
      e.in_power = block_energy / n;
      e.imbalance += rate_factor * n * ((s.i_energy / s.q_energy) - e.imbalance);
      const float dpn = s.dotprod / block_energy;
      e.sinphi += rate_factor * n * (dpn - e.sinphi);
      nk.gain_q = sqrtf(0.5f * (1 + e.imbalance));
      nk.gain_i = sqrtf((float)(0.5 * (1 + 1. / e.imbalance)));
      nk.secphi = 1 / sqrtf(1 - e.sinphi * e.sinphi);
      nk.tanphi = e.sinphi * nk.secphi;
    }
    e.clips += (unsigned long long)s.clips;
    nk.dc_re = e.dc_re;
    nk.dc_im = e.dc_im;
    *est = e;
    coef[blk + 1] = nk;
  }
}

struct CascadeArgs {
  const char2* raw;     // [n_in] int8 I/Q of the batch
  const FeCoef* coef;   // [nblk] coefficients of each callback block
  int cb_samples;       // samples per callback block
  long long sample0;    // stream index of raw[0] (Fs/4 rotation phase)
  int offset;           // rotation steps per sample (hackrf.c:68: 1 = tuner high by Fs/4)
  const float* hist;    // [2][S][16]: hist[p][l][k] = x_l[-k]
  float* tail;          // [2][S][16]
  float4 coeff;         // half-band coefficients c0..c3 (hackrf.c:229-237), the same at every level
  int S, n_in;
  float atten;          // 0.5^S (hackrf.c:469)
  short2* out;          // [n_in >> S] int16 I/Q
  float* out_energy;    // [1] sum s^2 over both planes (HackCD.out_power numerator, hackrf.c:308,325), may be null
};

// grid = (ceil(n_out / FE_C)), both planes in one CTA (the int8 pair is loaded once)
__global__ void __launch_bounds__(FE_THREADS) fe_cascade_kernel(const CascadeArgs a) {
  extern __shared__ float fe_sm[];
  float* bufp[2][2] = {{fe_sm, fe_sm + FE_BUF}, {fe_sm + 2 * FE_BUF, fe_sm + 3 * FE_BUF}};  // [plane][ping-pong]
  float(*hs)[FE_MAXS][16] = reinterpret_cast<float(*)[FE_MAXS][16]>(fe_sm + 4 * FE_BUF);
  const int t = threadIdx.x;
  const int S = a.S;
  for (int i = t; i < 2 * S * 16; i += FE_THREADS) hs[i / (16 * S)][(i / 16) % S][i & 15] = a.hist[i];
  const int n_out = a.n_in >> S;
  const int m0 = blockIdx.x * FE_C, m1 = min(m0 + FE_C, n_out);
  const bool last_cta = m1 == n_out;
  int lo[FE_MAXS + 1], hi[FE_MAXS + 1];
  lo[S] = m0;
  hi[S] = m1;
  for (int l = S - 1; l >= 0; l--) {
    lo[l] = 2 * lo[l + 1] - 13;
    hi[l] = 2 * hi[l + 1];
  }
  // level 0: correction (rx_callback) + Fs/4 rotation (process) fused into the load
  int clips = 0;
  for (int i = max(lo[0], 0) + t; i < hi[0]; i += FE_THREADS) {
    const float2 s = fe_correct(a.raw[i], a.coef[i / a.cb_samples], &clips);
    float re, im;
    switch ((int)(((a.sample0 + i) * a.offset) & 3)) {  // hackrf.c:273-289
      default:
      case 0: re = s.x; im = s.y; break;
      case 1: re = -s.y; im = s.x; break;
      case 2: re = -s.x; im = -s.y; break;
      case 3: re = s.y; im = -s.x; break;
    }
    bufp[0][0][i - lo[0]] = re;
    bufp[1][0][i - lo[0]] = im;
  }
  __syncthreads();
  const float4 c = a.coeff;
  int cur = 0;
  float e = 0.f;
  for (int l = 0; l < S; l++) {
    const int lo_in = lo[l];
    const bool final_level = l == S - 1;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const float* in = bufp[p][cur];
      float* out = bufp[p][cur ^ 1];
      const float* h = hs[p][l];
      auto X = [&](int i) -> float { return i >= 0 ? in[i - lo_in] : h[-i]; };
      if (last_cta && t >= 1 && t <= 13) {  // new history of this level: the last 13 samples of (history ++ input)
        const int n_l = a.n_in >> l;
        const int i = n_l - t;
        a.tail[(p * S + l) * 16 + t] = i >= 0 ? in[i - lo_in] : h[t - n_l];
      }
      for (int m = max(lo[l + 1], 0) + t; m < hi[l + 1]; m += FE_THREADS) {
        const int b = 2 * m;
        // same association as the portable reference loop (decimate.c:124-128)
        float r = X(b - 6);
        r += (X(b + 1) + X(b - 13)) * c.x;
        r += (X(b - 1) + X(b - 11)) * c.y;
        r += (X(b - 3) + X(b - 9)) * c.z;
        r += (X(b - 5) + X(b - 7)) * c.w;
        if (final_level) {
          const float s = r * a.atten;  // hackrf.c:307
          e += s * s;
          reinterpret_cast<short*>(a.out + m)[p] = (short)roundf(32767 * s);  // (short)round(32767 * s), hackrf.c:309
        } else {
          out[m - lo[l + 1]] = r;
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  if (a.out_energy) {
    e = warp_sum(e);
    if ((t & 31) == 0 && e != 0.f) atomicAdd(a.out_energy, e);
  }
}

}  // namespace

struct ka9q_frontend {
  ka9q_frontend_config cfg;
  int S = 0;
  cudaStream_t st = nullptr;
  char2* d_raw = nullptr;
  size_t raw_cap = 0;
  FeCoef* d_coef = nullptr;
  int coef_cap = 0;
  FeEst* d_est = nullptr;
  StatsPartial* d_partial = nullptr;
  unsigned* d_counter = nullptr;
  float *d_hist = nullptr, *d_tail = nullptr, *d_energy = nullptr;
  short2* d_out = nullptr;
  size_t out_cap = 0;
  long long samples = 0;  // input samples consumed so far
  FeCoef h_coef;          // coefficients in force for the next callback block (host mirror, refreshed on demand)
};

extern "C" {

int ka9q_frontend_create(ka9q_frontend** out, const ka9q_frontend_config* cfg) {
  if (!out || !cfg) {
    set_error("null argument");
    return -1;
  }
  *out = nullptr;
  int S = 0;
  while ((1 << S) < cfg->decimate) S++;
  if ((1 << S) != cfg->decimate || S < 1 || S > FE_MAXS) {
    set_error("decimation ratio must be a power of two, 2..64 (hackrf.c:464-467)");
    return -1;
  }
  if (cfg->callback_samples <= 0 || cfg->callback_samples % cfg->decimate != 0 || cfg->out_samprate <= 0) {
    set_error("callback_samples must be a positive multiple of the decimation ratio");
    return -1;
  }
  if (ka9q_device_count() <= cfg->device || cfg->device < 0) {
    set_error("CUDA device %d not available (no CPU fallback exists)", cfg->device);
    return -1;
  }
  K9_CUDA(cudaSetDevice(cfg->device));
  ka9q_frontend* f = new ka9q_frontend();
  f->cfg = *cfg;
  f->S = S;
  K9_CUDA(cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking));
  K9_CUDA(cudaMalloc(&f->d_est, sizeof(FeEst)));
  K9_CUDA(cudaMalloc(&f->d_partial, sizeof(StatsPartial) * FE_STATS_CTAS));
  K9_CUDA(cudaMalloc(&f->d_counter, sizeof(unsigned)));
  K9_CUDA(cudaMalloc(&f->d_hist, sizeof(float) * 2 * FE_MAXS * 16));
  K9_CUDA(cudaMalloc(&f->d_tail, sizeof(float) * 2 * FE_MAXS * 16));
  K9_CUDA(cudaMalloc(&f->d_energy, sizeof(float)));
  K9_CUDA(cudaMemset(f->d_counter, 0, sizeof(unsigned)));
  K9_CUDA(cudaMemset(f->d_hist, 0, sizeof(float) * 2 * FE_MAXS * 16));  // memset of the filter states, hackrf.c:213-214
  // HackCD is a zero-initialised global (hackrf.c:81): DC, imbalance and sin(phi) all start at 0 — the imbalance estimate
  // then takes ~Power_alpha seconds to reach the true I/Q power ratio, and the I gain is large meanwhile (hackrf.c:191),
  // exactly as in the reference; the gains in force for the first callback block are the file-scope 1, 1, 1, 0
  // (hackrf.c:121-124). ka9q_frontend_set_estimates starts from a calibrated state instead.
  FeEst e;
  memset(&e, 0, sizeof(e));
  K9_CUDA(cudaMemcpy(f->d_est, &e, sizeof(e), cudaMemcpyHostToDevice));
  f->h_coef = FeCoef{0.f, 0.f, 1.f, 1.f, 1.f, 0.f};
  K9_CUDA(cudaFuncSetAttribute(fe_cascade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(float) * (4 * FE_BUF + 2 * FE_MAXS * 16))));
  *out = f;
  return 0;
}

int ka9q_frontend_destroy(ka9q_frontend* f) {
  if (!f) return 0;
  cudaSetDevice(f->cfg.device);
  cudaStreamSynchronize(f->st);
  void* p[] = {f->d_raw, f->d_coef, f->d_est, f->d_partial, f->d_counter, f->d_hist, f->d_tail, f->d_energy, f->d_out};
  for (void* q : p)
    if (q) cudaFree(q);
  cudaStreamDestroy(f->st);
  delete f;
  return 0;
}

// Device work of one batch: `nsamples` int8 I/Q samples already at f->d_raw (a whole number of callback blocks).
static int frontend_run_device(ka9q_frontend* f, long long nsamples) {
  const int cb = f->cfg.callback_samples;
  const int nblk = (int)(nsamples / cb);
  if (f->coef_cap < nblk + 1) {
    if (f->d_coef) cudaFree(f->d_coef);
    K9_CUDA(cudaMalloc(&f->d_coef, sizeof(FeCoef) * (nblk + 1)));
    f->coef_cap = nblk + 1;
  }
  const size_t n_out = (size_t)(nsamples >> f->S);
  if (f->out_cap < n_out) {
    if (f->d_out) cudaFree(f->d_out);
    K9_CUDA(cudaMalloc(&f->d_out, sizeof(short2) * n_out));
    f->out_cap = n_out;
  }
  K9_CUDA(cudaMemcpyAsync(f->d_coef, &f->h_coef, sizeof(FeCoef), cudaMemcpyHostToDevice, f->st));
  const float rate_factor = 1. / ((double)f->cfg.decimate * f->cfg.out_samprate * f->cfg.power_alpha);  // hackrf.c:138
  for (int b = 0; b < nblk; b++)
    fe_stats_kernel<<<FE_STATS_CTAS, FE_THREADS, 0, f->st>>>(f->d_raw + (size_t)b * cb, cb, b, f->d_coef, f->d_est, f->d_partial,
                                                              f->d_counter, f->cfg.dc_alpha, rate_factor);
  CascadeArgs a;
  memset(&a, 0, sizeof(a));
  a.raw = f->d_raw;
  a.coef = f->d_coef;
  a.cb_samples = cb;
  a.sample0 = f->samples;
  a.offset = f->cfg.offset;
  a.hist = f->d_hist;
  a.tail = f->d_tail;
  a.coeff = make_float4((float)(-6. / 802), (float)(33. / 802), (float)(-116. / 802), (float)(490. / 802));  // hackrf.c:229-237
  a.S = f->S;
  a.n_in = (int)nsamples;
  a.atten = powf(.5, f->S);
  a.out = f->d_out;
  a.out_energy = nullptr;
  const size_t smem = sizeof(float) * (4 * FE_BUF + 2 * FE_MAXS * 16);
  fe_cascade_kernel<<<(unsigned)((n_out + FE_C - 1) / FE_C), FE_THREADS, smem, f->st>>>(a);
  K9_CUDA(cudaGetLastError());
  // the tails become the next call's histories; the last coefficient set is the next call's first
  K9_CUDA(cudaMemcpyAsync(f->d_hist, f->d_tail, sizeof(float) * 2 * FE_MAXS * 16, cudaMemcpyDeviceToDevice, f->st));
  K9_CUDA(cudaMemcpyAsync(&f->h_coef, f->d_coef + nblk, sizeof(FeCoef), cudaMemcpyDeviceToHost, f->st));
  f->samples += nsamples;
  return 0;
}

static int frontend_upload(ka9q_frontend* f, const void* iq8, long long nsamples) {
  if (!f || !iq8 || nsamples <= 0 || nsamples % f->cfg.callback_samples != 0) {
    set_error("nsamples must be a positive multiple of callback_samples");
    return -1;
  }
  if (nsamples >= (1ll << 30)) {
    set_error("batch too large");
    return -1;
  }
  K9_CUDA(cudaSetDevice(f->cfg.device));
  if (f->raw_cap < (size_t)nsamples) {
    if (f->d_raw) cudaFree(f->d_raw);
    K9_CUDA(cudaMalloc(&f->d_raw, sizeof(char2) * (size_t)nsamples));
    f->raw_cap = (size_t)nsamples;
  }
  K9_CUDA(cudaStreamSynchronize(f->st));  // h_coef of the previous call has landed
  K9_CUDA(cudaMemcpyAsync(f->d_raw, iq8, sizeof(char2) * (size_t)nsamples, cudaMemcpyHostToDevice, f->st));
  return 0;
}

// iq8: HOST int8 I/Q, nsamples complex samples = a whole number of callback blocks; out: HOST int16 I/Q, nsamples / decimate.
int ka9q_frontend_process(ka9q_frontend* f, const void* iq8, long long nsamples, int16_t* out) {
  if (frontend_upload(f, iq8, nsamples)) return -1;
  if (frontend_run_device(f, nsamples)) return -1;
  if (out) K9_CUDA(cudaMemcpyAsync(out, f->d_out, sizeof(short2) * (size_t)(nsamples >> f->S), cudaMemcpyDeviceToHost, f->st));
  K9_CUDA(cudaStreamSynchronize(f->st));
  return 0;
}

// Same, but the decimated stream goes device-to-device into a channelizer's ring (no host round trip).
int ka9q_frontend_process_to_stream(ka9q_frontend* f, const void* iq8, long long nsamples, ka9q_stream* s) {
  if (!s) {
    set_error("null stream");
    return -1;
  }
  if (frontend_upload(f, iq8, nsamples)) return -1;
  if (frontend_run_device(f, nsamples)) return -1;
  K9_CUDA(cudaStreamSynchronize(f->st));
  return ka9q_stream_push_device(s, f->d_out, nsamples >> f->S);
}

// Benchmark form: re-runs the device work on the batch already resident from the last process call.
int ka9q_frontend_rerun_resident(ka9q_frontend* f, long long nsamples, float* ms) {
  if (!f || !f->d_raw || (size_t)nsamples > f->raw_cap || nsamples % f->cfg.callback_samples != 0) {
    set_error("no resident batch of that size");
    return -1;
  }
  K9_CUDA(cudaSetDevice(f->cfg.device));
  K9_CUDA(cudaStreamSynchronize(f->st));
  cudaEvent_t e0, e1;
  K9_CUDA(cudaEventCreate(&e0));
  K9_CUDA(cudaEventCreate(&e1));
  K9_CUDA(cudaEventRecord(e0, f->st));
  const int r = frontend_run_device(f, nsamples);
  K9_CUDA(cudaEventRecord(e1, f->st));
  K9_CUDA(cudaStreamSynchronize(f->st));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms) *ms = t;
  return r;
}

// Start from known estimates instead of the reference's zeros; the gains follow as hackrf.c:190-193 derives them.
int ka9q_frontend_set_estimates(ka9q_frontend* f, float dc_i, float dc_q, float imbalance, float sinphi) {
  if (!f || !(imbalance > 0) || !(fabsf(sinphi) < 1)) {
    set_error("bad argument");
    return -1;
  }
  K9_CUDA(cudaSetDevice(f->cfg.device));
  K9_CUDA(cudaStreamSynchronize(f->st));
  FeEst e;
  K9_CUDA(cudaMemcpy(&e, f->d_est, sizeof(e), cudaMemcpyDeviceToHost));
  e.dc_re = dc_i;
  e.dc_im = dc_q;
  e.imbalance = imbalance;
  e.sinphi = sinphi;
  K9_CUDA(cudaMemcpy(f->d_est, &e, sizeof(e), cudaMemcpyHostToDevice));
  f->h_coef.dc_re = dc_i;
  f->h_coef.dc_im = dc_q;
  f->h_coef.gain_q = sqrtf(0.5f * (1 + imbalance));
  f->h_coef.gain_i = sqrtf((float)(0.5 * (1 + 1. / imbalance)));
  f->h_coef.secphi = 1 / sqrtf(1 - sinphi * sinphi);
  f->h_coef.tanphi = sinphi * f->h_coef.secphi;
  return 0;
}

int ka9q_frontend_get_status(ka9q_frontend* f, ka9q_frontend_status* out) {
  if (!f || !out) {
    set_error("null argument");
    return -1;
  }
  K9_CUDA(cudaSetDevice(f->cfg.device));
  K9_CUDA(cudaStreamSynchronize(f->st));
  FeEst e;
  K9_CUDA(cudaMemcpy(&e, f->d_est, sizeof(e), cudaMemcpyDeviceToHost));
  out->dc_i = e.dc_re;
  out->dc_q = e.dc_im;
  out->imbalance = e.imbalance;
  out->sinphi = e.sinphi;
  out->in_power = e.in_power;
  out->clips = (long long)e.clips;
  out->samples = f->samples;
  return 0;
}

}  // extern "C"
