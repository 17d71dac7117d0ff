// Batch layer host side (C++ behind the C ABI of include/ka9q_b200.h, section B): one I/Q stream on one GPU,
// K channels. Owns the device I/Q ring, the spectrum buffers, per-channel parameter/state/response arrays and the
// CUDA streams; designs filters (K4), runs the forward FFT (K1+K2) once per block and the fused channel kernels (K3).
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <map>
#include <vector>
#include "../../include/ka9q_b200.h"
#include "bigfft.cuh"
#include "chan.cuh"
#include "design.cuh"
#include "fft2048.cuh"
#include "util.cuh"
#include "stream_priv.cuh"

namespace k9 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

}  // namespace k9

using namespace k9;


// ------------------------------------------------------------------ helpers

static float default_headroom() { return (float)pow(10., -15. / 20); }  // main.c:117

// dB2voltage(x) = powf(10., x/20.)  (dsp.h:37)
static float dB2voltage_f(float x) { return powf(10., (x) / 20.); }

static int window_index(ka9q_stream* s, float beta) {
  for (size_t i = 0; i < s->betas.size(); i++)
    if (s->betas[i] == beta) return (int)i;
  s->betas.push_back(beta);
  return (int)s->betas.size() - 1;
}

// normalised set_filter edges exactly as each demodulator computes them:
//   FM:        low/dsamprate, dsamprate = (float)samprate / decimate            (fm.c:27,35)
//   AM/linear: samptime*low,  samptime  = decimate / (float)samprate            (am.c:21,41; linear.c:29,81)
static void normalised_edges(const ka9q_stream* s, const ka9q_chan_params& p, float* lo, float* hi) {
  if (p.demod_type == KA9Q_FM_DEMOD) {
    float const dsamprate = (float)s->cfg.samprate / s->cfg.decimate;
    *lo = p.low / dsamprate;
    *hi = p.high / dsamprate;
  } else {
    float const samptime = (float)s->cfg.decimate / (float)s->cfg.samprate;
    *lo = samptime * p.low;
    *hi = samptime * p.high;
  }
}

static float response_gain(const ka9q_stream* s, const ka9q_chan_params& p) {
  float gain = 1. / ((float)s->N);  // filter.c:518
  if (p.demod_type == KA9Q_LINEAR_DEMOD && (p.flags & KA9Q_FLAG_ISB)) gain *= M_SQRT1_2;  // CROSS_CONJ, filter.c:520-521
  return gain;
}

// Designs the responses of channels [first, first + count). Channels whose set_filter arguments are identical (edges,
// gain, window) get ONE row of the response table: the design runs once per distinct filter and the channel kernels of
// all of them read the same 16 KB (L1 / L2 resident) instead of one private copy each — at cfg5 all 8192 responses are
// the same filter. Row c is channel c's own row; resp_slot points a follower at its group's first channel. Channels
// outside the range keep their rows (retuning one channel after commit gives it a private row again, see set_filter).
static int design_channels(ka9q_stream* s, int first, int count) {
  std::vector<DesignSpec> all(count);
  for (int i = 0; i < count; i++) {
    const ka9q_chan_params& p = s->chans[first + i];
    normalised_edges(s, p, &all[i].low, &all[i].high);
    all[i].gain = response_gain(s, p);
    all[i].window = window_index(s, p.kaiser_beta);
    all[i].fine = (float)(s->fine_bins[first + i] / NDEC);
  }
  // group identical specs: leader[i] = index (within the range) of the first channel with the same filter
  std::vector<int> leader(count), uniq;
  {
    std::map<std::array<float, 5>, int> seen;
    for (int i = 0; i < count; i++) {
      const std::array<float, 5> key = {all[i].low, all[i].high, all[i].gain, (float)all[i].window, all[i].fine};
      auto it = seen.find(key);
      if (it == seen.end()) {
        seen[key] = i;
        leader[i] = i;
        uniq.push_back(i);
      } else {
        leader[i] = it->second;
      }
    }
  }
  const int nu = (int)uniq.size();
  std::vector<DesignSpec> specs(nu);
  for (int u = 0; u < nu; u++) specs[u] = all[uniq[u]];
  const int nb = (int)s->betas.size();
  std::vector<float> win((size_t)nb * s->mdec);
  for (int i = 0; i < nb; i++) kaiser_window_host(&win[(size_t)i * s->mdec], s->mdec, s->betas[i]);
  if (s->d_windows) cudaFree(s->d_windows);
  K9_CUDA(cudaMalloc(&s->d_windows, sizeof(float) * win.size()));
  K9_CUDA(cudaMemcpy(s->d_windows, win.data(), sizeof(float) * win.size(), cudaMemcpyHostToDevice));
  DesignSpec* d_specs = nullptr;
  float2 *d_work = nullptr, *d_out = nullptr;
  float* d_ng = nullptr;
  K9_CUDA(cudaMalloc(&d_specs, sizeof(DesignSpec) * nu));
  K9_CUDA(cudaMalloc(&d_work, sizeof(float2) * 2 * (size_t)nu * NDEC));
  K9_CUDA(cudaMalloc(&d_ng, sizeof(float) * nu));
  // all distinct: design straight into the channels' own rows; otherwise into a scratch table and copy the rows out
  const bool direct = nu == count;
  if (!direct) K9_CUDA(cudaMalloc(&d_out, sizeof(float2) * (size_t)nu * NDEC));
  float2* out = direct ? s->d_resp + (size_t)first * NDEC : d_out;
  K9_CUDA(cudaMemcpy(d_specs, specs.data(), sizeof(DesignSpec) * nu, cudaMemcpyHostToDevice));
  int r = design_complex_batch(&s->p2048, NDEC, s->mdec, d_specs, nu, s->d_windows, out, d_work, s->s_comp);
  // noise_gain = N * sum |H|^2, doubled for CROSS_CONJ (filter.c:472-497): computed unscaled, scaled on the host
  if (r == 0) r = noise_gain_device(out, NDEC, NDEC, nu, 1.0f, d_ng, s->s_comp);
  std::vector<float> ng(nu);
  if (r == 0) {
    cudaError_t e = cudaMemcpyAsync(ng.data(), d_ng, sizeof(float) * nu, cudaMemcpyDeviceToHost, s->s_comp);
    for (int u = 0; u < nu && e == cudaSuccess && !direct; u++)
      e = cudaMemcpyAsync(s->d_resp + (size_t)(first + uniq[u]) * NDEC, d_out + (size_t)u * NDEC, sizeof(float2) * NDEC,
                          cudaMemcpyDeviceToDevice, s->s_comp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->s_comp);
    if (e != cudaSuccess) {
      set_error("filter design failed: %s", cudaGetErrorString(e));
      r = -1;
    }
  }
  cudaFree(d_specs);
  cudaFree(d_work);
  cudaFree(d_ng);
  if (d_out) cudaFree(d_out);
  if (r) {
    if (!*get_error()) set_error("filter design launch failed");
    return -1;
  }
  s->h_noise_gain.resize(s->chans.size());
  std::vector<int> upos(count, -1);
  for (int u = 0; u < nu; u++) upos[uniq[u]] = u;
  for (int i = 0; i < count; i++) {
    const ka9q_chan_params& p = s->chans[first + i];
    const bool cc = p.demod_type == KA9Q_LINEAR_DEMOD && (p.flags & KA9Q_FLAG_ISB);
    const float g = ng[upos[leader[i]]];
    s->h_noise_gain[first + i] = cc ? 2 * s->N * g : s->N * g;
    s->h_params[first + i].resp_slot = first + leader[i];
  }
  return 0;
}

// FM post-detection audio responses (fm.c:39-66), one per distinct beta, stored full-length Hermitian
static int design_audio(ka9q_stream* s, const std::vector<float>& audio_betas) {
  const int na = (int)audio_betas.size();
  if (na == 0) return 0;
  const int AN = NDEC, AL = s->olen, AM = s->mdec;
  float const dsamprate = (float)s->cfg.samprate / s->cfg.decimate;
  float const filter_gain = 10. / AN;
  const int nh = AN / 2 + 1;
  std::vector<float2> half((size_t)na * nh, make_float2(0.f, 0.f));
  std::vector<float> win((size_t)na * AM);
  for (int a = 0; a < na; a++) {
    for (int j = 0; j <= AN / 2; j++) {
      float const f = (float)j * dsamprate / AN;
      if (f >= 300 && f <= 6000) half[(size_t)a * nh + j].x = filter_gain * 300. / f;
    }
    kaiser_window_host(&win[(size_t)a * AM], AM, audio_betas[a]);
  }
  (void)AL;
  float2 *d_half = nullptr, *d_work = nullptr;
  float* d_win = nullptr;
  K9_CUDA(cudaMalloc(&d_half, sizeof(float2) * half.size()));
  K9_CUDA(cudaMalloc(&d_work, sizeof(float2) * 2 * (size_t)AN));
  K9_CUDA(cudaMalloc(&d_win, sizeof(float) * win.size()));
  K9_CUDA(cudaMalloc(&s->d_audio_resp, sizeof(float2) * (size_t)na * AN));
  K9_CUDA(cudaMemcpy(d_half, half.data(), sizeof(float2) * half.size(), cudaMemcpyHostToDevice));
  K9_CUDA(cudaMemcpy(d_win, win.data(), sizeof(float) * win.size(), cudaMemcpyHostToDevice));
  int r = 0;
  for (int a = 0; a < na && r == 0; a++)
    r = window_rfilter_device(&s->p2048, AM, d_half + (size_t)a * nh, s->d_audio_resp + (size_t)a * AN, 1,
                              d_win + (size_t)a * AM, d_work, s->s_comp);
  cudaError_t e = cudaStreamSynchronize(s->s_comp);
  cudaFree(d_half);
  cudaFree(d_work);
  cudaFree(d_win);
  K9_CHECK(r == 0 && e == cudaSuccess, "audio response design failed");
  return 0;
}

// ------------------------------------------------------------------ C ABI

extern "C" {

const char* ka9q_last_error(void) { return k9::get_error(); }
const char* ka9q_version(void) { return "ka9q_b200 0.1 (sm_100a)"; }

int ka9q_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ka9q_stream_create(ka9q_stream** out, const ka9q_stream_config* cfg) {
  K9_CHECK(out && cfg, "null argument");
  *out = nullptr;
  K9_CHECK(cfg->samprate > 0 && cfg->L > 0 && cfg->M > 0 && cfg->decimate > 0, "bad geometry");
  const int N = cfg->L + cfg->M - 1;
  K9_CHECK(N % cfg->decimate == 0, "FFT size %d is not divisible by decimation ratio %d (filter.c:106-107)", N,
           cfg->decimate);
  K9_CHECK(N / cfg->decimate == NDEC, "N/decimate must be %d for the fused channel kernels (got %d)", NDEC,
           N / cfg->decimate);
  K9_CHECK(cfg->L % cfg->decimate == 0, "L must be a multiple of decimate");
  K9_CHECK(cfg->iq_format == KA9Q_IQ_S16 || cfg->iq_format == KA9Q_IQ_S8, "bad iq_format");
  K9_CHECK(cfg->max_blocks >= 1, "max_blocks must be >= 1");
  K9_CHECK(ka9q_device_count() > cfg->device && cfg->device >= 0,
           "CUDA device %d not available (no CPU fallback exists)", cfg->device);
  K9_CUDA(cudaSetDevice(cfg->device));
  ka9q_stream* s = new ka9q_stream();
  s->cfg = *cfg;
  s->N = N;
  s->olen = cfg->L / cfg->decimate;
  s->mdec = (cfg->M - 1) / cfg->decimate + 1;  // filter.c:514
  s->bytes_per_samp = cfg->iq_format == KA9Q_IQ_S16 ? 4 : 2;
  if (bigfft_plan_create(&s->fwd, N) != 0) {
    set_error("forward FFT size %d has no supported factorisation (2^a 3^b 5^c with pass sizes 16..400)", N);
    delete s;
    return -1;
  }
  if (bigfft_plan_create(&s->p2048, NDEC) != 0) {
    set_error("internal: 2048-point plan");
    delete s;
    return -1;
  }
  if (const char* e = getenv("KA9Q_B200_FFT_BLOCKS_PER_LAUNCH")) s->fft_blocks_per_launch = atoi(e);
  *out = s;
  return 0;
}

int ka9q_stream_add_channel(ka9q_stream* s, const ka9q_chan_params* p) {
  K9_CHECK(s && p, "null argument");
  K9_CHECK(!s->committed, "channels must be added before commit");
  K9_CHECK(p->demod_type >= 0 && p->demod_type <= 2, "bad demod_type");
  K9_CHECK(!(p->flags & (KA9Q_FLAG_PLL | KA9Q_FLAG_SQUARE)) || p->demod_type == KA9Q_LINEAR_DEMOD,
           "pll / square are options of the linear demodulator (modes.c:104-120)");
  K9_CHECK(!(p->hangtime < 0), "negative AGC hang time");
  K9_CHECK(!isnan(p->low) && !isnan(p->high), "filter edges must be set (set_filter returns -1 on NAN, filter.c:504)");
  ka9q_chan_params q = *p;
  if (q.flags & KA9Q_FLAG_SQUARE) q.flags |= KA9Q_FLAG_PLL;  // square implies pll (modes.c:113-116)
  if (q.low > q.high) std::swap(q.low, q.high);  // radio.c:347-353
  if (isnan(q.headroom)) q.headroom = default_headroom();
  if (q.channels != 1 && q.channels != 2) q.channels = (q.demod_type == KA9Q_LINEAR_DEMOD) ? 2 : 1;
  if (q.demod_type != KA9Q_LINEAR_DEMOD) q.channels = 1;  // fm.c:30, am.c:36
  q.bin %= s->N;
  if (q.bin < 0) q.bin += s->N;
  s->chans.push_back(q);
  s->fine_bins.push_back(0.0);
  return (int)s->chans.size() - 1;
}

// Off-grid carriers. The second LO of the reference is any double (radio.c:217,299); the shared forward
// FFT gives LOs on the N-point grid only. With f_LO = -(bin + e)/N: mixing before a filter h equals filtering with
// h[m] * exp(+j 2 pi e m / N) and mixing afterwards, so the grid part stays a bin rotation, the
// channel's impulse response is designed with that phase ramp (design.cu) and the kept samples are rotated by
// exp(-j 2 pi e n / N), n = m L + D i, at the output rate: in the FM discriminator's input, folded into the shift
// oscillator of the linear demodulator; the AM envelope does not see it. Not for ISB (the sideband fold does not commute
// with the rotation) or the coherent modes (the loop tracks the residue itself; use the grid).
int ka9q_stream_set_fine_lo(ka9q_stream* s, int chan, double bins) {
  K9_CHECK(s, "null argument");
  K9_CHECK(!s->committed, "ka9q_stream_set_fine_lo must be called before commit");
  K9_CHECK(chan >= 0 && chan < (int)s->chans.size(), "bad channel");
  K9_CHECK(bins >= -0.5 && bins <= 0.5, "the fine part of a carrier is at most half a bin");
  const ka9q_chan_params& p = s->chans[chan];
  K9_CHECK(bins == 0 || !(p.demod_type == KA9Q_LINEAR_DEMOD && (p.flags & (KA9Q_FLAG_ISB | KA9Q_FLAG_PLL))),
           "ISB and coherent (pll / square) channels take carriers on the bin grid only");
  s->fine_bins[chan] = bins;
  return 0;
}

// carrier frequency (Hz from the first LO, either sign) -> nearest grid bin and the fraction left over
int ka9q_stream_split_carrier(const ka9q_stream* s, double carrier_hz, long long* bin, double* fine_bins) {
  K9_CHECK(s && bin && fine_bins, "null argument");
  const double b = carrier_hz * (double)s->N / (double)s->cfg.samprate;
  const double r = nearbyint(b);
  *bin = (long long)r;
  *fine_bins = b - r;
  return 0;
}

static int commit_impl(ka9q_stream* s);
static void release_resources(ka9q_stream* s);

int ka9q_stream_commit(ka9q_stream* s) {
  K9_CHECK(s, "null argument");
  K9_CHECK(!s->committed, "already committed");
  K9_CHECK(!s->chans.empty(), "no channels");
  if (commit_impl(s)) {
    release_resources(s);  // nothing allocated so far outlives a failed commit (the error text is kept)
    return -1;
  }
  return 0;
}

static int commit_impl(ka9q_stream* s) {
  K9_CUDA(cudaSetDevice(s->cfg.device));
  const int K = (int)s->chans.size();
  const int B = s->cfg.max_blocks;
  const int L = s->cfg.L, M = s->cfg.M, N = s->N;
  K9_CUDA(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
  K9_CUDA(cudaStreamCreateWithFlags(&s->s_comp, cudaStreamNonBlocking));
  K9_CUDA(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
  K9_CUDA(cudaStreamCreateWithFlags(&s->s_fm, cudaStreamNonBlocking));
  K9_CUDA(cudaStreamCreateWithFlags(&s->s_am, cudaStreamNonBlocking));
  K9_CUDA(cudaStreamCreateWithFlags(&s->s_lin, cudaStreamNonBlocking));
  {
    // The forward FFT (and the multi-GPU exchange behind it) of batch k+1 becomes runnable at about the same moment as the
    // channel kernels of batch k. With equal priorities the channel kernels' long-lived CTAs (they loop over the blocks
    // of the batch) take every slot first and the transform waits for them to drain; at a higher priority the transform
    // goes first and the exchange then runs under the channel kernels. KA9Q_B200_FFT_PRIO=0 turns it off (A/B).
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const char* ev = getenv("KA9Q_B200_FFT_PRIO");
    const bool prio = ev ? atoi(ev) != 0 : false;  // (multi-GPU streams switch it on in ka9q_stream_mgpu_setup)
    K9_CUDA(cudaStreamCreateWithPriority(&s->s_fft, cudaStreamNonBlocking, prio ? hi : lo));
  }
  cudaEvent_t* evs[] = {&s->e_pushed, &s->e_fork, &s->e_am, &s->e_lin, &s->e_fm, &s->e_comp_done[0], &s->e_comp_done[1],
                        &s->e_fetched[0], &s->e_fetched[1], &s->e_spec_ready[0], &s->e_spec_ready[1], &s->e_spec_ready[2],
                        &s->e_spec_free[0], &s->e_spec_free[1], &s->e_spec_free[2]};
  for (auto e : evs) K9_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  K9_CUDA(cudaEventCreate(&s->e_fft0));
  K9_CUDA(cudaEventCreate(&s->e_fft1));
  K9_CUDA(cudaEventCreate(&s->e_chan1));

  // I/Q ring: history + two batches, so the copy of batch k+1 may overlap the compute of batch k
  s->ring_cap = (long long)(M - 1) + 2LL * B * L;
  K9_CUDA(cudaMalloc(&s->d_ring, (size_t)s->ring_cap * s->bytes_per_samp));
  K9_CUDA(cudaMemset(s->d_ring, 0, (size_t)s->ring_cap * s->bytes_per_samp));  // zero history (filter.c:77)
  {
    // Spectrum buffers. Two: the FFT of batch k+1 under the channel kernels of batch k. Three (default): the FFT and, on
    // multi-GPU runs, the exchange behind it may start a whole batch earlier, so that they are finished — however thinly
    // they are scheduled beside the channel kernels' long-lived CTAs — before the channel kernels of their batch could
    // start (measured at 2 GPUs: the exchange otherwise sits exposed between consecutive channel launches).
    const char* ev = getenv("KA9Q_B200_SPEC_BUFFERS");
    s->nspec = ev && atoi(ev) == 2 ? 2 : 3;
  }
  K9_CUDA(cudaMalloc(&s->d_spec, sizeof(float2) * s->nspec * (size_t)B * N));  // see issue_fft
  K9_CUDA(cudaMalloc(&s->d_tmp0, sizeof(float2) * (size_t)B * N));
  if (s->fwd.npass >= 3) K9_CUDA(cudaMalloc(&s->d_tmp1, sizeof(float2) * (size_t)B * N));
  K9_CUDA(cudaMalloc(&s->d_energy, sizeof(float) * B));
  {
    std::vector<float2> tw(FFT2048_TW_FLOAT2);
    fft2048_fill_twiddles(tw.data());
    K9_CUDA(cudaMalloc(&s->d_tw2048, sizeof(float2) * tw.size()));
    K9_CUDA(cudaMemcpy(s->d_tw2048, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
  }
  // per-channel arrays
  K9_CUDA(cudaMalloc(&s->d_resp, sizeof(float2) * (size_t)K * NDEC));
  K9_CUDA(cudaMalloc(&s->d_params, sizeof(ChanParams) * K));
  K9_CUDA(cudaMalloc(&s->d_state, sizeof(ChanState) * K));
  // PCM and status are double-buffered so the D2H copy of batch k overlaps the compute of batch k+1
  K9_CUDA(cudaMalloc(&s->d_status, sizeof(ChanStatus) * 2 * (size_t)B * K));
  K9_CUDA(cudaMemset(s->d_status, 0, sizeof(ChanStatus) * 2 * (size_t)B * K));

  if (s->n0_enabled) {
    K9_CUDA(cudaStreamCreateWithFlags(&s->s_n0, cudaStreamNonBlocking));
    K9_CUDA(cudaEventCreateWithFlags(&s->e_n0, cudaEventDisableTiming));
    K9_CUDA(cudaMalloc(&s->d_n0_chan, sizeof(N0Chan) * K));
    K9_CUDA(cudaMalloc(&s->d_n0_P, sizeof(float) * (size_t)B * N));
    K9_CUDA(cudaMalloc(&s->d_n0_partial, sizeof(double) * (size_t)B * N0_POWER_CTAS));
    K9_CUDA(cudaMalloc(&s->d_n0_blk, sizeof(N0Block) * B));
    K9_CUDA(cudaMalloc(&s->d_n0_T, sizeof(float) * (size_t)B * K));
    K9_CUDA(cudaMalloc(&s->d_n0_list, sizeof(float) * (size_t)B * N0_LIST_CAP));
    K9_CUDA(cudaMalloc(&s->d_n0_raw, sizeof(float) * 2 * (size_t)B * K));     // double-buffered like the PCM / status
    K9_CUDA(cudaMalloc(&s->d_n0_smooth, sizeof(float) * 2 * (size_t)B * K));
    K9_CUDA(cudaMalloc(&s->d_n0_state, sizeof(float) * K));
    K9_CUDA(cudaMemset(s->d_n0_state, 0, sizeof(float) * K));  // struct demod is zero-initialised: sig.n0 starts at 0
    std::vector<N0Chan> nc(K);
    for (int c = 0; c < K; c++) {
      const ka9q_chan_params& p = s->chans[c];
      nc[c].bin = (int)p.bin;
      n0_passband_bins(N, s->cfg.samprate, p.low, p.high, &nc[c].nlo, &nc[c].nhi);
      nc[c].alpha = p.demod_type == KA9Q_FM_DEMOD ? 0.01f : 0.001f;
    }
    K9_CUDA(cudaMemcpy(s->d_n0_chan, nc.data(), sizeof(N0Chan) * K, cudaMemcpyHostToDevice));
  }

  // parameter blocks, PCM layout, work lists
  s->h_params.resize(K);
  std::vector<ChanState> st(K);
  memset(st.data(), 0, sizeof(ChanState) * K);
  std::vector<float> audio_betas;
  std::map<int, std::vector<int>> fm_groups;  // audio slot -> channels
  std::vector<int2> w_fm, w_am, w_lin, w_pll;
  std::vector<PllParams> pll_params;
  long long off = 0;
  bool any_fm = false;
  float const dsamprate = (float)s->cfg.samprate / s->cfg.decimate;
  for (int c = 0; c < K; c++) {
    const ka9q_chan_params& p = s->chans[c];
    ChanParams& P = s->h_params[c];
    memset(&P, 0, sizeof(P));
    P.bin = p.bin;
    P.demod = p.demod_type;
    P.flags = p.flags;
    P.channels = p.channels;
    P.pcm_off = (int)off;
    off += (long long)s->olen * p.channels;
    P.headroom = p.headroom;
    P.audio_slot = -1;
    P.phase_step = (int)((p.bin % N) * (long long)(L % N) % N);
    float const samptime = (float)s->cfg.decimate / (float)s->cfg.samprate;
    if (p.demod_type == KA9Q_FM_DEMOD) {
      any_fm = true;
      // fm.c:86: (headroom * M_1_PI * dsamprate) / fabsf(low - high), evaluated in double, stored to float
      P.fm_gain = (p.headroom * M_1_PI * dsamprate) / fabsf(p.low - p.high);
      st[c].fm_state = make_float2(1.f, 0.f);  // fm.c:26
      P.shift_cycles = -s->fine_bins[c] / NDEC;  // FM has no shift oscillator: the field carries the fine LO alone
      if (!(p.flags & KA9Q_FLAG_FLAT)) {
        int slot = -1;
        for (size_t i = 0; i < audio_betas.size(); i++)
          if (audio_betas[i] == p.kaiser_beta) slot = (int)i;
        if (slot < 0) {
          audio_betas.push_back(p.kaiser_beta);
          slot = (int)audio_betas.size() - 1;
        }
        P.audio_slot = slot;
        fm_groups[slot].push_back(c);
      } else {
        w_fm.push_back(make_int2(c, -1));
      }
    } else {
      P.recovery_factor = dB2voltage_f(p.recovery_rate * samptime);  // am.c:27, linear.c:33
      P.hangmax = (int)(p.hangtime / samptime);                       // am.c:29, linear.c:37
      if (p.demod_type == KA9Q_AM_DEMOD) {
        st[c].agc_gain = dB2voltage_f(80.);  // am.c:30
        w_am.push_back(make_int2(c, -1));
      } else {
        st[c].agc_gain = dB2voltage_f(100.0);  // linear.c:39
        // radio.c:313: shift * decimate / samprate, cycles per output sample
        P.shift_cycles = (p.shift == 0) ? 0.0 : (double)p.shift * s->cfg.decimate / (double)s->cfg.samprate;
        P.shift_cycles -= s->fine_bins[c] / NDEC;  // the fine part of the second LO rides the same oscillator
        if (p.flags & KA9Q_FLAG_PLL) {
          // coherent modes: constants of linear.c:29-65 in the reference's own types and order of evaluation
          const bool square = p.flags & KA9Q_FLAG_SQUARE;
          PllParams Q;
          memset(&Q, 0, sizeof(Q));
          Q.samptime = samptime;
          Q.blocktime = samptime * L;
          float const snrthreshdb = 3;
          int const fftsize = 1 << 16;
          float const damping = M_SQRT1_2;
          float const lock_time = 1;
          Q.snrthresh = powf(10, snrthreshdb / 10);
          Q.lock_limit = (int)round(lock_time / samptime);
          Q.binsize = 1. / (fftsize * samptime);
          float const searchhigh = 300, searchlow = -300;
          Q.lowlimit = (int)round((square ? 2 : 1) * searchlow / Q.binsize);
          Q.highlimit = (int)round((square ? 2 : 1) * searchhigh / Q.binsize);
          float const loop_bw = 1;  // linear.c:26
          float const vcogain = 2 * M_PI, pdgain = 1;
          float const natfreq = loop_bw * 2 * M_PI;
          float const tau1 = vcogain * pdgain / (natfreq * natfreq);
          Q.integrator_gain = 1 / tau1;
          float const tau2 = 2 * damping / natfreq;
          Q.prop_gain = tau2 / tau1;
          w_pll.push_back(make_int2(c, (int)pll_params.size()));
          pll_params.push_back(Q);
        } else {
          w_lin.push_back(make_int2(c, -1));
        }
      }
    }
  }
  for (auto& g : fm_groups) {
    const std::vector<int>& v = g.second;
    for (size_t i = 0; i < v.size(); i += 2) w_fm.push_back(make_int2(v[i], i + 1 < v.size() ? v[i + 1] : -1));
  }
  s->pcm_stride = off;
  s->n_fm = (int)w_fm.size();
  s->n_am = (int)w_am.size();
  s->n_lin = (int)w_lin.size();
  s->n_pll = (int)w_pll.size();
  auto upload_work = [&](const std::vector<int2>& w, int2** d) -> int {
    if (w.empty()) return 0;
    K9_CUDA(cudaMalloc(d, sizeof(int2) * w.size()));
    K9_CUDA(cudaMemcpy(*d, w.data(), sizeof(int2) * w.size(), cudaMemcpyHostToDevice));
    return 0;
  };
  if (upload_work(w_fm, &s->d_work_fm) || upload_work(w_am, &s->d_work_am) || upload_work(w_lin, &s->d_work_lin) ||
      upload_work(w_pll, &s->d_work_pll))
    return -1;
  if (s->n_fm) {  // block-split form of the FM kernel: one sequence number per pair
    K9_CUDA(cudaMalloc(&s->d_fm_seq, sizeof(long long) * s->n_fm));
    K9_CUDA(cudaMemset(s->d_fm_seq, 0, sizeof(long long) * s->n_fm));
  }
  if (s->n_pll) {
    K9_CUDA(cudaStreamCreateWithFlags(&s->s_pll, cudaStreamNonBlocking));
    K9_CUDA(cudaEventCreateWithFlags(&s->e_pll, cudaEventDisableTiming));
    K9_CUDA(cudaMalloc(&s->d_pll_params, sizeof(PllParams) * s->n_pll));
    K9_CUDA(cudaMemcpy(s->d_pll_params, pll_params.data(), sizeof(PllParams) * s->n_pll, cudaMemcpyHostToDevice));
    std::vector<PllState> ps(s->n_pll);
    memset(ps.data(), 0, sizeof(PllState) * s->n_pll);  // sig.snr = 0 (linear.c:75), integrator, delta_f, lock_count = 0
    for (auto& x : ps) x.coarse_ph = x.fine_ph = make_double2(1.0, 0.0);  // linear.c:99,104
    K9_CUDA(cudaMalloc(&s->d_pll_state, sizeof(PllState) * s->n_pll));
    K9_CUDA(cudaMemcpy(s->d_pll_state, ps.data(), sizeof(PllState) * s->n_pll, cudaMemcpyHostToDevice));
    K9_CUDA(cudaMalloc(&s->d_pll_ring, sizeof(float2) * (size_t)s->n_pll * 65536));
    K9_CUDA(cudaMemset(s->d_pll_ring, 0, sizeof(float2) * (size_t)s->n_pll * 65536));
  }
  K9_CUDA(cudaMemcpy(s->d_state, st.data(), sizeof(ChanState) * K, cudaMemcpyHostToDevice));
  {
    const size_t rows_am = (size_t)B * s->n_am, rows_lin = (size_t)B * s->n_lin;
    if (rows_am) K9_CUDA(cudaMalloc(&s->d_agc_x_am, sizeof(float) * rows_am * s->olen));
    if (rows_lin) {
      K9_CUDA(cudaMalloc(&s->d_agc_x_lin, sizeof(float) * rows_lin * s->olen));
      K9_CUDA(cudaMalloc(&s->d_agc_y_lin, sizeof(float2) * rows_lin * s->olen));
    }
    if (rows_am + rows_lin) K9_CUDA(cudaMalloc(&s->d_agc_pow, sizeof(float) * 2 * (rows_am + rows_lin)));
  }
  if (s->pl_enabled && !w_fm.empty()) {
    // PL-tone analyser: one work item per de-emphasised FM channel (FLAT channels have no audio transform to tap)
    std::vector<PlWork> pw;
    for (size_t i = 0; i < w_fm.size(); i++) {
      const int ca = w_fm[i].x, cb = w_fm[i].y;
      if (s->h_params[ca].audio_slot < 0) continue;
      pw.push_back(PlWork{ca, (int)i, 0, 0});
      if (cb >= 0) pw.push_back(PlWork{cb, (int)i, 1, 0});
    }
    s->n_pl = (int)pw.size();
    if (s->n_pl) {
      K9_CUDA(cudaMalloc(&s->d_pl_work, sizeof(PlWork) * s->n_pl));
      K9_CUDA(cudaMemcpy(s->d_pl_work, pw.data(), sizeof(PlWork) * s->n_pl, cudaMemcpyHostToDevice));
      K9_CUDA(cudaMalloc(&s->d_pl_state, sizeof(PlState) * s->n_pl));
      K9_CUDA(cudaMemset(s->d_pl_state, 0, sizeof(PlState) * s->n_pl));  // struct demod is zero-initialised: plfreq = 0
      K9_CUDA(cudaMalloc(&s->d_pl_ring, sizeof(float) * (size_t)s->n_pl * 16384));
      K9_CUDA(cudaMemset(s->d_pl_ring, 0, sizeof(float) * (size_t)s->n_pl * 16384));
      K9_CUDA(cudaMalloc(&s->d_pl_spec, sizeof(float2) * (size_t)B * w_fm.size() * 65));
      K9_CUDA(cudaMemset(s->d_pl_spec, 0, sizeof(float2) * (size_t)B * w_fm.size() * 65));
      // slave response (fm.c:200-218): PL_N = AN/32 = 64, PL_L = AL/32, PL_M = PL_N - PL_L + 1; unity below 300 Hz on the
      // positive side, Kaiser beta 2, through the library's own window_rfilter
      const int PL_N = NDEC / 32, PL_L = s->olen / 32, PL_M = PL_N - PL_L + 1;
      std::vector<float2> pr(PL_N / 2 + 1, make_float2(0.f, 0.f));
      for (int j = 0; j <= PL_N / 2; j++) {
        float const f = (float)j * dsamprate / NDEC;
        if (f > 0 && f < 300) pr[j].x = 1;
      }
      K9_CHECK(window_rfilter(PL_L, PL_M, pr.data(), 2.0f) == 0, "PL response design failed: %s", get_error());
      K9_CUDA(cudaMalloc(&s->d_pl_resp, sizeof(float2) * pr.size()));
      K9_CUDA(cudaMemcpy(s->d_pl_resp, pr.data(), sizeof(float2) * pr.size(), cudaMemcpyHostToDevice));
    }
  }
  if (any_fm) {
    // one ring per channel + a spare all-zero ring that stands in for the missing partner of an unpaired FM channel
    K9_CUDA(cudaMalloc(&s->d_audio_hist, sizeof(float) * (size_t)(K + 1) * NDEC));
    K9_CUDA(cudaMemset(s->d_audio_hist, 0, sizeof(float) * (size_t)(K + 1) * NDEC));  // zero history (filter.c:87)
  }
  K9_CUDA(cudaMalloc(&s->d_pcm, sizeof(int16_t) * 2 * (size_t)B * s->pcm_stride));
  K9_CUDA(cudaMemset(s->d_pcm, 0, sizeof(int16_t) * 2 * (size_t)B * s->pcm_stride));
  if (s->cfg.capture_filter_output) K9_CUDA(cudaMalloc(&s->d_filt, sizeof(float2) * (size_t)B * K * s->olen));
  // pinned staging
  K9_CUDA(cudaHostAlloc(&s->h_iq, (size_t)B * L * s->bytes_per_samp, cudaHostAllocDefault));
  K9_CUDA(cudaHostAlloc((void**)&s->h_pcm, sizeof(int16_t) * (size_t)B * s->pcm_stride, cudaHostAllocDefault));
  K9_CUDA(cudaHostAlloc((void**)&s->h_status, sizeof(ChanStatus) * (size_t)B * K, cudaHostAllocDefault));

  // the forward FFT (and the multi-GPU exchange kernels) run beside the channel kernels: same shared-memory carve-out
  s->carveout = fm_carveout(s->n_am + s->n_lin + s->n_pll > 0);
  bigfft_set_carveout(s->carveout);
  if (design_channels(s, 0, K)) return -1;  // also assigns the response rows (resp_slot)
  K9_CUDA(cudaMemcpy(s->d_params, s->h_params.data(), sizeof(ChanParams) * K, cudaMemcpyHostToDevice));
  if (design_audio(s, audio_betas)) return -1;
  s->committed = true;
  return 0;
}

int ka9q_stream_set_filter(ka9q_stream* s, int chan, float low, float high, float kaiser_beta) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CHECK(chan >= 0 && chan < (int)s->chans.size(), "bad channel");
  K9_CHECK(!isnan(low) && !isnan(high), "NAN edge (filter.c:504-505)");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  ka9q_chan_params& p = s->chans[chan];
  if (p.demod_type == KA9Q_FM_DEMOD && !(p.flags & KA9Q_FLAG_FLAT))
    K9_CHECK(kaiser_beta == p.kaiser_beta, "changing kaiser_beta of a de-emphasised FM channel after commit is unsupported");
  p.low = std::min(low, high);
  p.high = std::max(low, high);
  p.kaiser_beta = kaiser_beta;
  if (ka9q_stream_sync(s)) return -1;
  // channels that share this channel's response row move to a row of their own first (the new group leader's)
  {
    int heir = -1;
    for (int c = 0; c < (int)s->chans.size(); c++) {
      if (c == chan || s->h_params[c].resp_slot != chan) continue;
      if (heir < 0) {
        heir = c;
        K9_CUDA(cudaMemcpy(s->d_resp + (size_t)heir * NDEC, s->d_resp + (size_t)chan * NDEC, sizeof(float2) * NDEC,
                           cudaMemcpyDeviceToDevice));
      }
      s->h_params[c].resp_slot = heir;
      K9_CUDA(cudaMemcpy(s->d_params + c, &s->h_params[c], sizeof(ChanParams), cudaMemcpyHostToDevice));
    }
  }
  if (design_channels(s, chan, 1)) return -1;  // into the channel's own row
  if (p.demod_type == KA9Q_FM_DEMOD) {
    float const dsamprate = (float)s->cfg.samprate / s->cfg.decimate;
    s->h_params[chan].fm_gain = (p.headroom * M_1_PI * dsamprate) / fabsf(p.low - p.high);
  }
  K9_CUDA(cudaMemcpy(s->d_params + chan, &s->h_params[chan], sizeof(ChanParams), cudaMemcpyHostToDevice));
  return 0;
}

// PL-tone analyser (pltask, fm.c:189-285) for every de-emphasised FM channel; call before commit. The tone frequency
// (demod->sig.plfreq: 0 until the first analysis, NAN when no tone stands out) is status row field reserved[1].
int ka9q_stream_enable_pl(ka9q_stream* s, int enable) {
  K9_CHECK(s, "null argument");
  K9_CHECK(!s->committed, "ka9q_stream_enable_pl must be called before commit");
  s->pl_enabled = enable != 0;
  return 0;
}

// Noise-density estimate (compute_n0, radio.c:383-425) for every channel and block; call before commit.
int ka9q_stream_enable_n0(ka9q_stream* s, int enable) {
  K9_CHECK(s, "null argument");
  K9_CHECK(!s->committed, "ka9q_stream_enable_n0 must be called before commit");
  s->n0_enabled = enable != 0;
  return 0;
}

// n0 rows of the last computed batch (like ka9q_stream_fetch: async on the output stream, ka9q_stream_sync / wait_fetch
// afterwards). raw: compute_n0() of each block; smooth: demod->sig.n0 after each block. [nblocks][nchan] floats each.
int ka9q_stream_fetch_n0(ka9q_stream* s, int nblocks, float* raw, float* smooth) {
  K9_CHECK(s && s->committed && s->n0_enabled, "n0 not enabled");
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks, "nblocks out of range");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamWaitEvent(s->s_out, s->e_comp_done[s->comp_parity], 0));
  const size_t off = (size_t)s->comp_parity * s->cfg.max_blocks * s->chans.size(), n = sizeof(float) * (size_t)nblocks * s->chans.size();
  if (raw) K9_CUDA(cudaMemcpyAsync(raw, s->d_n0_raw + off, n, cudaMemcpyDeviceToHost, s->s_out));
  if (smooth) K9_CUDA(cudaMemcpyAsync(smooth, s->d_n0_smooth + off, n, cudaMemcpyDeviceToHost, s->s_out));
  K9_CUDA(cudaEventRecord(s->e_fetched[s->comp_parity], s->s_out));
  return 0;
}

int ka9q_stream_num_channels(const ka9q_stream* s) { return s ? (int)s->chans.size() : -1; }
long long ka9q_stream_pcm_stride(const ka9q_stream* s) { return s ? s->pcm_stride : -1; }
int ka9q_stream_pcm_offset(const ka9q_stream* s, int chan) {
  if (!s || !s->committed || chan < 0 || chan >= (int)s->chans.size()) return -1;
  return s->h_params[chan].pcm_off;
}
int ka9q_stream_olen(const ka9q_stream* s) { return s ? s->olen : -1; }
long long ka9q_stream_blocks_done(const ka9q_stream* s) { return s ? s->block0 : -1; }
int ka9q_stream_fft_size(const ka9q_stream* s) { return s ? s->N : -1; }
int ka9q_stream_launches_per_call(const ka9q_stream* s) {
  if (!s) return -1;
  return s->fwd.npass + (s->n_fm ? 1 : 0) + (s->n_am ? 1 : 0) + (s->n_lin ? 1 : 0) + (s->n_pll ? 1 : 0);
}

int ka9q_stream_push(ka9q_stream* s, const void* iq, int nblocks) {
  K9_CHECK(s && s->committed && iq, "bad argument");
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks, "nblocks out of range");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  const long long n = (long long)nblocks * s->cfg.L;
  // the region about to be overwritten was last read by the compute issued two calls ago
  K9_CHECK(s->pushed + n - s->block0 * (long long)s->cfg.L <= 2LL * s->cfg.max_blocks * s->cfg.L,
           "push would overwrite samples that have not been computed yet");
  // The ring holds two batches: this push overwrites the batch before the one most recently handed to compute, so it
  // must wait for the latest ISSUED compute's predecessor at least; waiting for the latest issued compute itself is
  // correct for every call order (push k+1 issued before compute k then overlaps compute k).
  K9_CUDA(cudaStreamWaitEvent(s->s_in, s->e_comp_done[s->comp_parity], 0));
  long long pos = (s->pushed + (s->cfg.M - 1)) % s->ring_cap;
  long long done = 0;
  while (done < n) {
    long long chunk = std::min(n - done, s->ring_cap - pos);
    K9_CUDA(cudaMemcpyAsync((char*)s->d_ring + pos * s->bytes_per_samp, (const char*)iq + done * s->bytes_per_samp,
                            (size_t)chunk * s->bytes_per_samp, cudaMemcpyHostToDevice, s->s_in));
    done += chunk;
    pos = (pos + chunk) % s->ring_cap;
  }
  s->pushed += n;
  K9_CUDA(cudaEventRecord(s->e_pushed, s->s_in));
  return 0;
}

// Device-resident producer (the front-end service): nsamples complex samples in device memory, any count.
int ka9q_stream_push_device(ka9q_stream* s, const void* d_iq, long long nsamples) {
  K9_CHECK(s && s->committed && d_iq && nsamples > 0, "bad argument");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CHECK(s->pushed + nsamples - s->block0 * (long long)s->cfg.L <= 2LL * s->cfg.max_blocks * s->cfg.L,
           "push would overwrite samples that have not been computed yet");
  K9_CUDA(cudaStreamWaitEvent(s->s_in, s->e_comp_done[s->comp_parity], 0));
  long long pos = (s->pushed + (s->cfg.M - 1)) % s->ring_cap, done = 0;
  while (done < nsamples) {
    const long long chunk = std::min(nsamples - done, s->ring_cap - pos);
    K9_CUDA(cudaMemcpyAsync((char*)s->d_ring + pos * s->bytes_per_samp, (const char*)d_iq + done * s->bytes_per_samp,
                            (size_t)chunk * s->bytes_per_samp, cudaMemcpyDeviceToDevice, s->s_in));
    done += chunk;
    pos = (pos + chunk) % s->ring_cap;
  }
  s->pushed += nsamples;
  K9_CUDA(cudaEventRecord(s->e_pushed, s->s_in));
  return 0;
}

static void fill_launch(ka9q_stream* s, ChanLaunch& a, int nblocks) {
  memset(&a, 0, sizeof(a));
  a.spec = s->d_spec;
  a.spec_stride = s->N;
  a.N = s->N;
  a.L = s->cfg.L;
  a.M = s->cfg.M;
  a.olen = s->olen;
  a.dsamprate = (float)s->cfg.samprate / s->cfg.decimate;
  a.nblocks = nblocks;
  a.block0 = s->phase_block;
  {
    long long st0 = (s->phase_block % s->N) * (long long)(s->cfg.L % s->N) % s->N - (long long)((s->cfg.M - 1) % s->N);
    st0 %= s->N;
    if (st0 < 0) st0 += s->N;
    a.start0 = (int)st0;
  }
  a.twN_lo = s->fwd.tw_lo;
  a.twN_hi = s->fwd.tw_hi;
  a.tw2048 = s->d_tw2048;
  a.params = s->d_params;
  a.state = s->d_state;
  a.resp = s->d_resp;
  a.audio_resp = s->d_audio_resp;
  a.audio_hist = s->d_audio_hist;
  const int wbuf = s->comp_parity ^ 1;  // the buffer this batch writes; fetch reads it once the batch is recorded
  a.pcm = s->d_pcm + (size_t)wbuf * s->cfg.max_blocks * s->pcm_stride;
  a.pcm_stride = s->pcm_stride;
  a.status = s->d_status + (size_t)wbuf * s->cfg.max_blocks * s->chans.size();
  a.nchan_total = (int)s->chans.size();
  a.filt_dbg = s->d_filt;
}

// The spectrum is double-buffered and the forward FFT (+ the multi-GPU collective) runs on its own stream, so the FFT
// of batch k+1 overlaps the channel kernels of batch k:
//   s_fft : [wait ring pushed, buffer p free] FFT -> (collective) -> record ready[p]
//   s_comp: [wait ready[p]] channel kernels -> record free[p]
float2* spec_buf(ka9q_stream* s, int p) { return s->d_spec + (size_t)p * s->cfg.max_blocks * s->N; }

int issue_fft(ka9q_stream* s, long long first_block, int blk_first, int blk_count) {
  const int p = s->spec_wr;
  K9_CUDA(cudaStreamWaitEvent(s->s_fft, s->e_pushed, 0));
  K9_CUDA(cudaStreamWaitEvent(s->s_fft, s->e_spec_free[p], 0));
  if (!s->overlap) K9_CUDA(cudaStreamWaitEvent(s->s_fft, s->e_comp_done[s->comp_parity], 0));
  if (blk_count <= 0) return 0;
  BigFftIn in;
  in.in_mode = s->cfg.iq_format == KA9Q_IQ_S16 ? IN_RING_S16 : IN_RING_S8;
  in.in = s->d_ring;
  in.ring_cap = s->ring_cap;
  in.ring_off = ((first_block + blk_first) * (long long)s->cfg.L) % s->ring_cap;
  in.ring_step = s->cfg.L;
  in.scale = s->cfg.iq_format == KA9Q_IQ_S16 ? (float)(1. / 32767) : (float)(1. / 127);  // radio.c:38-39
  in.gain = s->cfg.gain_factor;
  in.stat_from = s->cfg.M - 1;
  in.energy = s->d_energy + blk_first;
  if (s->mg_fused && s->mg_route_now) {  // channel-sharded multi-GPU: the last pass stores every row where it is read (mgpu.cu)
    in.route_mask = s->d_mg_mask;
    for (int r = 0; r < K9_MAX_RANKS; r++) in.route_delta[r] = s->mg_delta[r];
  }
  K9_CUDA(cudaMemsetAsync(s->d_energy + blk_first, 0, sizeof(float) * blk_count, s->s_fft));
  K9_CUDA(cudaEventRecord(s->e_fft0, s->s_fft));
  {
    TimedRegion tr(s, TC_FFT, s->s_fft);
    // all blocks of the batch per launch by default; KA9Q_B200_FFT_BLOCKS_PER_LAUNCH=1 keeps one block's ping-pong
    // buffers L2-resident across its passes but was measured slower (0.265 vs 0.221 ms per 4 blocks: tail effects)
    const int per_launch = s->fft_blocks_per_launch > 0 ? s->fft_blocks_per_launch : blk_count;
    for (int b = 0; b < blk_count; b += per_launch) {
      const int nb = std::min(per_launch, blk_count - b);
      BigFftIn inb = in;
      inb.ring_off = (in.ring_off + (long long)b * s->cfg.L) % s->ring_cap;
      inb.energy = in.energy + b;
      if (bigfft_exec(&s->fwd, inb, spec_buf(s, p) + (size_t)(blk_first + b) * s->N, s->N, s->d_tmp0, s->d_tmp1, nb, -1,
                      s->s_fft)) {
        set_error("forward FFT launch failed");
        return -1;
      }
    }
  }
  K9_CUDA(cudaEventRecord(s->e_fft1, s->s_fft));
  s->fft_pending = true;
  return 0;
}

// the spectrum buffer being written is complete (FFT and, on multi-GPU runs, the collective): hand it to the channels
int publish_spectrum(ka9q_stream* s) {
  K9_CUDA(cudaEventRecord(s->e_spec_ready[s->spec_wr], s->pub_stream ? s->pub_stream : s->s_fft));
  s->pub_stream = nullptr;
  s->spec_wr = (s->spec_wr + 1) % s->nspec;
  s->spec_published++;
  s->fft_pending = false;
  return 0;
}

int issue_channels(ka9q_stream* s, int nblocks) {
  if (s->fft_pending && publish_spectrum(s)) return -1;
  K9_CHECK(s->spec_published > 0, "no spectrum has been computed or received for this batch");
  s->spec_published--;
  const int p = s->spec_rd;
  K9_CUDA(cudaStreamWaitEvent(s->s_comp, s->e_spec_ready[p], 0));
  K9_CUDA(cudaStreamWaitEvent(s->s_comp, s->e_fetched[s->comp_parity ^ 1], 0));  // that PCM buffer's results have left
  K9_CHECK(mgpu_wait_ready(s, p) == 0, "multi-GPU wait kernel launch failed");     // peers' arcs have landed (P2P transport)
  ChanLaunch a;
  fill_launch(s, a, nblocks);
  a.spec = spec_buf(s, p);
  // the three demodulator families run concurrently on sibling streams
  K9_CUDA(cudaEventRecord(s->e_fork, s->s_comp));
  if (s->n_am) {
    K9_CUDA(cudaStreamWaitEvent(s->s_am, s->e_fork, 0));
    a.work = s->d_work_am;
    a.nwork = s->n_am;
    a.agc_x = s->d_agc_x_am;
    a.agc_y = nullptr;
    a.agc_pow = s->d_agc_pow;
    {
      TimedRegion tr(s, TC_AM, s->s_am);
      K9_CHECK(launch_am(a, s->s_am) == 0, "am kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    K9_CUDA(cudaEventRecord(s->e_am, s->s_am));
  }
  if (s->n_lin) {
    K9_CUDA(cudaStreamWaitEvent(s->s_lin, s->e_fork, 0));
    a.work = s->d_work_lin;
    a.nwork = s->n_lin;
    a.agc_x = s->d_agc_x_lin;
    a.agc_y = s->d_agc_y_lin;
    a.agc_pow = s->d_agc_pow + 2 * (size_t)s->cfg.max_blocks * s->n_am;
    {
      TimedRegion tr(s, TC_LIN, s->s_lin);
      K9_CHECK(launch_linear(a, s->s_lin) == 0, "linear kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    K9_CUDA(cudaEventRecord(s->e_lin, s->s_lin));
  }
  if (s->n_pll) {
    K9_CUDA(cudaStreamWaitEvent(s->s_pll, s->e_fork, 0));
    a.work = s->d_work_pll;
    a.nwork = s->n_pll;
    a.pll_params = s->d_pll_params;
    a.pll_state = s->d_pll_state;
    a.pll_ring = s->d_pll_ring;
    {
      TimedRegion tr(s, TC_LIN, s->s_pll);
      K9_CHECK(launch_pll(a, s->s_pll) == 0, "pll kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    K9_CUDA(cudaEventRecord(s->e_pll, s->s_pll));
  }
  if (s->n_fm) {
    a.work = s->d_work_fm;
    a.nwork = s->n_fm;
    a.fm_seq = s->d_fm_seq;
    if (s->n_pl) {
      a.pl_spec = s->d_pl_spec;
      a.pl_npairs = s->n_fm;
      a.pl_work = s->d_pl_work;
      a.pl_state = s->d_pl_state;
      a.pl_ring = s->d_pl_ring;
      a.pl_resp = s->d_pl_resp;
    }
    {
      TimedRegion tr(s, TC_FM, s->s_comp);
      K9_CHECK(launch_fm(a, s->s_comp, s->n_am + s->n_lin + s->n_pll > 0) == 0, "fm kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (s->n_pl) K9_CHECK(launch_pl(a, s->n_pl, s->s_comp) == 0, "pl kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (s->n0_enabled) {
    // K6 beside the channel kernels: reads the same spectrum buffer, writes the n0 rows of this batch
    K9_CUDA(cudaStreamWaitEvent(s->s_n0, s->e_fork, 0));
    N0Launch n;
    n.spec = a.spec;
    n.spec_stride = s->N;
    n.N = s->N;
    n.samprate = s->cfg.samprate;
    n.nblocks = nblocks;
    n.nchan = (int)s->chans.size();
    n.chan = s->d_n0_chan;
    n.P = s->d_n0_P;
    n.partial = s->d_n0_partial;
    n.blk = s->d_n0_blk;
    n.T = s->d_n0_T;
    n.list = s->d_n0_list;
    n.list_cap = N0_LIST_CAP;
    const int wbuf = s->comp_parity ^ 1;
    n.n0_raw = s->d_n0_raw + (size_t)wbuf * s->cfg.max_blocks * s->chans.size();
    n.n0_smooth = s->d_n0_smooth + (size_t)wbuf * s->cfg.max_blocks * s->chans.size();
    n.state = s->d_n0_state;
    K9_CHECK(n0_launch(n, s->s_n0) == 0, "n0 kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    K9_CUDA(cudaEventRecord(s->e_n0, s->s_n0));
    K9_CUDA(cudaStreamWaitEvent(s->s_comp, s->e_n0, 0));
  }
  if (s->n_am) K9_CUDA(cudaStreamWaitEvent(s->s_comp, s->e_am, 0));
  if (s->n_lin) K9_CUDA(cudaStreamWaitEvent(s->s_comp, s->e_lin, 0));
  if (s->n_pll) K9_CUDA(cudaStreamWaitEvent(s->s_comp, s->e_pll, 0));
  K9_CUDA(cudaEventRecord(s->e_chan1, s->s_comp));
  K9_CUDA(cudaEventRecord(s->e_spec_free[p], s->s_comp));
  s->spec_rd = (s->spec_rd + 1) % s->nspec;
  s->phase_block += nblocks;
  s->comp_parity ^= 1;
  K9_CUDA(cudaEventRecord(s->e_comp_done[s->comp_parity], s->s_comp));
  s->last_nblocks = nblocks;
  return 0;
}

// which blocks of the ring a batch covers: streaming mode consumes the next unprocessed blocks, resident mode re-runs
// the most recently pushed ones (benchmarks with the input already in HBM)
static int batch_first_block(ka9q_stream* s, int nblocks, bool resident, long long* first_block) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks, "nblocks out of range");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (resident) {
    *first_block = s->pushed / s->cfg.L - nblocks;
    K9_CHECK(*first_block >= 0, "not enough samples resident in the ring");
    if (s->phase_block < *first_block) s->phase_block = *first_block;
  } else {
    *first_block = s->block0;
    K9_CHECK((*first_block + nblocks) * (long long)s->cfg.L <= s->pushed, "compute ahead of pushed samples");
  }
  return 0;
}

static int compute_impl(ka9q_stream* s, int nblocks, bool resident, bool do_fft, bool do_chan, int blk_first = 0,
                        int blk_count = -1) {
  long long first_block = 0;
  if (batch_first_block(s, nblocks, resident, &first_block)) return -1;
  if (blk_count < 0) blk_count = nblocks - blk_first;
  K9_CHECK(blk_first >= 0 && blk_first + blk_count <= nblocks, "block range out of the batch");
  if (do_fft && issue_fft(s, first_block, blk_first, blk_count)) return -1;
  if (do_fft && do_chan && publish_spectrum(s)) return -1;
  if (do_chan && issue_channels(s, nblocks)) return -1;
  if (do_fft || !resident) {
    if (resident)
      s->block0 = s->pushed / s->cfg.L;  // everything pushed so far counts as consumed
    else if (do_chan || !do_fft)
      s->block0 += nblocks;
  }
  return 0;
}

int ka9q_stream_compute(ka9q_stream* s, int nblocks) { return compute_impl(s, nblocks, false, true, true); }
int ka9q_stream_compute_resident(ka9q_stream* s, int nblocks) { return compute_impl(s, nblocks, true, true, true); }
int ka9q_stream_compute_fft_only(ka9q_stream* s, int nblocks) { return compute_impl(s, nblocks, true, true, false); }
int ka9q_stream_compute_fft_blocks(ka9q_stream* s, int nblocks, int first, int count) {
  return compute_impl(s, nblocks, true, true, false, first, count);
}
int ka9q_stream_compute_channels_only(ka9q_stream* s, int nblocks) { return compute_impl(s, nblocks, true, false, true); }

int ka9q_stream_fetch(ka9q_stream* s, int nblocks, int16_t* pcm, ka9q_chan_status* status) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks, "nblocks out of range");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamWaitEvent(s->s_out, s->e_comp_done[s->comp_parity], 0));
  if (pcm)
    K9_CUDA(cudaMemcpyAsync(pcm, s->d_pcm + (size_t)s->comp_parity * s->cfg.max_blocks * s->pcm_stride,
                            sizeof(int16_t) * (size_t)nblocks * s->pcm_stride, cudaMemcpyDeviceToHost,
                            s->s_out));
  if (status)
    K9_CUDA(cudaMemcpyAsync(status, s->d_status + (size_t)s->comp_parity * s->cfg.max_blocks * s->chans.size(),
                            sizeof(ChanStatus) * (size_t)nblocks * s->chans.size(),
                            cudaMemcpyDeviceToHost, s->s_out));
  K9_CUDA(cudaEventRecord(s->e_fetched[s->comp_parity], s->s_out));
  return 0;
}

// Wait only for the D2H copies issued by ka9q_stream_fetch so far (the compute of the next batch keeps running).
int ka9q_stream_wait_fetch(ka9q_stream* s) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamSynchronize(s->s_out));
  return 0;
}

int ka9q_stream_wait_fetched(ka9q_stream* s, int batches_ago) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CHECK(batches_ago == 0 || batches_ago == 1, "batches_ago must be 0 or 1 (the device PCM is double-buffered)");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  // e_fetched[p] is recorded behind the copies of the batch computed into PCM buffer p; comp_parity is the latest batch's
  K9_CUDA(cudaEventSynchronize(s->e_fetched[s->comp_parity ^ batches_ago]));
  return 0;
}

// Wait until every H2D copy issued by push / push_at has finished reading its host buffer (the compute may still run).
int ka9q_stream_sync_input(ka9q_stream* s) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamSynchronize(s->s_in));
  return 0;
}

int ka9q_stream_sync(ka9q_stream* s) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamSynchronize(s->s_in));
  K9_CUDA(cudaStreamSynchronize(s->s_fft));
  if (s->s_mgx) K9_CUDA(cudaStreamSynchronize(s->s_mgx));
  if (s->s_mgwait) K9_CUDA(cudaStreamSynchronize(s->s_mgwait));
  K9_CUDA(cudaStreamSynchronize(s->s_comp));
  if (s->s_mgsig) K9_CUDA(cudaStreamSynchronize(s->s_mgsig));
  K9_CUDA(cudaStreamSynchronize(s->s_am));
  K9_CUDA(cudaStreamSynchronize(s->s_lin));
  if (s->s_n0) K9_CUDA(cudaStreamSynchronize(s->s_n0));
  if (s->s_pll) K9_CUDA(cudaStreamSynchronize(s->s_pll));
  K9_CUDA(cudaStreamSynchronize(s->s_out));
  return 0;
}

int ka9q_stream_process(ka9q_stream* s, const void* iq, int nblocks, int16_t* pcm, ka9q_chan_status* status) {
  static_assert(sizeof(ka9q_chan_status) == sizeof(ChanStatus), "status layout");
  if (ka9q_stream_push(s, iq, nblocks)) return -1;
  if (ka9q_stream_compute(s, nblocks)) return -1;
  if (ka9q_stream_fetch(s, nblocks, pcm, status)) return -1;
  return ka9q_stream_sync(s);
}

int ka9q_stream_last_timing(ka9q_stream* s, float* total_ms, float* fft_ms, float* chan_ms) {
  K9_CHECK(s && s->committed, "stream not committed");
  float f = 0, c = 0;
  K9_CUDA(cudaEventElapsedTime(&f, s->e_fft0, s->e_fft1));
  K9_CUDA(cudaEventElapsedTime(&c, s->e_fft1, s->e_chan1));
  if (fft_ms) *fft_ms = f;
  if (chan_ms) *chan_ms = c;
  if (total_ms) *total_ms = f + c;
  return 0;
}

int ka9q_stream_spectrum_ptr(ka9q_stream* s, void** dev_ptr, long long* bytes_per_block) {
  K9_CHECK(s && s->committed, "stream not committed");
  // the buffer compute_fft_only just wrote (not yet handed to the channel kernels), else the one the last launch read
  if (dev_ptr) *dev_ptr = spec_buf(s, s->fft_pending ? s->spec_wr : (s->spec_rd + s->nspec - 1) % s->nspec);
  if (bytes_per_block) *bytes_per_block = (long long)sizeof(float2) * s->N;
  return 0;
}

int ka9q_stream_get_response(ka9q_stream* s, int chan, void* out2048, float* noise_gain) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CHECK(chan >= 0 && chan < (int)s->chans.size(), "bad channel");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (ka9q_stream_sync(s)) return -1;  // the work streams are non-blocking: order behind everything in flight
  if (out2048)
    K9_CUDA(cudaMemcpy(out2048, s->d_resp + (size_t)s->h_params[chan].resp_slot * NDEC, sizeof(float2) * NDEC,
                       cudaMemcpyDeviceToHost));
  if (noise_gain) *noise_gain = s->h_noise_gain[chan];
  return 0;
}

int ka9q_stream_get_filter_output(ka9q_stream* s, int chan, int nblocks, void* out) {
  K9_CHECK(s && s->committed && s->d_filt, "filter-output capture not enabled");
  K9_CHECK(chan >= 0 && chan < (int)s->chans.size() && nblocks >= 1 && nblocks <= s->cfg.max_blocks, "bad argument");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (ka9q_stream_sync(s)) return -1;  // the work streams are non-blocking: order behind everything in flight
  const int K = (int)s->chans.size();
  K9_CUDA(cudaMemcpy2D(out, sizeof(float2) * s->olen, s->d_filt + (size_t)chan * s->olen, sizeof(float2) * (size_t)K * s->olen,
                       sizeof(float2) * s->olen, nblocks, cudaMemcpyDeviceToHost));
  return 0;
}

int ka9q_stream_get_spectrum(ka9q_stream* s, int block, void* outN) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CHECK(block >= 0 && block < s->cfg.max_blocks, "bad block");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (ka9q_stream_sync(s)) return -1;
  K9_CUDA(cudaMemcpy(outN, spec_buf(s, s->fft_pending ? s->spec_wr : (s->spec_rd + s->nspec - 1) % s->nspec) + (size_t)block * s->N, sizeof(float2) * s->N, cudaMemcpyDeviceToHost));
  return 0;
}

int ka9q_stream_get_if_energy(ka9q_stream* s, int nblocks, float* energy) {
  K9_CHECK(s && s->committed && energy, "bad argument");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (ka9q_stream_sync(s)) return -1;  // the work streams are non-blocking: order behind everything in flight
  K9_CUDA(cudaMemcpy(energy, s->d_energy, sizeof(float) * nblocks, cudaMemcpyDeviceToHost));
  return 0;
}

// Frees every device / host / stream / event resource of a stream (idempotent). Used by destroy and by the failure paths
// of commit, so a stream that could not be committed leaks nothing.
static void release_resources(ka9q_stream* s) {
  cudaSetDevice(s->cfg.device);
  cudaDeviceSynchronize();
  mgpu_release(s);
  if (s->nccl_comm && p_ncclCommDestroy) p_ncclCommDestroy((k9_ncclComm_t)s->nccl_comm);
  s->nccl_comm = nullptr;
  void** dev[] = {(void**)&s->d_agc_x_am, (void**)&s->d_agc_x_lin, (void**)&s->d_agc_y_lin, (void**)&s->d_agc_pow, &s->d_ring,
                  (void**)&s->d_spec, (void**)&s->d_tmp0, (void**)&s->d_tmp1, (void**)&s->d_energy, (void**)&s->d_tw2048,
                  (void**)&s->d_params, (void**)&s->d_state, (void**)&s->d_resp, (void**)&s->d_audio_resp,
                  (void**)&s->d_audio_hist, (void**)&s->d_pcm, (void**)&s->d_status, (void**)&s->d_filt,
                  (void**)&s->d_windows, (void**)&s->d_work_fm, (void**)&s->d_work_am, (void**)&s->d_work_lin,
                  (void**)&s->d_pl_spec, (void**)&s->d_pl_resp, (void**)&s->d_pl_work, (void**)&s->d_pl_state, (void**)&s->d_pl_ring,
                  (void**)&s->d_work_pll, (void**)&s->d_pll_params, (void**)&s->d_pll_state, (void**)&s->d_pll_ring,
                  (void**)&s->d_n0_chan, (void**)&s->d_n0_P, (void**)&s->d_n0_T, (void**)&s->d_n0_list, (void**)&s->d_n0_raw,
                  (void**)&s->d_n0_smooth, (void**)&s->d_n0_state, (void**)&s->d_n0_partial, (void**)&s->d_n0_blk,
                  (void**)&s->d_fm_seq};
  for (void** p : dev) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  void** host[] = {&s->h_iq, (void**)&s->h_pcm, (void**)&s->h_status};
  for (void** p : host) {
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
  }
  cudaStream_t* sts[] = {&s->s_in, &s->s_comp, &s->s_out, &s->s_fm, &s->s_am, &s->s_lin, &s->s_fft, &s->s_n0, &s->s_pll,
                         &s->s_mgwait, &s->s_mgsig, &s->s_mgx};
  for (auto st : sts) {
    if (*st) cudaStreamDestroy(*st);
    *st = nullptr;
  }
  cudaEvent_t* evs[] = {&s->e_pushed, &s->e_fft0, &s->e_fft1, &s->e_chan1, &s->e_fork, &s->e_am, &s->e_lin, &s->e_fm,
                        &s->e_comp_done[0], &s->e_comp_done[1], &s->e_fetched[0], &s->e_fetched[1], &s->e_spec_ready[0],
                        &s->e_spec_ready[1], &s->e_spec_ready[2], &s->e_spec_free[0], &s->e_spec_free[1], &s->e_spec_free[2], &s->e_t0, &s->e_t1, &s->e_n0, &s->e_pll,
                        &s->e_mg_ready, &s->e_mg_chan, &s->e_mg_fft};
  for (auto e : evs) {
    if (*e) cudaEventDestroy(*e);
    *e = nullptr;
  }
  for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
  s->ev_pool.clear();
  s->ev_used.clear();
  s->ev_next = 0;
  s->committed = false;
}

int ka9q_stream_destroy(ka9q_stream* s) {
  if (!s) return 0;
  release_resources(s);
  bigfft_plan_destroy(&s->fwd);
  bigfft_plan_destroy(&s->p2048);
  delete s;
  return 0;
}

// Overlap of the forward FFT (batch k+1) with the channel kernels (batch k) is on by default; switching it off
// serialises them so that per-kernel event timings are those of each kernel running alone.
int ka9q_stream_set_overlap(ka9q_stream* s, int enable) {
  K9_CHECK(s, "null argument");
  s->overlap = enable != 0;
  return 0;
}

// ------------------------------------------------------------------ timed region (bench.py)
// timer_start/stop bracket a region with CUDA events on the compute stream (the stream every kernel of this library
// is launched on or joined into); while active, every forward FFT and every channel-kernel launch is also bracketed
// by its own event pair so the per-kernel-class device time inside the region can be reported.
int ka9q_stream_timer_start(ka9q_stream* s) {
  K9_CHECK(s && s->committed, "stream not committed");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (!s->e_t0) {
    K9_CUDA(cudaEventCreate(&s->e_t0));
    K9_CUDA(cudaEventCreate(&s->e_t1));
  }
  s->ev_used.clear();
  s->ev_next = 0;
  s->timing = true;
  s->timing_regions = true;
  K9_CUDA(cudaEventRecord(s->e_t0, s->s_comp));
  return 0;
}
// the outer event pair only: nothing is recorded around the individual kernels (their event records cost a few
// microseconds per step, which shows at 8-GPU step times); timer_stop then reports zeros for the classes
int ka9q_stream_timer_start_plain(ka9q_stream* s) {
  if (ka9q_stream_timer_start(s)) return -1;
  s->timing_regions = false;
  return 0;
}
// ms_total: region time. class_ms[5] / class_launches[5]: summed device time and launch count of
// {forward FFT (all passes), FM kernel, AM kernel, linear kernel, NCCL spectrum broadcast}.
int ka9q_stream_timer_stop(ka9q_stream* s, float* ms_total, float* class_ms, int* class_launches) {
  K9_CHECK(s && s->committed && s->timing, "timer not started");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaEventRecord(s->e_t1, s->s_comp));
  s->timing = false;
  if (ka9q_stream_sync(s)) return -1;
  float ms = 0;
  K9_CUDA(cudaEventElapsedTime(&ms, s->e_t0, s->e_t1));
  if (ms_total) *ms_total = ms;
  float acc[TC_COUNT] = {0, 0, 0, 0, 0};
  int cnt[TC_COUNT] = {0, 0, 0, 0, 0};
  for (auto& u : s->ev_used) {
    float t = 0;
    if (u.first >= TC_COUNT) continue;  // timeline-only regions (waits on peer flags)
    if (cudaEventElapsedTime(&t, u.second.first, u.second.second) == cudaSuccess) {
      acc[u.first] += t;
      cnt[u.first]++;
    }
  }
  for (int i = 0; i < TC_COUNT; i++) {
    if (class_ms) class_ms[i] = acc[i];
    if (class_launches) class_launches[i] = cnt[i];
  }
  return 0;
}

// After timer_stop: the bracketed regions of the timed interval in issue order, as (class, start, end) in ms from
// timer_start. Classes as in timer_stop, plus 5 = waiting for peer flags (multi-GPU). `start` is when the region's stream
// reached it (its earlier work had finished), `end` when its last kernel had finished. Returns the number of regions.
int ka9q_stream_timer_timeline(ka9q_stream* s, int max_regions, int* cls, float* start_ms, float* end_ms) {
  K9_CHECK(s && s->committed && !s->timing && s->e_t0, "call after ka9q_stream_timer_stop");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  int n = 0;
  for (auto& u : s->ev_used) {
    if (n >= max_regions) break;
    float a = 0, b = 0;
    if (cudaEventElapsedTime(&a, s->e_t0, u.second.first) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&b, s->e_t0, u.second.second) != cudaSuccess) continue;
    cls[n] = u.first;
    start_ms[n] = a;
    end_ms[n] = b;
    n++;
  }
  return n;
}

// ------------------------------------------------------------------ pinned host memory

void* ka9q_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 64, cudaHostAllocDefault) != cudaSuccess) {
    set_error("ka9q_host_alloc: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
void ka9q_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ NCCL spectrum broadcast (libnccl dlopen'ed)
// The only inter-GPU exchange of the path: the rank that ingested the block broadcasts N*8 bytes of
// spectrum per block; channels never move.

int (*p_ncclGetUniqueId)(k9_ncclUniqueId*);
int (*p_ncclCommInitRank)(k9_ncclComm_t*, int, k9_ncclUniqueId, int);
int (*p_ncclBroadcast)(const void*, void*, size_t, int, int, k9_ncclComm_t, cudaStream_t);
int (*p_ncclCommDestroy)(k9_ncclComm_t);
int (*p_ncclAllGather)(const void*, void*, size_t, int, k9_ncclComm_t, cudaStream_t);
int (*p_ncclSend)(const void*, size_t, int, int, k9_ncclComm_t, cudaStream_t);
int (*p_ncclRecv)(void*, size_t, int, int, k9_ncclComm_t, cudaStream_t);
int (*p_ncclGroupStart)(void);
int (*p_ncclGroupEnd)(void);
const char* (*p_ncclGetErrorString)(int);

int load_nccl() {
  static int state = 0;  // 0 untried, 1 ok, -1 failed
  if (state) return state > 0 ? 0 : -1;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    state = -1;
    set_error("cannot load libnccl.so.2: %s", dlerror());
    return -1;
  }
  p_ncclGetUniqueId = (int (*)(k9_ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
  p_ncclCommInitRank = (int (*)(k9_ncclComm_t*, int, k9_ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
  p_ncclBroadcast = (int (*)(const void*, void*, size_t, int, int, k9_ncclComm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
  p_ncclCommDestroy = (int (*)(k9_ncclComm_t))dlsym(h, "ncclCommDestroy");
  p_ncclAllGather = (int (*)(const void*, void*, size_t, int, k9_ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
  p_ncclGetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  p_ncclSend = (int (*)(const void*, size_t, int, int, k9_ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
  p_ncclRecv = (int (*)(void*, size_t, int, int, k9_ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
  p_ncclGroupStart = (int (*)(void))dlsym(h, "ncclGroupStart");
  p_ncclGroupEnd = (int (*)(void))dlsym(h, "ncclGroupEnd");
  if (!p_ncclGetUniqueId || !p_ncclCommInitRank || !p_ncclBroadcast) {
    state = -1;
    set_error("libnccl.so.2 lacks required symbols");
    return -1;
  }
  state = 1;
  return 0;
}

int ka9q_nccl_unique_id(void* id128) {
  K9_CHECK(id128, "null argument");
  if (load_nccl()) return -1;
  k9_ncclUniqueId id;
  int r = p_ncclGetUniqueId(&id);
  K9_CHECK(r == 0, "ncclGetUniqueId: %s", p_ncclGetErrorString ? p_ncclGetErrorString(r) : "error");
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int ka9q_stream_nccl_init(ka9q_stream* s, const void* id128, int rank, int nranks) {
  K9_CHECK(s && s->committed && id128, "bad argument");
  if (load_nccl()) return -1;
  K9_CUDA(cudaSetDevice(s->cfg.device));
  k9_ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  k9_ncclComm_t comm = nullptr;
  int r = p_ncclCommInitRank(&comm, nranks, id, rank);
  K9_CHECK(r == 0, "ncclCommInitRank: %s", p_ncclGetErrorString ? p_ncclGetErrorString(r) : "error");
  s->nccl_comm = comm;
  s->nccl_rank = rank;
  s->nccl_nranks = nranks;
  return 0;
}

// Both collectives run on the FFT stream right behind the forward transform(s) of the batch being written and then
// publish that spectrum buffer to the channel kernels, so they overlap the channel work of the previous batch.
int ka9q_stream_nccl_broadcast_spectrum(ka9q_stream* s, int nblocks, int root) {
  K9_CHECK(s && s->committed && s->nccl_comm, "NCCL communicator not initialised");
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks, "nblocks out of range");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamWaitEvent(s->s_fft, s->e_spec_free[s->spec_wr], 0));
  float2* buf = spec_buf(s, s->spec_wr);
  const size_t count = (size_t)2 * s->N * nblocks;  // float32 count: 2 floats per bin
  int r;
  {
    TimedRegion tr(s, TC_BCAST, s->s_fft);
    r = p_ncclBroadcast(buf, buf, count, /*ncclFloat32=*/7, root, (k9_ncclComm_t)s->nccl_comm, s->s_fft);
  }
  K9_CHECK(r == 0, "ncclBroadcast: %s", p_ncclGetErrorString ? p_ncclGetErrorString(r) : "error");
  return publish_spectrum(s);
}

// Forward FFT sharded by block: rank r transformed blocks [r*nblocks/nranks, (r+1)*nblocks/nranks) of the batch
// (ka9q_stream_compute_fft_blocks; every rank holds the int16 input, which is 4x smaller than the spectrum); the
// all-gather completes every rank's copy of all nblocks spectra. nblocks must be a multiple of nranks.
int ka9q_stream_nccl_allgather_spectrum(ka9q_stream* s, int nblocks) {
  K9_CHECK(s && s->committed && s->nccl_comm && p_ncclAllGather, "NCCL communicator not initialised");
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks && nblocks % s->nccl_nranks == 0,
           "nblocks must be a multiple of the number of ranks");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  K9_CUDA(cudaStreamWaitEvent(s->s_fft, s->e_spec_free[s->spec_wr], 0));
  float2* buf = spec_buf(s, s->spec_wr);
  const size_t per_rank = (size_t)2 * s->N * (nblocks / s->nccl_nranks);
  int r;
  {
    TimedRegion tr(s, TC_BCAST, s->s_fft);
    r = p_ncclAllGather((const float*)buf + per_rank * s->nccl_rank, buf, per_rank, /*ncclFloat32=*/7,
                        (k9_ncclComm_t)s->nccl_comm, s->s_fft);
  }
  K9_CHECK(r == 0, "ncclAllGather: %s", p_ncclGetErrorString ? p_ncclGetErrorString(r) : "error");
  return publish_spectrum(s);
}

// ------------------------------------------------------------------ generic FFT on host buffers (tests / cross-checks)

int ka9q_fft_plan_describe(int n, int* sizes) {
  int R1[4], R2[4];
  int np = bigfft_factorize(n, R1, R2);
  if (np < 0) return -1;
  for (int i = 0; i < np; i++)
    if (sizes) sizes[i] = R1[i] * R2[i];
  return np;
}

int ka9q_fft_c2c(int device, int n, int batch, int sign, const void* in, void* out) {
  K9_CHECK(in && out && n >= 16 && batch >= 1, "bad argument");
  K9_CHECK(ka9q_device_count() > device && device >= 0, "CUDA device %d not available (no CPU fallback exists)", device);
  K9_CUDA(cudaSetDevice(device));
  BigFftPlan plan;
  K9_CHECK(bigfft_plan_create(&plan, n) == 0, "FFT size %d unsupported", n);
  float2 *d_in = nullptr, *d_out = nullptr, *d_t0 = nullptr, *d_t1 = nullptr;
  const size_t bytes = sizeof(float2) * (size_t)n * batch;
  K9_CUDA(cudaMalloc(&d_in, bytes));
  K9_CUDA(cudaMalloc(&d_out, bytes));
  K9_CUDA(cudaMalloc(&d_t0, bytes));
  K9_CUDA(cudaMalloc(&d_t1, bytes));
  K9_CUDA(cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice));
  BigFftIn bi;
  bi.in = d_in;
  bi.in_batch_stride = n;
  int r = bigfft_exec(&plan, bi, d_out, n, d_t0, d_t1, batch, sign, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (r == 0 && e == cudaSuccess) e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  cudaFree(d_t0);
  cudaFree(d_t1);
  bigfft_plan_destroy(&plan);
  K9_CHECK(r == 0 && e == cudaSuccess, "fft failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
