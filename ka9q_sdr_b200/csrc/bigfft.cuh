// K1+K2: large batched 1-D complex FFT (N = 2^a 3^b 5^c up to ~2^31) as 1..4 Stockham autosort passes over HBM/L2,
// each pass computing T independent R-point column FFTs per CTA in registers + one shared-memory exchange.
//
// Replaces fftwf_execute(master->fwd_plan) (reference filter.c:151) and, with the int16 ring input mode, the
// sample ingest of proc_samples (reference radio.c:106-123: int16 * SCALE16 * gain_factor) plus the
// overlap-save window assembly / memmove (filter.c:159-170): the first pass reads the N-sample window
// [m*L-(M-1), m*L+L) straight out of the device I/Q ring, so no fp32 time-domain buffer ever exists in HBM.
//
// Pass definition (decimation in frequency, autosort): for current length n, stride s (n*s == N), radix R, m=n/R:
//   y[q + s*(R*p + j)] = w_n^(p*j) * sum_r x[q + s*(p + m*r)] * W_R^(r*j),   q in [0,s), p in [0,m)
// Column index c = q + s*p runs over [0, N/R); x index = c + (N/R)*r, so loads are coalesced across columns.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace k9 {

enum InMode : int {
  IN_C32 = 0,       // float2 array, natural order (intermediate passes, generic FFT)
  IN_RING_S16 = 1,  // int16 I/Q ring (RTP IQ_PT payload, reference radio.c:113-114)
  IN_RING_S8 = 2,   // int8 I/Q ring (IQ_PT8, radio.c:116-117)
  IN_RING_C32 = 3,  // float2 ring (drop-in filter_in with COMPLEX input)
  IN_RING_R32 = 4,  // float ring (drop-in filter_in with REAL input; imaginary part = 0)
};

struct PassArgs {
  const void* in;        // IN_C32: float2[batch][N]; ring modes: ring base
  float2* out;           // float2[batch][N]
  long long in_batch_stride;   // elements, IN_C32 only
  long long out_batch_stride;  // elements
  int N;                 // transform length
  int n_cur;             // current Stockham sub-length (n); stride s = N / n_cur
  int ncols;             // N / R
  int in_mode;
  // ring modes (first pass only)
  long long ring_cap;    // samples in ring
  long long ring_off;    // ring position of logical index 0 for batch 0
  long long ring_step;   // added per batch (= L)
  float scale;           // sample scale (SCALE16 / SCALE8), applied first
  float gain;            // sdr.gain_factor, applied second (radio.c:122)
  int stat_from;         // logical index where the "new" L samples start (M-1); energy accumulated from here
  float* energy;         // [batch] sum |x|^2 over new samples (if_power numerator, radio.c:123), may be null
  // twiddles
  const float2* tw_lo;   // W_N^a, a in [0,1024)           (forward sign; conjugated for SIGN=+1)
  const float2* tw_hi;   // W_N^(1024 b), b in [0, ceil(N/1024))
  const float2* tw_r;    // W_R^a, a in [0,R)
  // Multi-GPU routing of the LAST pass (lean 160-point kernel only; null = plain store to `out`): every 128-byte output
  // row (16 bins) goes to the ranks whose channels read it — route_mask[bin / 16] has one bit per rank — at the same
  // offset inside that rank's spectrum allocation: (char*)(out + idx) + route_delta[rank]; route_delta[own rank] = 0.
  const unsigned short* route_mask;
  long long route_delta[16];
};

struct BigFftPlan {
  int N = 0;
  int npass = 0;
  int R1[4] = {0, 0, 0, 0}, R2[4] = {0, 0, 0, 0};
  float2* tw_lo = nullptr;
  float2* tw_hi = nullptr;
  float2* tw_r[4] = {nullptr, nullptr, nullptr, nullptr};
  int device = 0;
};

// Host API (bigfft.cu)
// Returns 0 on success; -1 if N has no supported factorisation.
int bigfft_plan_create(BigFftPlan* plan, int N);
void bigfft_plan_destroy(BigFftPlan* plan);
// Describe how N would be factorised: fills R[] with per-pass sizes, returns number of passes or -1.
int bigfft_factorize(int N, int* R1, int* R2);

struct BigFftIn {
  int in_mode = IN_C32;
  const void* in = nullptr;
  long long in_batch_stride = 0;
  long long ring_cap = 0, ring_off = 0, ring_step = 0;
  float scale = 1.f, gain = 1.f;
  int stat_from = 0;
  float* energy = nullptr;
  // multi-GPU routing of the last pass (see PassArgs); honoured only when bigfft_can_route(plan)
  const unsigned short* route_mask = nullptr;
  long long route_delta[16] = {0};
};
// Executes `batch` transforms. tmp0/tmp1: scratch float2[batch][N] (tmp1 only needed when npass >= 3; for IN_C32 input
// with npass >= 2 the input is NOT modified). sign: -1 forward, +1 backward (unnormalised).
int bigfft_exec(const BigFftPlan* plan, const BigFftIn& in, float2* out, long long out_batch_stride, float2* tmp0,
                float2* tmp1, int batch, int sign, cudaStream_t stream);
// preferred shared-memory carve-out (percent, or cudaSharedmemCarveoutMaxShared) of the lean pass kernels on the current
// device: set it to the channel kernels' so that the two can share SMs (see bigfft_r128.cuh)
void bigfft_set_carveout(int pct);
// true if the plan's last pass is the lean 160-point kernel, which can store its output rows straight into peer memory
bool bigfft_can_route(const BigFftPlan* plan);
// number of kernel launches one bigfft_exec performs
inline int bigfft_launches(const BigFftPlan* plan) { return plan->npass; }

}  // namespace k9
