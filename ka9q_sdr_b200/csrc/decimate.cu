// K5: half-band decimators on the device (reference decimate.c:44-162, driven as hackrf.c:297-318).
//
//   hb15: y[m] = x[2m-6] + sum_{i<4} c[i] * (x[2m+1-2i] + x[2m-13+2i])     (15-tap folded half-band, unity centre tap)
//   hb3 : y[m] = 2*x[2m] + x[2m+1] + x[2m-1]                                 (decimate.c:148-162)
// State layout is the reference's struct hb15_state (decimate.h:4-9): after a call
//   even_samples[1..3] = x[-2], x[-4], x[-6]; odd_samples[1..3] = x[-1], x[-3], x[-5];
//   old_odd_samples[3..0] = x[-7], x[-9], x[-11], x[-13]   (indices relative to the next call's x[0]).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/ka9q_b200.h"
#include "util.cuh"

using namespace k9;

namespace {

// hist[k] = x[-k], k = 1..13 (hist[0] unused)
__global__ void hb15_kernel(const float* __restrict__ x, const float* __restrict__ hist, float c0, float c1, float c2,
                            float c3, float* __restrict__ y, int cnt) {
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < cnt; m += gridDim.x * blockDim.x) {
    auto X = [&](int i) -> float { return i >= 0 ? x[i] : hist[-i]; };
    const int b = 2 * m;
    // same association as the portable reference loop: centre, then taps from the tails inwards (decimate.c:124-128)
    float r = X(b - 6);
    r += (X(b + 1) + X(b - 13)) * c0;
    r += (X(b - 1) + X(b - 11)) * c1;
    r += (X(b - 3) + X(b - 9)) * c2;
    r += (X(b - 5) + X(b - 7)) * c3;
    y[m] = r;
  }
}

__global__ void hb3_kernel(const float* __restrict__ x, float prev, float* __restrict__ y, int cnt) {
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < cnt; m += gridDim.x * blockDim.x) {
    const float xm1 = m > 0 ? x[2 * m - 1] : prev;
    y[m] = 2 * x[2 * m] + x[2 * m + 1] + xm1;
  }
}

// Whole cascade in one launch (stages <= 6, the hackrf /64 case): each CTA produces FUSED_C final outputs and recomputes
// the halo it needs at every level in shared memory (out[m] needs in[2m-13 .. 2m+1], so a chunk [a, b) of a level's
// output needs [2a-13, 2b) of its input). Level l = 0 is the highest rate; its state is states[S-1-l]
// (hackrf.c:297-301). Samples at negative indices of a level come from that level's carried history, exactly as in the
// per-stage kernel, and every output is formed by the same expression, so the result is bit-identical.
constexpr int FUSED_MAXS = 6;
constexpr int FUSED_C = 64;
constexpr int FUSED_BUF = (FUSED_C << FUSED_MAXS) + 13 * ((1 << FUSED_MAXS) - 1) + 8;

struct FusedArgs {
  const float* x;     // level-0 input, n_in samples
  const float* hist;  // [S][16]: hist[l][k] = x_l[-k], k = 1..13
  float4 coeff[FUSED_MAXS];
  int S, n_in;
  float* y;     // n_in >> S outputs
  float* tail;  // [S][16]: tail[l][k] = last samples of (history ++ input) of level l, k = 1..13
};

__global__ void __launch_bounds__(256) hb15_cascade_kernel(const FusedArgs a) {
  __shared__ float buf[2][FUSED_BUF];
  __shared__ float hs[FUSED_MAXS][16];
  const int t = threadIdx.x;
  const int S = a.S;
  if (t < 16 * S) hs[t >> 4][t & 15] = a.hist[t];
  const int n_out = a.n_in >> S;
  const int m0 = blockIdx.x * FUSED_C, m1 = min(m0 + FUSED_C, n_out);
  const bool last_cta = m1 == n_out;
  // needed range of every level's input, from the final outputs back to level 0
  int lo[FUSED_MAXS + 1], hi[FUSED_MAXS + 1];
  lo[S] = m0;
  hi[S] = m1;
  for (int l = S - 1; l >= 0; l--) {
    lo[l] = 2 * lo[l + 1] - 13;
    hi[l] = 2 * hi[l + 1];
  }
  // level 0 from global memory; buffer slot i - lo[l] holds sample i of level l (i >= 0 only)
  for (int i = max(lo[0], 0) + t; i < hi[0]; i += blockDim.x) buf[0][i - lo[0]] = a.x[i];
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < S; l++) {
    const float* in = buf[cur];
    float* out = buf[cur ^ 1];
    const float* h = hs[l];
    const int lo_in = lo[l];
    auto X = [&](int i) -> float { return i >= 0 ? in[i - lo_in] : h[-i]; };
    if (last_cta && t < 16) {  // new state of this level: the last 13 samples of (history ++ input)
      const int n_l = a.n_in >> l;
      if (t >= 1 && t <= 13) {
        const int i = n_l - t;
        a.tail[16 * l + t] = i >= 0 ? in[i - lo_in] : h[t - n_l];
      }
    }
    const float4 c = a.coeff[l];
    const bool final_level = l == S - 1;
    for (int m = max(lo[l + 1], 0) + t; m < hi[l + 1]; m += blockDim.x) {
      const int b = 2 * m;
      // same association as the portable reference loop: centre, then taps from the tails inwards (decimate.c:124-128)
      float r = X(b - 6);
      r += (X(b + 1) + X(b - 13)) * c.x;
      r += (X(b - 1) + X(b - 11)) * c.y;
      r += (X(b - 3) + X(b - 9)) * c.z;
      r += (X(b - 5) + X(b - 7)) * c.w;
      if (final_level)
        a.y[m] = r;
      else
        out[m - lo[l + 1]] = r;
    }
    __syncthreads();
    cur ^= 1;
  }
}

void state_to_hist(const struct hb15_state* st, float* hist) {
  hist[0] = 0;
  hist[2] = st->even_samples[1];
  hist[4] = st->even_samples[2];
  hist[6] = st->even_samples[3];
  hist[1] = st->odd_samples[1];
  hist[3] = st->odd_samples[2];
  hist[5] = st->odd_samples[3];
  hist[7] = st->old_odd_samples[3];
  hist[9] = st->old_odd_samples[2];
  hist[11] = st->old_odd_samples[1];
  hist[13] = st->old_odd_samples[0];
  hist[8] = hist[10] = hist[12] = hist[14] = hist[15] = 0;
}

// new state from the last 13 samples of (history ++ input)
void hist_to_state(struct hb15_state* st, const float* tail /* tail[k] = x_end[-k], k=1..13 */) {
  st->even_samples[0] = tail[2];
  st->even_samples[1] = tail[2];
  st->even_samples[2] = tail[4];
  st->even_samples[3] = tail[6];
  st->odd_samples[0] = tail[1];
  st->odd_samples[1] = tail[1];
  st->odd_samples[2] = tail[3];
  st->odd_samples[3] = tail[5];
  st->old_odd_samples[3] = tail[7];
  st->old_odd_samples[2] = tail[9];
  st->old_odd_samples[1] = tail[11];
  st->old_odd_samples[0] = tail[13];
}

struct Scratch {
  float *d_a = nullptr, *d_b = nullptr, *d_hist = nullptr;
  size_t cap = 0;
  int device = -1;
};
thread_local Scratch g_scr;

int ensure_scratch(int device, size_t nfloats) {
  if (ka9q_device_count() <= device) {
    set_error("no CUDA device (no CPU fallback)");
    return -1;
  }
  cudaSetDevice(device);
  if (g_scr.device != device || g_scr.cap < nfloats) {
    if (g_scr.d_a) cudaFree(g_scr.d_a);
    if (g_scr.d_b) cudaFree(g_scr.d_b);
    if (g_scr.d_hist) cudaFree(g_scr.d_hist);
    g_scr = Scratch();
    if (cudaMalloc(&g_scr.d_a, sizeof(float) * nfloats) != cudaSuccess) return -1;
    if (cudaMalloc(&g_scr.d_b, sizeof(float) * nfloats) != cudaSuccess) return -1;
    if (cudaMalloc(&g_scr.d_hist, sizeof(float) * 16 * 64) != cudaSuccess) return -1;
    g_scr.cap = nfloats;
    g_scr.device = device;
  }
  return 0;
}

int grid_for(int n) { return (n + 255) / 256 > 2048 ? 2048 : (n + 255) / 256; }

// one stage on device buffers; updates host state from host-visible tails
int run_stage(struct hb15_state* st, const float* d_in, float* d_out, int cnt, float* d_hist, const float* h_tail_src,
              const float* h_prev_hist) {
  (void)h_tail_src;
  (void)h_prev_hist;
  float hist[16];
  state_to_hist(st, hist);
  if (cudaMemcpy(d_hist, hist, sizeof(hist), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  hb15_kernel<<<grid_for(cnt), 256>>>(d_in, d_hist, st->coeffs[0], st->coeffs[1], st->coeffs[2], st->coeffs[3], d_out, cnt);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// tail[k] = x_end[-k] for k = 1..13 where x = (old history) ++ (n new samples at d_in)
int fetch_tail(const struct hb15_state* st, const float* d_in, int n, float* tail) {
  float hist[16];
  state_to_hist(st, hist);
  float last[13];
  const int have = n < 13 ? n : 13;
  if (have > 0 && cudaMemcpy(last, d_in + (n - have), sizeof(float) * have, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  for (int k = 1; k <= 13; k++) {
    if (k <= have)
      tail[k] = last[have - k];
    else
      tail[k] = hist[k - have];
  }
  tail[0] = 0;
  return 0;
}

int pick_device() {
  const char* e = getenv("KA9Q_B200_DEVICE");
  return e ? atoi(e) : 0;
}

}  // namespace

extern "C" {

int ka9q_hb15_cascade(int device, int stages, struct hb15_state* states, const float* in, int n_in, float* out) {
  if (!states || !in || !out || stages < 1 || n_in <= 0 || (n_in % (1 << stages)) != 0) {
    set_error("ka9q_hb15_cascade: bad argument");
    return -1;
  }
  if (ensure_scratch(device, (size_t)n_in)) return -1;
  float* a = g_scr.d_a;
  float* b = g_scr.d_b;
  if (cudaMemcpy(a, in, sizeof(float) * n_in, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  if (stages <= FUSED_MAXS && !(getenv("KA9Q_B200_HB15_FUSED") && atoi(getenv("KA9Q_B200_HB15_FUSED")) == 0)) {
    // one launch, one upload of all histories, one download of all tails
    float hist[FUSED_MAXS * 16], tail[FUSED_MAXS * 16];
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    for (int l = 0; l < stages; l++) {
      const struct hb15_state* st = &states[stages - 1 - l];
      state_to_hist(st, hist + 16 * l);
      fa.coeff[l] = make_float4(st->coeffs[0], st->coeffs[1], st->coeffs[2], st->coeffs[3]);
    }
    float* d_hist = g_scr.d_hist;            // [64][16] floats: histories in the first half, tails in the second
    float* d_tail = g_scr.d_hist + 16 * 32;
    if (cudaMemcpy(d_hist, hist, sizeof(float) * 16 * stages, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    fa.x = a;
    fa.hist = d_hist;
    fa.S = stages;
    fa.n_in = n_in;
    fa.y = b;
    fa.tail = d_tail;
    const int n_out = n_in >> stages;
    hb15_cascade_kernel<<<(n_out + FUSED_C - 1) / FUSED_C, 256>>>(fa);
    if (cudaGetLastError() != cudaSuccess) return -1;
    if (cudaMemcpy(out, b, sizeof(float) * n_out, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(tail, d_tail, sizeof(float) * 16 * stages, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int l = 0; l < stages; l++) hist_to_state(&states[stages - 1 - l], tail + 16 * l);
    return 0;
  }
  int n = n_in;
  // highest-rate stage first: stage index stages-1 down to 0, each with its own state (hackrf.c:297-301)
  for (int j = stages - 1; j >= 0; j--) {
    const int cnt = n / 2;
    float tail[16];
    if (fetch_tail(&states[j], a, n, tail)) return -1;
    if (run_stage(&states[j], a, b, cnt, g_scr.d_hist + 16 * (j % 64), nullptr, nullptr)) return -1;
    hist_to_state(&states[j], tail);
    float* t = a;
    a = b;
    b = t;
    n = cnt;
  }
  if (cudaMemcpy(out, a, sizeof(float) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return 0;
}

void hb15_block(struct hb15_state* state, float* output, float* input, int cnt) {
  if (cnt <= 0) return;
  if (ka9q_hb15_cascade(pick_device(), 1, state, input, 2 * cnt, output) != 0)
    fprintf(stderr, "ka9q_b200: hb15_block failed: %s\n", ka9q_last_error());
}

void hb3_block(float* state, float* output, float* input, int cnt) {
  if (cnt <= 0) return;
  if (ensure_scratch(pick_device(), (size_t)2 * cnt)) {
    fprintf(stderr, "ka9q_b200: hb3_block failed: %s\n", ka9q_last_error());
    return;
  }
  const float prev = *state;
  const float last = input[2 * cnt - 1];  // read before the (possibly in-place) output overwrites it
  cudaMemcpy(g_scr.d_a, input, sizeof(float) * 2 * cnt, cudaMemcpyHostToDevice);
  hb3_kernel<<<grid_for(cnt), 256>>>(g_scr.d_a, prev, g_scr.d_b, cnt);
  cudaMemcpy(output, g_scr.d_b, sizeof(float) * cnt, cudaMemcpyDeviceToHost);
  *state = last;
}

}  // extern "C"
