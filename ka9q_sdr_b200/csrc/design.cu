// K4 implementation: see design.cuh.
#include <math.h>
#include "design.cuh"
#include "fft_regs.cuh"
#include "util.cuh"

namespace k9 {

// Modified Bessel function I0 by its power series, <= 40 terms, stop at term < 1e-12*sum (filter.c:282-293)
static float i0f_series(float const x) {
  const float t = 0.25 * x * x;
  float sum = 1 + t;
  float term = t;
  for (int k = 2; k < 40; k++) {
    term *= t / (k * k);
    sum += term;
    if (term < 1e-12 * sum) break;
  }
  return sum;
}

void kaiser_window_host(float* window, unsigned M, float beta) {
  // filter.c:337-357: half computed, mirrored; middle forced to 1 for odd M
  float const numc = M_PI * beta;
  float const inv_denom = 1. / i0f_series(numc);
  float const pc = 2.0 / (M - 1);
  for (unsigned n = 0; n < M / 2; n++) {
    float const p = pc * n - 1;
    window[M - 1 - n] = window[n] = i0f_series(numc * sqrtf(1 - p * p)) * inv_denom;
  }
  if (M & 1) window[(M - 1) / 2] = 1;
}

// brick-wall response over signed frequency f = n/ndec (filter.c:524-535), inclusive float compares
__global__ void brickwall_kernel(const DesignSpec* __restrict__ specs, int ndec, float2* __restrict__ out) {
  const int c = blockIdx.y;
  const DesignSpec s = specs[c];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < ndec; n += gridDim.x * blockDim.x) {
    float f;
    if (n <= ndec / 2)
      f = (float)n / ndec;
    else
      f = (float)(n - ndec) / ndec;
    const bool in = (f >= s.low && f <= s.high);
    out[(long long)c * ndec + n] = make_float2(in ? s.gain : 0.f, 0.f);
  }
}

// Shift by M/2, window, scale by 1/N, zero-pad (filter.c:386-392). The reference does this in place for
// n = M-1 down to 0; that equals the out-of-place form whenever N - M/2 >= M. Otherwise a few low-n entries
// read values already rewritten in the same loop; `resolve` follows that chain so the result is identical.
__device__ __forceinline__ float2 shifted_value(const float2* __restrict__ h, const float* __restrict__ w, int n, int N,
                                                int M, float gain, bool real_only) {
  // chain of at most a few hops: src > idx and src < M means the in-place loop had already rewritten it
  int chain[8];
  int depth = 0;
  int idx = n;
  int src;
  while (true) {
    chain[depth++] = idx;
    src = (idx - M / 2 + N) % N;
    if (src > idx && src < M && depth < 8)
      idx = src;
    else
      break;
  }
  float2 v = h[src];
  if (real_only) v.y = 0.f;
  for (int d = depth - 1; d >= 0; d--) {
    const float ww = w[chain[d]];
    v = make_float2(v.x * ww * gain, v.y * ww * gain);
  }
  return v;
}

__global__ void shift_window_kernel(const float2* __restrict__ h, const DesignSpec* __restrict__ specs,
                                    const float* __restrict__ windows, int fixed_window, int N, int M, bool real_only,
                                    float2* __restrict__ out) {
  const int c = blockIdx.y;
  const int wi = specs ? specs[c].window : fixed_window;
  const float* w = windows + (long long)wi * M;
  const float gain = 1. / N;
  const float fine = specs ? specs[c].fine : 0.f;
  const float2* hc = h + (long long)c * N;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (n < M) v = shifted_value(hc, w, n, N, M, gain, real_only);
    if (fine != 0.f && n < M) {  // h[n] * exp(+j 2 pi fine n): the filter an off-grid LO is equivalent to (stream.cu)
      float sn, cs;
      sincospif(2.f * fine * (float)n, &sn, &cs);
      v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    }
    out[(long long)c * N + n] = v;
  }
}

// Hermitian extension of a half spectrum (what c2r assumes: filter.c:436-437): X[N-k] = conj(X[k]);
// imaginary parts of DC and (even N) Nyquist ignored.
__global__ void hermitian_extend_kernel(const float2* __restrict__ half, int N, float2* __restrict__ full) {
  const int c = blockIdx.y;
  const int nh = N / 2 + 1;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float2 v;
    if (n < nh) {
      v = half[(long long)c * nh + n];
      if (n == 0 || (2 * n == N)) v.y = 0.f;
    } else {
      v = half[(long long)c * nh + (N - n)];
      v.y = -v.y;
    }
    full[(long long)c * N + n] = v;
  }
}

__global__ void noise_gain_kernel(const float2* __restrict__ resp, int stride, int bins, float scale,
                                  float* __restrict__ out) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int n = threadIdx.x; n < bins; n += blockDim.x) {
    const float2 v = resp[(long long)c * stride + n];
    s += v.x * v.x + v.y * v.y;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[c] = scale * v;
  }
}

static dim3 grid2(int n, int count) { return dim3((n + 255) / 256 > 64 ? 64 : (n + 255) / 256, count); }

int design_complex_batch(const BigFftPlan* plan, int ndec, int mdec, const DesignSpec* d_specs, int count,
                         const float* d_windows, float2* d_resp, float2* d_work, cudaStream_t st) {
  if (count <= 0) return 0;
  float2* w0 = d_work;
  float2* w1 = d_work + (long long)count * ndec;
  brickwall_kernel<<<grid2(ndec, count), 256, 0, st>>>(d_specs, ndec, d_resp);
  BigFftIn in;
  in.in = d_resp;
  in.in_batch_stride = ndec;
  // impulse response: unnormalised backward transform (filter.c:376-377)
  if (bigfft_exec(plan, in, w0, ndec, w1, d_resp /*unused 3rd buffer when npass<3*/, count, +1, st)) return -1;
  shift_window_kernel<<<grid2(ndec, count), 256, 0, st>>>(w0, d_specs, d_windows, 0, ndec, mdec, false, w1);
  in.in = w1;
  if (bigfft_exec(plan, in, d_resp, ndec, w0, nullptr, count, -1, st)) return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int window_filter_device(const BigFftPlan* plan, int M, float2* d_resp, int count, const float* d_window, float2* d_work,
                         cudaStream_t st) {
  const int N = plan->N;
  float2* w0 = d_work;
  float2* w1 = d_work + (long long)count * N;
  BigFftIn in;
  in.in = d_resp;
  in.in_batch_stride = N;
  if (plan->npass >= 3) return -2;  // scratch layout below assumes <= 2 passes (N <= 102400)
  if (bigfft_exec(plan, in, w0, N, w1, nullptr, count, +1, st)) return -1;
  shift_window_kernel<<<grid2(N, count), 256, 0, st>>>(w0, nullptr, d_window, 0, N, M, false, w1);
  in.in = w1;
  if (bigfft_exec(plan, in, d_resp, N, w0, nullptr, count, -1, st)) return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int window_rfilter_device(const BigFftPlan* plan, int M, const float2* d_half, float2* d_full, int count,
                          const float* d_window, float2* d_work, cudaStream_t st) {
  const int N = plan->N;
  float2* w0 = d_work;
  float2* w1 = d_work + (long long)count * N;
  if (plan->npass >= 3) return -2;
  hermitian_extend_kernel<<<grid2(N, count), 256, 0, st>>>(d_half, N, d_full);
  BigFftIn in;
  in.in = d_full;
  in.in_batch_stride = N;
  if (bigfft_exec(plan, in, w0, N, w1, nullptr, count, +1, st)) return -1;
  // time domain is real (c2r output): drop the rounding-level imaginary part
  shift_window_kernel<<<grid2(N, count), 256, 0, st>>>(w0, nullptr, d_window, 0, N, M, true, w1);
  in.in = w1;
  if (bigfft_exec(plan, in, d_full, N, w0, nullptr, count, -1, st)) return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int noise_gain_device(const float2* d_resp, int stride, int bins, int count, float scale, float* d_out, cudaStream_t st) {
  if (count <= 0) return 0;
  noise_gain_kernel<<<count, 256, 0, st>>>(d_resp, stride, bins, scale, d_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace k9
