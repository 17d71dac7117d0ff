// K6: noise-density estimate, compute_n0 (reference radio.c:383-425), once per stream instead of once per channel.
//
// The reference calls compute_n0 per channel and block (fm.c:78, am.c:46-49, linear.c:123-126): the power spectrum of ALL N
// master bins, a first average over the bins outside the channel's passband, then a second average that drops every bin
// at or above twice (+3 dB) the first — O(N) per channel, N = 2.6 M at cfg5. A channel's master spectrum is the shared
// spectrum rotated by its carrier bin, |.|^2 is phase blind, so all channels look at the SAME power
// spectrum P[n] = |X[n]|^2 and differ only in (a) which window of bins is the passband and (b) the threshold
// T_c = 2 * avg1_c, which is nearly the same number for every channel. Per block:
//
//   n0_power_kernel     P = |X|^2 (N floats) and its total (per-CTA double partials, summed in a fixed order)
//   n0_chan_a_kernel    per channel (one warp): passband sum from P -> avg1_c, T_c; block-wide min / max of T_c
//   n0_global_kernel    one pass over P: sum / count of the bins below T_min, and the (few) bins in [T_min, T_max) compacted
//                       into a list
//   n0_chan_b_kernel    per channel (one warp): + list entries below T_c, - passband bins below T_c -> second average,
//                       / (2 N Fs) (radio.c:424), then the reference's smoothing (fm.c:79-82: 0.01; am.c:46-49,
//                       linear.c:123-126: 0.001; demod->sig.n0 starts at 0 in the zero-initialised struct demod)
//
// O(N + K * passband) per block. Sums are carried in double (the reference accumulates 2.6 M floats sequentially in a float,
// ~1e-4 relative rounding at cfg5); the parity bar is 1e-3 relative on the smoothed value (tests/test_gpu_parity_configs.py).
#include <math.h>
#include "n0.cuh"
#include "util.cuh"

namespace k9 {

constexpr int N0_THREADS = 256;

__global__ void __launch_bounds__(N0_THREADS) n0_power_kernel(const float2* __restrict__ X, long long spec_stride, int N,
                                                               float* __restrict__ P, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const float4* x4 = reinterpret_cast<const float4*>(X + (long long)b * spec_stride);
  float2* p2 = reinterpret_cast<float2*>(P + (long long)b * N);
  double acc = 0.0;
  for (int i = blockIdx.x * N0_THREADS + threadIdx.x; i < N / 2; i += gridDim.x * N0_THREADS) {
    const float4 v = __ldg(x4 + i);
    const float a = v.x * v.x + v.y * v.y, c = v.z * v.z + v.w * v.w;  // cnrmf (radio.c:396)
    p2[i] = make_float2(a, c);
    acc += (double)a + (double)c;
  }
  __shared__ double red[N0_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < N0_THREADS / 32; w++) s += red[w];
    partial[(long long)b * gridDim.x + blockIdx.x] = s;
  }
}

// sums the per-CTA partials of a block in a fixed order; also resets the block's threshold bounds and list counter
__global__ void n0_finish_total_kernel(const double* __restrict__ partial, int nparts, N0Block* __restrict__ blk) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; i++) s += partial[(long long)b * nparts + i];
    blk[b].total = s;
    blk[b].tmin_bits = 0x7f800000u;  // +inf
    blk[b].tmax_bits = 0u;
    blk[b].list_count = 0;
    blk[b].base_sum = 0.0;
    blk[b].base_cnt = 0ull;
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per (channel, block): first average (radio.c:401-421, iteration 0: every bin outside the passband counts)
__global__ void n0_chan_a_kernel(const float* __restrict__ P, int N, const N0Chan* __restrict__ ch, int nchan,
                                 N0Block* __restrict__ blk, float* __restrict__ T) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  if (c >= nchan) return;
  const N0Chan q = ch[c];
  const float* Pb = P + (long long)b * N;
  double s = 0.0;
  for (int n = q.nlo + lane; n <= q.nhi; n += 32) {
    int i = q.bin + n;
    i = i < 0 ? i + N : (i >= N ? i - N : i);
    s += (double)Pb[i];
  }
  s = warp_sum_d(s);
  if (lane == 0) {
    const int npass = q.nhi >= q.nlo ? q.nhi - q.nlo + 1 : 0;
    const float avg1 = (float)(blk[b].total - s) / (float)(N - npass);  // new_avg_n /= noisebins (radio.c:420)
    const float t = avg1 * 2;                                          // +3 dB threshold (radio.c:415)
    T[(long long)b * nchan + c] = t;
    atomicMin(&blk[b].tmin_bits, __float_as_uint(fmaxf(t, 0.f)));
    atomicMax(&blk[b].tmax_bits, __float_as_uint(fmaxf(t, 0.f)));
  }
}

// one pass over the power spectrum: everything below the smallest threshold is summed once for all channels; the bins
// between the smallest and the largest threshold go to a list the per-channel pass decides on
__global__ void __launch_bounds__(N0_THREADS) n0_global_kernel(const float* __restrict__ P, int N, N0Block* __restrict__ blk,
                                                                float* __restrict__ list, int list_cap) {
  const int b = blockIdx.y;
  const float* Pb = P + (long long)b * N;
  const float tmin = __uint_as_float(blk[b].tmin_bits), tmax = __uint_as_float(blk[b].tmax_bits);
  double acc = 0.0;
  unsigned long long cnt = 0;
  for (int i = blockIdx.x * N0_THREADS + threadIdx.x; i < N; i += gridDim.x * N0_THREADS) {
    const float v = Pb[i];
    if (v < tmin) {
      acc += (double)v;
      cnt++;
    } else if (v < tmax) {
      const int slot = atomicAdd(&blk[b].list_count, 1);
      if (slot < list_cap) list[(long long)b * list_cap + slot] = v;
    }
  }
  acc = warp_sum_d(acc);
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&blk[b].base_sum, acc);  // double atomics: the order of the additions moves the sum by ~1e-16 relative
    atomicAdd(&blk[b].base_cnt, cnt);
  }
}

// one warp per channel, blocks in stream order (the smoothing is a recurrence): second average and n0
__global__ void n0_chan_b_kernel(const float* __restrict__ P, int N, const N0Chan* __restrict__ ch, int nchan, int nblocks,
                                 const N0Block* __restrict__ blk, const float* __restrict__ T, const float* __restrict__ list,
                                 int list_cap, double scale /* 1 / (2 N Fs) */, float* __restrict__ n0_raw,
                                 float* __restrict__ n0_smooth, float* __restrict__ state) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= nchan) return;
  const N0Chan q = ch[c];
  float sm = state[c];
  for (int b = 0; b < nblocks; b++) {
    const float* Pb = P + (long long)b * N;
    const float t = T[(long long)b * nchan + c];
    double s = 0.0;
    long long k = 0;
    const int nl = min(blk[b].list_count, list_cap);
    for (int i = lane; i < nl; i += 32) {
      const float v = list[(long long)b * list_cap + i];
      if (v < t) {
        s += (double)v;
        k++;
      }
    }
    for (int n = q.nlo + lane; n <= q.nhi; n += 32) {  // the passband is skipped (radio.c:411-412)
      int i = q.bin + n;
      i = i < 0 ? i + N : (i >= N ? i - N : i);
      const float v = Pb[i];
      if (v < t) {
        s -= (double)v;
        k--;
      }
    }
    s = warp_sum_d(s);
    for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    if (lane == 0) {
      const double sum = blk[b].base_sum + s;
      const long long cnt = (long long)blk[b].base_cnt + k;
      const float avg2 = (float)(sum / (double)cnt);  // new_avg_n /= noisebins (0/0 = NaN as in the reference)
      const float n0 = (float)((double)avg2 * scale);  // radio.c:424
      // demod->sig.n0 += a * (n0 - demod->sig.n0), a = .01 (fm.c:82) or .001 (am.c:47, linear.c:124) as a double constant
      sm = isnan(sm) ? n0 : (float)((double)sm + (double)q.alpha * (double)(n0 - sm));
      n0_raw[(long long)b * nchan + c] = n0;
      n0_smooth[(long long)b * nchan + c] = sm;
    }
    sm = __shfl_sync(0xffffffffu, sm, 0);
  }
  if (lane == 0) state[c] = sm;
}

int n0_launch(const N0Launch& a, cudaStream_t st) {
  if (a.nchan <= 0 || a.nblocks <= 0) return 0;
  const int nparts = N0_POWER_CTAS;
  n0_power_kernel<<<dim3(nparts, a.nblocks), N0_THREADS, 0, st>>>(a.spec, a.spec_stride, a.N, a.P, a.partial);
  n0_finish_total_kernel<<<a.nblocks, 32, 0, st>>>(a.partial, nparts, a.blk);
  const int wpb = 8;  // warps (channels) per CTA
  n0_chan_a_kernel<<<dim3((a.nchan + wpb - 1) / wpb, a.nblocks), 32 * wpb, 0, st>>>(a.P, a.N, a.chan, a.nchan, a.blk, a.T);
  n0_global_kernel<<<dim3(nparts, a.nblocks), N0_THREADS, 0, st>>>(a.P, a.N, a.blk, a.list, a.list_cap);
  n0_chan_b_kernel<<<(a.nchan + wpb - 1) / wpb, 32 * wpb, 0, st>>>(a.P, a.N, a.chan, a.nchan, a.nblocks, a.blk, a.T, a.list,
                                                                  a.list_cap, 1.0 / (2.0 * (double)a.N * (double)a.samprate),
                                                                  a.n0_raw, a.n0_smooth, a.state);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Passband of a channel as signed master-bin offsets [nlo, nhi] from its carrier: the bins compute_n0 skips,
// low <= f <= high with f = (float)(n * samprate) / N evaluated in float exactly as radio.c:406-412 does (the product in
// 64 bits: the reference's int product overflows above ~2^31 / samprate bins, i.e. at rates it was never run at).
void n0_passband_bins(int N, int samprate, float low, float high, int* nlo, int* nhi) {
  auto f_of = [&](long long n) { return (float)(n * (long long)samprate) / N; };
  long long a = (long long)floor((double)low * N / samprate) - 2, e = (long long)ceil((double)high * N / samprate) + 2;
  if (a < -(long long)(N - 1) / 2) a = -(long long)(N - 1) / 2;  // signed grid: n in (-N/2, N/2]
  if (e > N / 2) e = N / 2;
  while (a <= e && !(f_of(a) >= low)) a++;
  while (e >= a && !(f_of(e) <= high)) e--;
  *nlo = (int)a;
  *nhi = (int)e;
}

}  // namespace k9
