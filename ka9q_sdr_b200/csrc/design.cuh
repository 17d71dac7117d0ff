// K4: filter design on the device — set_filter / window_filter / window_rfilter / make_kaiser / noise_gain
// (reference filter.c:500-546, :365-415, :420-469, :337-357, :472-497), batched over channels.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "bigfft.cuh"

namespace k9 {

// Host: Kaiser window exactly as filter.c:337-357 (+ i0, filter.c:282-293), fp32 arithmetic.
void kaiser_window_host(float* w, unsigned M, float beta);

struct DesignSpec {
  float low, high;   // cycles/sample at the output rate, as passed to set_filter (filter.c:500)
  float gain;        // 1/N, times M_SQRT1_2 for REAL / CROSS_CONJ outputs (filter.c:518-522)
  int window;        // index into the window table
  float fine;        // off-grid part of the carrier, cycles per output sample (bins / ndec): phase ramp on the impulse response
};

// Batched set_filter core for complex responses of length ndec (plan must be for ndec):
//   resp[c][n] = FFT-( shift/window/scale( FFT+( brickwall_c ) ) )[n]
// windows: device float [nwindows][mdec]. work: device scratch float2[2][count][ndec].
int design_complex_batch(const BigFftPlan* plan, int ndec, int mdec, const DesignSpec* d_specs, int count,
                         const float* d_windows, float2* d_resp, float2* d_work, cudaStream_t st);

// window_filter (filter.c:365) on a device buffer in place: resp[count][n], n = L+M-1 = plan->N.
int window_filter_device(const BigFftPlan* plan, int M, float2* d_resp, int count, const float* d_window, float2* d_work,
                         cudaStream_t st);
// window_rfilter (filter.c:420): input d_half[count][n/2+1] (Hermitian half), output FULL-length Hermitian
// spectrum d_full[count][n] (bins 0..n/2 are the reference's result; the rest is its conjugate mirror).
int window_rfilter_device(const BigFftPlan* plan, int M, const float2* d_half, float2* d_full, int count,
                          const float* d_window, float2* d_work, cudaStream_t st);

// noise_gain (filter.c:472-497): out[c] = scale * sum_n |resp[c][n]|^2 over `bins` bins
int noise_gain_device(const float2* d_resp, int stride, int bins, int count, float scale, float* d_out, cudaStream_t st);

}  // namespace k9
