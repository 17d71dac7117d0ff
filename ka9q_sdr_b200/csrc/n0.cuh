// K6: compute_n0 (reference radio.c:383-425) once per stream — declarations. See n0.cu.
#pragma once
#include <cuda_runtime.h>

namespace k9 {

constexpr int N0_POWER_CTAS = 592;  // 148 SMs x 4: CTAs of the two passes over the power spectrum
constexpr int N0_LIST_CAP = 1 << 16;

struct N0Chan {
  int bin;        // carrier bin in [0, N)
  int nlo, nhi;   // passband as signed bin offsets from the carrier (the bins compute_n0 skips)
  float alpha;    // smoothing constant of demod->sig.n0: 0.01 FM (fm.c:82), 0.001 AM / linear (am.c:47, linear.c:124)
};
struct N0Block {
  double total;                 // sum of P over all N bins
  double base_sum;              // sum over the bins below the smallest threshold
  unsigned long long base_cnt;
  unsigned tmin_bits, tmax_bits;  // smallest / largest channel threshold of the block (float bits; positive floats order like uints)
  int list_count;
  int pad;
};
struct N0Launch {
  const float2* spec;  // [nblocks][N] forward spectra
  long long spec_stride;
  int N, samprate, nblocks, nchan;
  const N0Chan* chan;
  float* P;            // [nblocks][N] power spectrum scratch
  double* partial;     // [nblocks][N0_POWER_CTAS]
  N0Block* blk;        // [nblocks]
  float* T;            // [nblocks][nchan] per-channel thresholds
  float* list;         // [nblocks][list_cap]
  int list_cap;
  float* n0_raw;       // [nblocks][nchan] compute_n0() of the block
  float* n0_smooth;    // [nblocks][nchan] demod->sig.n0 after the block
  float* state;        // [nchan] demod->sig.n0 carried between launches
};

int n0_launch(const N0Launch& a, cudaStream_t st);
void n0_passband_bins(int N, int samprate, float low, float high, int* nlo, int* nhi);

}  // namespace k9
