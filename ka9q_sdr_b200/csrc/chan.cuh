// K3: per-channel fused kernels — shared declarations (device structs + host launchers).
//
// One work item = what one reference `radio` process does per 20 ms block after its forward FFT:
// execute_filter_output (reference filter.c:175-252) -> demod_fm / demod_am / demod_linear per-sample loops
// (fm.c:72-173, am.c:43-79, linear.c:114-311) -> scaleclip (audio.c:22-28).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace k9 {

constexpr int NDEC = 2048;    // decimated FFT size the fused kernels are specialised for
constexpr int OLEN_MAX = 1024;

enum DemodType : int { DEMOD_LINEAR = 0, DEMOD_AM = 1, DEMOD_FM = 2 };  // reference radio.h:20-24
enum ChanFlags : int { CH_ISB = 1, CH_FLAT = 2, CH_PLL = 4, CH_SQUARE = 8 };

struct ChanParams {
  long long bin;         // carrier position on the N-point grid, normalised to [0,N); second LO = -bin*Fs/N
  int demod;
  int flags;
  int channels;          // PCM channels (1 mono, 2 stereo I/Q)
  int pcm_off;           // int16 offset of this channel inside one block row of the PCM output
  float fm_gain;         // (headroom * M_1_PI * dsamprate) / fabsf(low - high)   (fm.c:86)
  float headroom;        // demod->agc.headroom (main.c:117)
  float recovery_factor; // dB2voltage(recovery_rate * samptime)  (am.c:27, linear.c:33)
  int hangmax;           // hangtime / samptime (am.c:29, linear.c:37)
  double shift_cycles;   // post-detection shift, cycles per output sample (radio.c:313)
  int audio_slot;        // FM: index of the audio (de-emphasis) response, -1 = flat
  int phase_step;        // (bin * L) mod N: per-block advance of the LO phase index
  int resp_slot;         // row of the response table this channel reads: channels with identical filters share one row
};

struct ChanState {
  float2 fm_state;       // conj(last good sample) (fm.c:26,132)
  float fm_lastaudio;    // fm.c:68
  int fm_below;          // snr_below_threshold (fm.c:69)
  float fm_foffset, fm_pdeviation;  // persist while squelch closed (fm.c:145-154)
  float agc_gain;        // demod->agc.gain
  float am_dc;           // DC_filter (am.c:33)
  int hang;              // hangcount
  int pad_;
  double shift_phase;    // turns, phase of the post-detection shift oscillator at the start of the next block
};

// Coherent (pll / square) linear channels: constants of linear.c:29-65 evaluated on the host in the reference's types
struct PllParams {
  float samptime, blocktime;        // linear.c:29-30
  float snrthresh;                  // linear.c:49
  int lock_limit;                   // linear.c:50
  float binsize;                    // linear.c:51
  int lowlimit, highlimit;          // linear.c:55-56
  float integrator_gain, prop_gain; // linear.c:63,65
};
struct PllState {
  double2 coarse_ph, fine_ph;       // NCO phasors (struct osc, osc.h:9-17)
  double coarse_freq, fine_freq;    // cycles per sample
  float integrator, delta_f;        // linear.c:107-108
  int lock_count, pll_lock;         // linear.c:110, demod->sig.pll_lock
  int fft_samples, fft_ptr;         // linear.c:112,87
  float snr, foffset, cphase;       // demod->sig.*
  int pad_;
};

// PL-tone analyser (fm.c:189-285): one work item per de-emphasised FM channel
struct PlWork {
  int chan, pair, half;  // channel index, index of its pair in the FM work list, 0 = the pair's first channel (real part)
  int pad_;
};
struct PlState {
  int fft_ptr, last_fft;  // fm.c:231-232
  float plfreq;           // demod->sig.plfreq
  int pad_;
};

struct ChanStatus {      // mirrors the demod->sig.* scalars the reference demodulators publish
  float bb_power;        // fm.c:99, am.c:78, linear.c:302
  float snr;             // fm.c:102-103 (NAN for AM / non-PLL linear, linear.c:309)
  float foffset;         // fm.c:147
  float pdeviation;      // fm.c:152
  float agc_gain;        // demod->agc.gain after the block
  int squelch_open;      // FM: snr_below_threshold < 2
  float reserved[2];
};

struct ChanLaunch {
  // shared per-stream data
  const float2* spec;        // [nblocks][N] forward spectra
  long long spec_stride;     // N
  int N;                     // forward FFT size
  int L, M;                  // block length / impulse length at the input rate (for the per-block LO phase)
  int olen;                  // output samples per block (960)
  float dsamprate;           // decimated (output) sample rate, (float)samprate / decimate (fm.c:27)
  int nblocks;
  long long block0;          // index of the first block of this launch since stream start
  int start0;                // (block0*L - (M-1)) mod N: stream index of the first window sample, mod N
  const float2* twN_lo;      // forward-FFT twiddle tables: W_N^a = lo[a & 1023] * hi[a >> 10]
  const float2* twN_hi;
  const float2* tw2048;      // W_2048 table
  // per-channel arrays
  const ChanParams* params;
  ChanState* state;
  const float2* resp;        // [nchan][2048] channel responses H (filter_out.response)
  const float2* audio_resp;  // [naudio][2048] full-length (Hermitian-extended) FM audio responses
  float* audio_hist;         // [nchan][2048] FM audio ring (history M_audio-1 = 1088 samples + 960 new)
  int16_t* pcm;              // [nblocks][pcm_stride]
  long long pcm_stride;
  ChanStatus* status;        // [nblocks][nchan_total]
  int nchan_total;
  float2* filt_dbg;          // optional [nblocks][nchan_total][olen] raw filter output capture (parity tests), or null
  // work list for this launch
  const int2* work;          // FM: (chanA, chanB or -1); AM / linear: (chan, -1)
  int nwork;
  // FM block-split form (chan_kernels.cu: fm_split_wait): per-pair count of discriminated blocks; CTAs per pair
  long long* fm_seq;         // [nwork] or null (then never split)
  int fm_split;              // set by launch_fm
  // AM / linear scratch between the front, recurrence and output kernels (rows indexed by block*nwork + work index)
  float* agc_x;              // [nblocks*nwork][olen]: amplitude in; AM: (s - DC)*gain out, linear: gain out
  float2* agc_y;             // [nblocks*nwork][olen]: kept filter output (linear only)
  float* agc_pow;            // [nblocks*nwork][2]: block sums (am.c:56-58 / linear.c:256-261)
  // coherent linear channels (work = (chan, pll slot))
  const PllParams* pll_params;
  PllState* pll_state;
  float2* pll_ring;          // [npll][65536] carrier-search ring (linear.c:84-93)
  // PL-tone analyser (optional)
  float2* pl_spec;           // [nblocks][pl_npairs][65]: low bins of every pair's audio transform, written by fm_kernel
  int pl_npairs;
  const PlWork* pl_work;
  PlState* pl_state;
  float* pl_ring;            // [npl][16384]
  const float2* pl_resp;     // [33] slave response (fm.c:208-218)
};

// mixed: AM / linear kernels of the same stream run beside this one (they need the maximum shared-memory carve-out; CTAs
// of kernels with different carve-outs cannot share an SM, so the FM kernel then gives up its larger L1)
int launch_fm(const ChanLaunch& a, cudaStream_t st, bool mixed);
// the carve-out (percent or cudaSharedmemCarveoutMaxShared) launch_fm asks for: kernels meant to run beside it use the same
int fm_carveout(bool mixed);
int launch_am(const ChanLaunch& a, cudaStream_t st);
int launch_linear(const ChanLaunch& a, cudaStream_t st);
int launch_pll(const ChanLaunch& a, cudaStream_t st);
int launch_pl(const ChanLaunch& a, int nchan_pl, cudaStream_t st);

}  // namespace k9
