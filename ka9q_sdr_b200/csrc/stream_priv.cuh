// Private to the batch layer: the stream object and the issue helpers shared by stream.cu and mgpu.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>
#include <vector>
#include "../../include/ka9q_b200.h"
#include "bigfft.cuh"
#include "chan.cuh"
#include "n0.cuh"

#include "util.cuh"
using namespace k9;  // private header of two translation units

constexpr int K9_MAX_RANKS = 16;

// Channel-sharded multi-GPU state (mgpu.cu)
constexpr int K9_MAX_SPEC = 3;           // spectrum buffers per stream (ka9q_stream::nspec of them in use)
struct MgpuFlags {                       // lives in device memory of every rank, written by its peers over NVLink
  int ready[K9_MAX_SPEC][K9_MAX_RANKS];  // [spectrum buffer][producer]: last batch whose sub-bands from that producer landed
  int freed[K9_MAX_SPEC][K9_MAX_RANKS];  // [spectrum buffer][consumer]: last batch that consumer finished reading
  int error;                             // a wait timed out
  int pad[31];
};
struct CopyJob {                         // one arc of one block for one peer, in 16-byte units
  const int4* src;
  int4* dst;
  long long n16;
  long long first;                       // index of this job's first unit in the concatenation of all jobs
};
struct MgpuSeg {                         // one contiguous piece of a rank's needed arc, in bins
  long long lo, len;
};

struct ka9q_stream {
  ka9q_stream_config cfg;
  int N = 0, olen = 0, mdec = 0;
  int bytes_per_samp = 4;
  int carveout = 0;            // shared-memory carve-out every kernel of this stream asks for (so they can share SMs)
  bool committed = false;
  BigFftPlan fwd, p2048;
  // device
  void* d_ring = nullptr;
  long long ring_cap = 0;
  long long pushed = 0;        // samples pushed since stream start
  long long block0 = 0;        // blocks computed since stream start
  long long phase_block = 0;   // block counter used for LO phase / audio ring (advances in resident mode too)
  float2 *d_spec = nullptr, *d_tmp0 = nullptr, *d_tmp1 = nullptr;
  float* d_energy = nullptr;
  float2* d_tw2048 = nullptr;
  std::vector<ka9q_chan_params> chans;
  std::vector<double> fine_bins;  // off-grid part of each channel's carrier, in bins of the N-point grid, |.| <= 0.5
  std::vector<ChanParams> h_params;
  std::vector<float> h_noise_gain;
  ChanParams* d_params = nullptr;
  ChanState* d_state = nullptr;
  float2* d_resp = nullptr;
  float2* d_audio_resp = nullptr;
  float* d_audio_hist = nullptr;
  // AM / linear scratch between the front, recurrence and output kernels: [max_blocks][n][olen]
  float* d_agc_x_am = nullptr;
  float* d_agc_x_lin = nullptr;
  float2* d_agc_y_lin = nullptr;
  float* d_agc_pow = nullptr;  // [max_blocks][n_am + n_lin][2]
  int16_t* d_pcm = nullptr;
  ChanStatus* d_status = nullptr;
  float2* d_filt = nullptr;
  float* d_windows = nullptr;
  std::vector<float> betas;  // distinct Kaiser betas -> window table rows
  long long* d_fm_seq = nullptr;  // FM block-split form: discriminated blocks per pair
  int2 *d_work_fm = nullptr, *d_work_am = nullptr, *d_work_lin = nullptr, *d_work_pll = nullptr;
  int n_fm = 0, n_am = 0, n_lin = 0, n_pll = 0;
  PllParams* d_pll_params = nullptr;
  PllState* d_pll_state = nullptr;
  float2* d_pll_ring = nullptr;
  cudaStream_t s_pll = nullptr;
  cudaEvent_t e_pll = nullptr;
  long long pcm_stride = 0;
  // pinned staging
  void* h_iq = nullptr;
  int16_t* h_pcm = nullptr;
  ChanStatus* h_status = nullptr;
  // streams / events
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr, s_fm = nullptr, s_am = nullptr, s_lin = nullptr;
  cudaEvent_t e_pushed = nullptr, e_fft0 = nullptr, e_fft1 = nullptr, e_chan1 = nullptr, e_fork = nullptr, e_am = nullptr,
              e_lin = nullptr, e_fm = nullptr, e_comp_done[2] = {nullptr, nullptr}, e_fetched[2] = {nullptr, nullptr};
  int comp_parity = 0;
  int last_nblocks = 0;
  cudaStream_t s_fft = nullptr;
  cudaStream_t s_mgx = nullptr, pub_stream = nullptr;  // multi-GPU: exchange stream; stream the next publish_spectrum records on
  cudaEvent_t e_mg_fft = nullptr;
  cudaStream_t s_mgwait = nullptr, s_mgsig = nullptr;  // multi-GPU: flag waits / flag signals, off the channel stream's critical path
  cudaEvent_t e_mg_ready = nullptr, e_mg_chan = nullptr;
  bool timing_regions = true;  // false: timer_start_plain, only the outer event pair is recorded
  cudaEvent_t e_spec_ready[K9_MAX_SPEC] = {}, e_spec_free[K9_MAX_SPEC] = {};
  int nspec = 3;  // spectrum buffers: the forward FFT (+ exchange) may run up to nspec - 1 batches ahead of the channel kernels
  int spec_wr = 0, spec_rd = 0, spec_published = 0;
  bool fft_pending = false;
  int fft_blocks_per_launch = 0;  // 0 = all blocks of the batch in one launch per pass (measured faster than per-block)
  bool overlap = true;  // false: the forward FFT waits for the previous batch's channel kernels (per-kernel timing)
  // PL-tone analyser (pl_kernel.cuh): enabled by ka9q_stream_enable_pl before commit
  bool pl_enabled = false;
  int n_pl = 0;
  float2 *d_pl_spec = nullptr, *d_pl_resp = nullptr;
  PlWork* d_pl_work = nullptr;
  PlState* d_pl_state = nullptr;
  float* d_pl_ring = nullptr;
  // K6 noise density (n0.cu): enabled by ka9q_stream_enable_n0 before commit
  bool n0_enabled = false;
  cudaStream_t s_n0 = nullptr;
  cudaEvent_t e_n0 = nullptr;
  N0Chan* d_n0_chan = nullptr;
  float *d_n0_P = nullptr, *d_n0_T = nullptr, *d_n0_list = nullptr, *d_n0_raw = nullptr, *d_n0_smooth = nullptr,
        *d_n0_state = nullptr;
  double* d_n0_partial = nullptr;
  N0Block* d_n0_blk = nullptr;
  // NCCL (dlopen'ed)
  void* nccl_comm = nullptr;
  int nccl_rank = 0, nccl_nranks = 1;
  // channel-sharded multi-GPU exchange (mgpu.cu)
  int mg_rank = 0, mg_nranks = 0, mg_transport = 0;  // transport: 1 NCCL send/recv, 2 peer-memory stores (NVLink P2P)
  int mg_seq = 0;                                    // batches exchanged so far
  std::vector<std::vector<MgpuSeg>> mg_need;         // per rank: the arc of the spectrum its channels read (1-2 pieces)
  MgpuFlags* d_flags = nullptr;                      // own flags (peers write into them)
  float2* mg_peer_spec[K9_MAX_RANKS] = {};           // peer spectrum buffers mapped into this process (IPC), [rank]
  MgpuFlags* mg_peer_flags[K9_MAX_RANKS] = {};
  bool mg_ipc_opened[K9_MAX_RANKS] = {};
  void* d_mg_jobs = nullptr;                         // copy-job list of the scatter kernel
  unsigned* d_mg_counter = nullptr;                  // CTAs of the scatter kernel that have finished
  int mg_njobs = 0, mg_jobs_nblocks = 0;
  std::vector<CopyJob> mg_host_jobs[K9_MAX_SPEC];              // host copy of the job lists (copy-engine transport)
  bool mg_route_now = false;                         // set by mgpu_compute around its issue_fft
  bool mg_pull = false;                              // consumers load their arcs from the producers' buffers (default)
  bool mg_fused = false;                             // the forward FFT's last pass stores the arcs itself (no copy kernel)
  unsigned short* d_mg_mask = nullptr;               // [N / 16] ranks that read each 16-bin row
  long long mg_delta[K9_MAX_RANKS] = {};             // byte offset from this rank's spectrum allocation to each peer's
  int mg_wait_ready = 0;                             // sequence number the next channel launch has to wait for (P2P)
  MgpuFlags** d_mg_peer_flag_ptrs = nullptr;         // device copy of mg_peer_flags
  // live timing of the timed region (bench.py): event pairs around every forward FFT and every FM/AM/linear launch
  bool timing = false;
  cudaEvent_t e_t0 = nullptr, e_t1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_used;  // (class, (start, stop))
  size_t ev_next = 0;
};

// NCCL entry points (libnccl is dlopen'ed on first use: load_nccl in stream.cu)
typedef struct ncclComm* k9_ncclComm_t;
typedef struct {
  char internal[128];
} k9_ncclUniqueId;
extern "C" {
extern int (*p_ncclGetUniqueId)(k9_ncclUniqueId*);
extern int (*p_ncclCommInitRank)(k9_ncclComm_t*, int, k9_ncclUniqueId, int);
extern int (*p_ncclBroadcast)(const void*, void*, size_t, int, int, k9_ncclComm_t, cudaStream_t);
extern int (*p_ncclCommDestroy)(k9_ncclComm_t);
extern int (*p_ncclAllGather)(const void*, void*, size_t, int, k9_ncclComm_t, cudaStream_t);
extern int (*p_ncclSend)(const void*, size_t, int, int, k9_ncclComm_t, cudaStream_t);
extern int (*p_ncclRecv)(void*, size_t, int, int, k9_ncclComm_t, cudaStream_t);
extern int (*p_ncclGroupStart)(void);
extern int (*p_ncclGroupEnd)(void);
extern const char* (*p_ncclGetErrorString)(int);
int load_nccl();
}

enum TimeClass { TC_FFT = 0, TC_FM = 1, TC_AM = 2, TC_LIN = 3, TC_BCAST = 4, TC_COUNT = 5, TC_WAIT = 5 };  // TC_WAIT: timeline only

static inline cudaEvent_t timing_event(ka9q_stream* s) {
  if (s->ev_next == s->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    s->ev_pool.push_back(e);
  }
  return s->ev_pool[s->ev_next++];
}
struct TimedRegion {
  ka9q_stream* s;
  cudaStream_t st;
  cudaEvent_t e1 = nullptr;
  TimedRegion(ka9q_stream* s_, int cls, cudaStream_t st_) : s(s_), st(st_) {
    if (!s->timing || !s->timing_regions || s->ev_used.size() >= 4096) return;
    cudaEvent_t e0 = timing_event(s);
    e1 = timing_event(s);
    if (!e0 || !e1) {
      e1 = nullptr;
      return;
    }
    cudaEventRecord(e0, st);
    s->ev_used.push_back({cls, {e0, e1}});
  }
  ~TimedRegion() {
    if (e1) cudaEventRecord(e1, st);
  }
};


// issue helpers (stream.cu; defined inside its extern "C" block)
extern "C" {
float2* spec_buf(ka9q_stream* s, int p);
int issue_fft(ka9q_stream* s, long long first_block, int blk_first, int blk_count);
int publish_spectrum(ka9q_stream* s);
int issue_channels(ka9q_stream* s, int nblocks);
}
void mgpu_release(ka9q_stream* s);         // mgpu.cu: closes IPC mappings, frees the exchange buffers
int mgpu_wait_ready(ka9q_stream* s, int p);  // mgpu.cu: channel stream waits for every producer's arcs (P2P transport)

#define K9_CHECK(cond, ...)        \
  do {                             \
    if (!(cond)) {                 \
      k9::set_error(__VA_ARGS__);  \
      return -1;                   \
    }                              \
  } while (0)
