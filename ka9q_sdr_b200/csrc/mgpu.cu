// Channel-sharded multi-GPU channelizer (BASELINE.json north_star: "channels are sharded across the 8 GPUs
// of one box"). One process per GPU; every rank holds a contiguous (in carrier frequency) share of the channels and never
// moves channel state. Per batch of nblocks 20 ms blocks:
//
//   1. the forward FFT is sharded BY BLOCK: rank q transforms blocks [q*nb/G, (q+1)*nb/G) of the batch (it needs the int16
//      samples of those blocks plus the M-1 samples of overlap in front of them, nothing else);
//   2. exchange: a rank's channels only read the arc [min bin - 1023, max bin + 1024] of the N-bin spectrum (1/G of it plus
//      a 2048-bin halo), so the producer of a block sends every peer just that peer's arc — 21 MB x (1/G + halo) per block
//      and peer instead of the 21 MB a broadcast moves. Two transports:
//        KA9Q_MGPU_P2P   (default) a copy kernel of this library stores the arcs straight into the peers' spectrum buffers
//                        over NVLink (peer memory mapped with CUDA IPC), then raises a per-producer sequence flag in the
//                        peer's memory; the consumer's channel stream spins on its flags in a one-warp kernel. Flow
//                        control in the other direction (the consumer has finished reading buffer p) uses the same flags.
//                        No host round trip, no collective library on the data path.
//        KA9Q_MGPU_NCCL  grouped ncclSend/ncclRecv of the same arcs (cross-check and fallback).
//   3. every rank runs its channel kernels on its own share.
//
// The spectrum is double-buffered, so the FFT + exchange of batch k+1 overlap the channel kernels of batch k.
// Results equal the single-GPU run: AM / linear bit for bit, FM within 1 LSB where the pair partner differs.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "stream_priv.cuh"

namespace {

constexpr int SCATTER_THREADS = 256;
constexpr long long SPIN_TIMEOUT_CYCLES = 6000000000ll;  // ~3 s: a dead peer must not hang the GPU

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Stores the arcs of this rank's freshly transformed blocks into the peers' spectrum buffers (NVLink peer stores), then —
// once every CTA's stores are performed at system scope — raises ready[p][me] = seq in every peer's flag block.
__global__ void __launch_bounds__(SCATTER_THREADS) mgpu_scatter_kernel(const CopyJob* __restrict__ jobs, int njobs,
                                                                       unsigned* counter, MgpuFlags* const* peer_flags,
                                                                       int nranks, int me, int p, int seq) {
  // All jobs form ONE index space, so every thread keeps eight independent 16-byte loads in flight whatever the size of
  // the single arcs (a loop over the jobs, each spread over the whole grid, left one short round trip per job: 52 us
  // for 18 MB at 8 GPUs, half the NVLink rate).
  __shared__ CopyJob sj[32];
  for (int j = threadIdx.x; j < njobs; j += blockDim.x) sj[j] = jobs[j];
  __syncthreads();
  const long long total = sj[njobs - 1].first + sj[njobs - 1].n16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto locate = [&](long long i, const int4*& src, int4*& dst) {
    int j = 0;
    while (j + 1 < njobs && i >= sj[j + 1].first) j++;
    src = sj[j].src + (i - sj[j].first);
    dst = sj[j].dst + (i - sj[j].first);
  };
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 8 * stride) {
    int4 v[8];
    int4* d[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long i = i0 + u * stride;
      d[u] = nullptr;
      if (i < total) {
        const int4* sp;
        locate(i, sp, d[u]);
        v[u] = __ldg(sp);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (d[u]) *d[u] = v[u];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *counter = 0;
    __threadfence_system();
    if (threadIdx.x < nranks && threadIdx.x != me) st_release_sys(&peer_flags[threadIdx.x]->ready[p][me], seq);
  }
}

// one warp: lane r waits until flags[r] >= want (r != me, r < nranks)
__global__ void mgpu_wait_kernel(MgpuFlags* flags, int which /* 0 ready, 1 freed */, int p, int nranks, int me, int want) {
  const int r = threadIdx.x;
  if (r >= nranks || r == me) return;
  const int* f = which ? &flags->freed[p][r] : &flags->ready[p][r];
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < want) {
    __nanosleep(200);
    if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) {
      flags->error = 1;
      break;
    }
  }
}

// producer -> consumers after copy-engine transfers: the arcs of batch seq have landed in every peer's buffer p
__global__ void mgpu_signal_ready_kernel(MgpuFlags* const* peer_flags, int nranks, int me, int p, int seq) {
  const int r = threadIdx.x;
  if (r >= nranks || r == me) return;
  __threadfence_system();
  st_release_sys(&peer_flags[r]->ready[p][me], seq);
}

// consumer -> producers: this rank has finished reading its spectrum buffer p for batch seq
__global__ void mgpu_signal_free_kernel(MgpuFlags* const* peer_flags, int nranks, int me, int p, int seq) {
  const int r = threadIdx.x;
  if (r >= nranks || r == me) return;
  __threadfence_system();
  st_release_sys(&peer_flags[r]->freed[p][me], seq);
}

struct IpcBlob {  // what ka9q_stream_mgpu_export writes (128 bytes)
  cudaIpcMemHandle_t spec, flags;
};
static_assert(sizeof(IpcBlob) == 128, "IPC blob layout");

}  // namespace

void mgpu_release(ka9q_stream* s) {
  for (int r = 0; r < K9_MAX_RANKS; r++) {
    if (s->mg_ipc_opened[r]) {
      if (s->mg_peer_spec[r]) cudaIpcCloseMemHandle(s->mg_peer_spec[r]);
      if (s->mg_peer_flags[r]) cudaIpcCloseMemHandle(s->mg_peer_flags[r]);
    }
    s->mg_peer_spec[r] = nullptr;
    s->mg_peer_flags[r] = nullptr;
    s->mg_ipc_opened[r] = false;
  }
  if (s->d_flags) cudaFree(s->d_flags);
  if (s->d_mg_jobs) cudaFree(s->d_mg_jobs);
  if (s->d_mg_counter) cudaFree(s->d_mg_counter);
  if (s->d_mg_peer_flag_ptrs) cudaFree(s->d_mg_peer_flag_ptrs);
  if (s->d_mg_mask) cudaFree(s->d_mg_mask);
  s->d_mg_mask = nullptr;
  s->mg_fused = false;
  s->d_flags = nullptr;
  s->d_mg_jobs = nullptr;
  s->d_mg_counter = nullptr;
  s->d_mg_peer_flag_ptrs = nullptr;
  s->mg_nranks = 0;
}

extern "C" {

// The arc of the spectrum this rank's channels read: bins [lo, lo + len) mod N, 16-bin (128-byte) aligned. Channels are
// taken on the signed grid (bin > N/2 counts as bin - N), which is how a frequency-contiguous shard is contiguous.
int ka9q_stream_needed_bins(ka9q_stream* s, long long* lo, long long* len) {
  K9_CHECK(s && s->committed && lo && len, "bad argument");
  const long long N = s->N;
  long long mn = 0, mx = 0;
  bool first = true;
  for (const ka9q_chan_params& p : s->chans) {
    long long b = p.bin > N / 2 ? p.bin - N : p.bin;
    if (first || b < mn) mn = b;
    if (first || b > mx) mx = b;
    first = false;
  }
  long long a = mn - 1023, e = mx + 1024 + 1;  // [a, e)
  a = (a >= 0 ? a / 16 : -((-a + 15) / 16)) * 16;
  e = (e >= 0 ? (e + 15) / 16 : -((-e) / 16)) * 16;
  long long n = e - a;
  if (n >= N) {
    *lo = 0;
    *len = N;
  } else {
    *lo = ((a % N) + N) % N;
    *len = n;
  }
  return 0;
}

// 128-byte handle blob of this rank's spectrum buffers and flag block, to be all-gathered by the host plumbing
int ka9q_stream_mgpu_export(ka9q_stream* s, void* blob128) {
  K9_CHECK(s && s->committed && blob128, "bad argument");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (!s->d_flags) {
    K9_CUDA(cudaMalloc(&s->d_flags, sizeof(MgpuFlags)));
    K9_CUDA(cudaMemset(s->d_flags, 0, sizeof(MgpuFlags)));
  }
  IpcBlob b;
  K9_CUDA(cudaIpcGetMemHandle(&b.spec, s->d_spec));
  K9_CUDA(cudaIpcGetMemHandle(&b.flags, s->d_flags));
  memcpy(blob128, &b, sizeof(b));
  return 0;
}

// transport: KA9Q_MGPU_NCCL (needs ka9q_stream_nccl_init first) or KA9Q_MGPU_P2P (needs the peers' blobs).
// lo_all / len_all: every rank's ka9q_stream_needed_bins. blobs: nranks x 128 bytes (P2P only, may be NULL for NCCL).
int ka9q_stream_mgpu_setup(ka9q_stream* s, int transport, int rank, int nranks, const long long* lo_all,
                           const long long* len_all, const void* blobs) {
  K9_CHECK(s && s->committed && lo_all && len_all, "bad argument");
  K9_CHECK(nranks >= 1 && nranks <= K9_MAX_RANKS && rank >= 0 && rank < nranks, "bad rank / nranks");
  K9_CHECK(transport == KA9Q_MGPU_NCCL || transport == KA9Q_MGPU_P2P, "bad transport");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  if (transport == KA9Q_MGPU_NCCL)
    K9_CHECK(nranks == 1 || (s->nccl_comm && s->nccl_nranks == nranks && s->nccl_rank == rank && p_ncclSend && p_ncclRecv &&
                             p_ncclGroupStart && p_ncclGroupEnd),
             "NCCL transport: call ka9q_stream_nccl_init(rank, nranks) first");
  const long long N = s->N;
  s->mg_need.assign(nranks, {});
  for (int r = 0; r < nranks; r++) {
    long long lo = lo_all[r], len = len_all[r];
    K9_CHECK(lo >= 0 && lo < N && len > 0 && len <= N && lo % 16 == 0 && len % 16 == 0, "bad arc for rank %d", r);
    if (lo + len <= N) {
      s->mg_need[r].push_back({lo, len});
    } else {
      s->mg_need[r].push_back({lo, N - lo});
      s->mg_need[r].push_back({0, lo + len - N});
    }
  }
  {  // the exchange kernels run beside the channel kernels of the previous batch: same carve-out (see bigfft_r128.cuh)
    cudaFuncSetAttribute(mgpu_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, s->carveout);
    cudaFuncSetAttribute(mgpu_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, s->carveout);
    cudaFuncSetAttribute(mgpu_signal_free_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, s->carveout);
  }
  if (nranks > 1) {
    // The transform and the exchange of the next batch go ahead of the channel kernels' pending CTAs (stream.cu, where
    // s_fft is created). Measured at 2 GPUs: 0.298 -> 0.282 ms per step; the exchange then runs under the channel kernels.
    const char* ev = getenv("KA9Q_B200_FFT_PRIO");
    if (!ev || atoi(ev) != 0) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      K9_CUDA(cudaStreamSynchronize(s->s_fft));
      K9_CUDA(cudaStreamDestroy(s->s_fft));
      s->s_fft = nullptr;
      K9_CUDA(cudaStreamCreateWithPriority(&s->s_fft, cudaStreamNonBlocking, hi));
    }
    // flag waits and flag signals run on streams of their own: between two channel launches the channel stream then
    // only waits for an event instead of launching two one-warp kernels (about 10 us of an 8-GPU step)
    {
      const char* ex = getenv("KA9Q_B200_MGPU_XSTREAM");
      if (!s->s_mgx && (!ex || atoi(ex) != 0)) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        K9_CUDA(cudaStreamCreateWithPriority(&s->s_mgx, cudaStreamNonBlocking, hi));
        K9_CUDA(cudaEventCreateWithFlags(&s->e_mg_fft, cudaEventDisableTiming));
      }
    }
    if (!s->s_mgwait) K9_CUDA(cudaStreamCreateWithFlags(&s->s_mgwait, cudaStreamNonBlocking));
    if (!s->s_mgsig) K9_CUDA(cudaStreamCreateWithFlags(&s->s_mgsig, cudaStreamNonBlocking));
    if (!s->e_mg_ready) K9_CUDA(cudaEventCreateWithFlags(&s->e_mg_ready, cudaEventDisableTiming));
    if (!s->e_mg_chan) K9_CUDA(cudaEventCreateWithFlags(&s->e_mg_chan, cudaEventDisableTiming));
  }
  s->mg_rank = rank;
  s->mg_nranks = nranks;
  s->mg_transport = transport;
  s->mg_seq = 0;
  if (transport == KA9Q_MGPU_P2P && nranks > 1) {
    K9_CHECK(blobs, "P2P transport needs the peers' handle blobs");
    K9_CHECK(s->d_flags, "call ka9q_stream_mgpu_export before setup");
    std::vector<MgpuFlags*> fl(K9_MAX_RANKS, nullptr);
    for (int r = 0; r < nranks; r++) {
      if (r == rank) {
        s->mg_peer_spec[r] = s->d_spec;
        s->mg_peer_flags[r] = s->d_flags;
      } else {
        IpcBlob b;
        memcpy(&b, (const char*)blobs + (size_t)r * sizeof(IpcBlob), sizeof(b));
        void *ps = nullptr, *pf = nullptr;
        K9_CUDA(cudaIpcOpenMemHandle(&ps, b.spec, cudaIpcMemLazyEnablePeerAccess));
        K9_CUDA(cudaIpcOpenMemHandle(&pf, b.flags, cudaIpcMemLazyEnablePeerAccess));
        s->mg_peer_spec[r] = (float2*)ps;
        s->mg_peer_flags[r] = (MgpuFlags*)pf;
        s->mg_ipc_opened[r] = true;
      }
      fl[r] = s->mg_peer_flags[r];
    }
    K9_CUDA(cudaMalloc(&s->d_mg_peer_flag_ptrs, sizeof(MgpuFlags*) * K9_MAX_RANKS));
    K9_CUDA(cudaMemcpy(s->d_mg_peer_flag_ptrs, fl.data(), sizeof(MgpuFlags*) * K9_MAX_RANKS, cudaMemcpyHostToDevice));
    K9_CUDA(cudaMalloc(&s->d_mg_counter, sizeof(unsigned)));
    K9_CUDA(cudaMemset(s->d_mg_counter, 0, sizeof(unsigned)));
    K9_CHECK(s->cfg.max_blocks % nranks == 0, "max_blocks must be a multiple of the number of ranks");
    // Three forms of the P2P exchange (all parity-tested, tests/test_gpu_multi.py):
    //   push  (default)               producers STORE the arcs into the consumers' buffers (copy kernel, or the copy
    //                                 engines with KA9Q_B200_MGPU_CE=1)
    //   pull  (KA9Q_B200_MGPU_PULL=1) consumers LOAD their arcs out of the producers' spectrum buffers
    //   fused (KA9Q_B200_MGPU_FUSED=1) the forward FFT's last pass stores each 128-byte output row straight into the
    //                                 spectrum buffer of every rank that reads it (plans ending in the lean 160-point pass)
    const char* ef = getenv("KA9Q_B200_MGPU_FUSED");
    s->mg_fused = bigfft_can_route(&s->fwd) && ef && atoi(ef) != 0 && nranks <= 16;   // opt-in: measured slower
    const char* ep = getenv("KA9Q_B200_MGPU_PULL");
    s->mg_pull = !s->mg_fused && ep && atoi(ep) != 0;                                   // opt-in: measured slower too
    K9_CHECK(!(s->mg_fused && s->n0_enabled), "the noise-density estimate needs the whole spectrum on every rank: not "
                                               "available with the sharded exchange");
    if (s->mg_fused) {
      std::vector<unsigned short> mask((size_t)N / 16, 0);
      for (int r = 0; r < nranks; r++)
        for (const MgpuSeg& sg : s->mg_need[r])
          for (long long tile = sg.lo / 16; tile < (sg.lo + sg.len) / 16; tile++) mask[(size_t)tile] |= (unsigned short)(1u << r);
      K9_CUDA(cudaMalloc(&s->d_mg_mask, sizeof(unsigned short) * mask.size()));
      K9_CUDA(cudaMemcpy(s->d_mg_mask, mask.data(), sizeof(unsigned short) * mask.size(), cudaMemcpyHostToDevice));
      for (int r = 0; r < nranks; r++)
        s->mg_delta[r] = (long long)((const char*)s->mg_peer_spec[r] - (const char*)s->d_spec);
    }
  }
  return 0;
}

// Sample range (absolute stream positions) this rank has to hold in its ring to transform its share of the batch that
// starts at block `first_block`: its blocks plus the M-1 samples of overlap in front of them.
int ka9q_stream_mgpu_input_range(ka9q_stream* s, long long first_block, int nblocks, long long* first_sample,
                                 long long* nsamples) {
  K9_CHECK(s && s->committed && first_sample && nsamples, "bad argument");
  const int G = s->mg_nranks > 0 ? s->mg_nranks : 1;
  K9_CHECK(nblocks >= 1 && nblocks % G == 0, "nblocks must be a multiple of the number of ranks");
  const long long L = s->cfg.L, M = s->cfg.M;
  const int cnt = nblocks / G, bf = s->mg_rank * cnt;
  long long a = (first_block + bf) * L - (M - 1), e = (first_block + bf + cnt) * L;
  if (a < 0) a = 0;
  *first_sample = a;
  *nsamples = e - a;
  return 0;
}

// H2D copy of samples [first_sample, first_sample + nsamples) of the stream to their place in the device ring (the
// block-sharded forward FFT reads only part of every batch). Async on the copy stream, like ka9q_stream_push.
int ka9q_stream_push_at(ka9q_stream* s, const void* iq, long long first_sample, long long nsamples) {
  K9_CHECK(s && s->committed && iq && first_sample >= 0 && nsamples > 0, "bad argument");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  const long long end = first_sample + nsamples;
  K9_CHECK(end - s->block0 * (long long)s->cfg.L <= 2LL * s->cfg.max_blocks * s->cfg.L,
           "push would overwrite samples that have not been computed yet");
  K9_CUDA(cudaStreamWaitEvent(s->s_in, s->e_comp_done[s->comp_parity], 0));
  long long pos = (first_sample + (s->cfg.M - 1)) % s->ring_cap, done = 0;
  while (done < nsamples) {
    const long long chunk = std::min(nsamples - done, s->ring_cap - pos);
    K9_CUDA(cudaMemcpyAsync((char*)s->d_ring + pos * s->bytes_per_samp, (const char*)iq + done * s->bytes_per_samp,
                            (size_t)chunk * s->bytes_per_samp, cudaMemcpyHostToDevice, s->s_in));
    done += chunk;
    pos = (pos + chunk) % s->ring_cap;
  }
  if (end > s->pushed) s->pushed = end;
  K9_CUDA(cudaEventRecord(s->e_pushed, s->s_in));
  return 0;
}

static int exchange_nccl(ka9q_stream* s, int nblocks, int p) {
  const int G = s->mg_nranks, me = s->mg_rank, cnt = nblocks / G;
  float2* buf = spec_buf(s, p);
  int r = p_ncclGroupStart();
  for (int peer = 0; peer < G && r == 0; peer++) {
    if (peer == me) continue;
    for (int b = me * cnt; b < (me + 1) * cnt && r == 0; b++)
      for (const MgpuSeg& sg : s->mg_need[peer])
        if (r == 0) r = p_ncclSend(buf + (size_t)b * s->N + sg.lo, (size_t)2 * sg.len, /*ncclFloat32*/ 7, peer, (k9_ncclComm_t)s->nccl_comm, s->s_fft);
    for (int b = peer * cnt; b < (peer + 1) * cnt && r == 0; b++)
      for (const MgpuSeg& sg : s->mg_need[me])
        if (r == 0) r = p_ncclRecv(buf + (size_t)b * s->N + sg.lo, (size_t)2 * sg.len, 7, peer, (k9_ncclComm_t)s->nccl_comm, s->s_fft);
  }
  const int r2 = p_ncclGroupEnd();
  if (r == 0) r = r2;
  K9_CHECK(r == 0, "NCCL sub-band exchange: %s", p_ncclGetErrorString ? p_ncclGetErrorString(r) : "error");
  return 0;
}

// plain copy of a job list (used by the pull form; the push form's mgpu_scatter_kernel also raises the flags)
__global__ void __launch_bounds__(SCATTER_THREADS) mgpu_gather_kernel(const CopyJob* __restrict__ jobs, int njobs) {
  __shared__ CopyJob sj[32];
  for (int j = threadIdx.x; j < njobs; j += blockDim.x) sj[j] = jobs[j];
  __syncthreads();
  const long long total = sj[njobs - 1].first + sj[njobs - 1].n16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 8 * stride) {
    int4 v[8];
    int4* d[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long i = i0 + u * stride;
      d[u] = nullptr;
      if (i < total) {
        int j = 0;
        while (j + 1 < njobs && i >= sj[j + 1].first) j++;
        // peer memory is read around the local L1 / L2 (ld.global.cg: nothing stale from two batches ago can be hit)
        v[u] = __ldcg(sj[j].src + (i - sj[j].first));
        d[u] = sj[j].dst + (i - sj[j].first);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (d[u]) *d[u] = v[u];
  }
}

// Pull form: this rank loads its arcs of the peers' blocks out of the peers' spectrum buffers. On the FFT stream, after
// this rank's own transform:
//   ready[p][me] = seq on every peer -> wait for every peer's ready -> gather -> freed[p][me] = seq on every peer
// (a producer may overwrite its buffer p two batches later only when every consumer has pulled: waited for before its FFT)
static int exchange_pull(ka9q_stream* s, int nblocks, int p, int seq) {
  const int G = s->mg_nranks, me = s->mg_rank, cnt = nblocks / G;
  if (s->mg_jobs_nblocks != nblocks) {
    std::vector<CopyJob> jobs[K9_MAX_SPEC];
    for (int par = 0; par < s->nspec; par++)
      for (int peer = 0; peer < G; peer++) {
        if (peer == me) continue;
        for (int b = peer * cnt; b < (peer + 1) * cnt; b++)
          for (const MgpuSeg& sg : s->mg_need[me]) {
            const size_t off = ((size_t)par * s->cfg.max_blocks + b) * s->N + sg.lo;
            const long long first = jobs[par].empty() ? 0 : jobs[par].back().first + jobs[par].back().n16;
            jobs[par].push_back({(const int4*)(s->mg_peer_spec[peer] + off), (int4*)(s->d_spec + off), sg.len / 2, first});
          }
      }
    K9_CHECK(jobs[0].size() <= 32, "too many exchange segments (max 32 per batch)");
    s->mg_njobs = (int)jobs[0].size();
    if (s->d_mg_jobs) cudaFree(s->d_mg_jobs);
    s->d_mg_jobs = nullptr;
    if (s->mg_njobs) {
      K9_CUDA(cudaMalloc(&s->d_mg_jobs, sizeof(CopyJob) * s->nspec * s->mg_njobs));
      for (int par = 0; par < s->nspec; par++)
        K9_CUDA(cudaMemcpy((CopyJob*)s->d_mg_jobs + (size_t)par * s->mg_njobs, jobs[par].data(), sizeof(CopyJob) * s->mg_njobs,
                           cudaMemcpyHostToDevice));
    }
    s->mg_jobs_nblocks = nblocks;
  }
  static int ctas = 0;
  if (!ctas) {
    const char* e = getenv("KA9Q_B200_SCATTER_CTAS");
    ctas = e && atoi(e) > 0 ? atoi(e) : 592;
  }
  mgpu_signal_ready_kernel<<<1, 32, 0, s->s_fft>>>(s->d_mg_peer_flag_ptrs, G, me, p, seq);
  mgpu_wait_kernel<<<1, 32, 0, s->s_fft>>>(s->d_flags, 0, p, G, me, seq);
  mgpu_gather_kernel<<<ctas, SCATTER_THREADS, 0, s->s_fft>>>((const CopyJob*)s->d_mg_jobs + (size_t)p * s->mg_njobs, s->mg_njobs);
  mgpu_signal_free_kernel<<<1, 32, 0, s->s_fft>>>(s->d_mg_peer_flag_ptrs, G, me, p, seq);
  K9_CHECK(cudaGetLastError() == cudaSuccess, "gather kernel launch failed");
  return 0;
}

static int exchange_p2p(ka9q_stream* s, int nblocks, int p, int seq, cudaStream_t xs) {
  const int G = s->mg_nranks, me = s->mg_rank, cnt = nblocks / G;
  // copy-job list of this (parity, nblocks) shape, built once and cached on the device
  if (s->mg_jobs_nblocks != nblocks) {
    std::vector<CopyJob> jobs[K9_MAX_SPEC];
    for (int par = 0; par < s->nspec; par++)
      for (int peer = 0; peer < G; peer++) {
        if (peer == me) continue;
        for (int b = me * cnt; b < (me + 1) * cnt; b++)
          for (const MgpuSeg& sg : s->mg_need[peer]) {
            const size_t off = ((size_t)par * s->cfg.max_blocks + b) * s->N + sg.lo;
            const long long first = jobs[par].empty() ? 0 : jobs[par].back().first + jobs[par].back().n16;
            jobs[par].push_back({(const int4*)(s->d_spec + off), (int4*)(s->mg_peer_spec[peer] + off), sg.len / 2, first});
          }
      }
    K9_CHECK(jobs[0].size() <= 32, "too many exchange segments (max 32 per batch)");
    s->mg_njobs = (int)jobs[0].size();
    for (int par = 0; par < s->nspec; par++) {
      s->mg_host_jobs[par].clear();
      for (const CopyJob& j : jobs[par]) s->mg_host_jobs[par].push_back(j);
    }
    if (s->d_mg_jobs) cudaFree(s->d_mg_jobs);
    s->d_mg_jobs = nullptr;
    if (s->mg_njobs) {
      K9_CUDA(cudaMalloc(&s->d_mg_jobs, sizeof(CopyJob) * s->nspec * s->mg_njobs));
      for (int par = 0; par < s->nspec; par++)
        K9_CUDA(cudaMemcpy((CopyJob*)s->d_mg_jobs + (size_t)par * s->mg_njobs, jobs[par].data(), sizeof(CopyJob) * s->mg_njobs,
                           cudaMemcpyHostToDevice));
    }
    s->mg_jobs_nblocks = nblocks;
  }
  // the peers have finished reading their buffer p of two batches ago
  if (seq > s->nspec) {
    TimedRegion tw(s, TC_WAIT, xs);
    mgpu_wait_kernel<<<1, 32, 0, xs>>>(s->d_flags, 1, p, G, me, seq - s->nspec);
  }
  static int ctas = 0, use_ce = -1;
  if (!ctas) {
    const char* e = getenv("KA9Q_B200_SCATTER_CTAS");
    ctas = e && atoi(e) > 0 ? atoi(e) : 592;
    e = getenv("KA9Q_B200_MGPU_CE");
    use_ce = e ? atoi(e) != 0 : 1;
  }
  if (use_ce) {
    // Default: the arcs go through the copy engines (peer-to-peer cudaMemcpyAsync, no SM involved), then a one-warp
    // kernel raises the flags. The transfers take longer than the copy kernel's (0.09 vs 0.05 ms at 8 GPUs: 14 small
    // copies in a row) but the exchange is hidden under the channel kernels of the previous batch either way, and the
    // engines do not compete with them for SM slots and the L1: 0.161 -> 0.149 ms per step at 8 GPUs, 0.244 -> 0.228 at
    // 2. KA9Q_B200_MGPU_CE=0 selects the copy kernel.
    const std::vector<CopyJob>& hj = s->mg_host_jobs[p];
    for (const CopyJob& j : hj)
      K9_CUDA(cudaMemcpyAsync(j.dst, j.src, (size_t)j.n16 * 16, cudaMemcpyDeviceToDevice, xs));
    mgpu_signal_ready_kernel<<<1, 32, 0, xs>>>(s->d_mg_peer_flag_ptrs, G, me, p, seq);
    K9_CHECK(cudaGetLastError() == cudaSuccess, "signal kernel launch failed");
    return 0;
  }
  mgpu_scatter_kernel<<<ctas, SCATTER_THREADS, 0, xs>>>((const CopyJob*)s->d_mg_jobs + (size_t)p * s->mg_njobs, s->mg_njobs,
                                                            s->d_mg_counter, s->d_mg_peer_flag_ptrs, G, me, p, seq);
  K9_CHECK(cudaGetLastError() == cudaSuccess, "scatter kernel launch failed");
  return 0;
}

// One batch in channel-sharded mode: forward FFT of this rank's blocks, exchange of the arcs, channel kernels.
// resident != 0: re-run the most recently pushed batch (every rank holds the whole batch in its ring; benchmarks with the
// input already in HBM); resident == 0: streaming (push / push_at beforehand).
int ka9q_stream_mgpu_compute(ka9q_stream* s, int nblocks, int resident) {
  K9_CHECK(s && s->committed && s->mg_nranks >= 1, "call ka9q_stream_mgpu_setup first");
  const int G = s->mg_nranks, me = s->mg_rank;
  K9_CHECK(nblocks >= 1 && nblocks <= s->cfg.max_blocks && nblocks % G == 0, "nblocks must be a multiple of the number of ranks");
  K9_CUDA(cudaSetDevice(s->cfg.device));
  const int cnt = nblocks / G, bf = me * cnt;
  long long first_block;
  if (resident) {
    first_block = s->pushed / s->cfg.L - nblocks;
    K9_CHECK(first_block >= 0, "not enough samples resident in the ring");
    if (s->phase_block < first_block) s->phase_block = first_block;
  } else {
    first_block = s->block0;
    K9_CHECK((first_block + bf + cnt) * (long long)s->cfg.L <= s->pushed, "compute ahead of pushed samples");
  }
  const int p = s->spec_wr;
  const int seq = ++s->mg_seq;
  const bool fused = G > 1 && s->mg_transport == KA9Q_MGPU_P2P && s->mg_fused;
  const bool pull = G > 1 && s->mg_transport == KA9Q_MGPU_P2P && s->mg_pull;
  if ((fused || pull) && seq > s->nspec) {
    // fused: the last pass writes into the peers' buffer p, which they must have finished reading (two batches ago);
    // pull: the transform overwrites this rank's buffer p, which every peer must have finished pulling from
    K9_CUDA(cudaStreamWaitEvent(s->s_fft, s->e_spec_free[p], 0));
    mgpu_wait_kernel<<<1, 32, 0, s->s_fft>>>(s->d_flags, 1, p, G, me, seq - s->nspec);
  }
  s->mg_route_now = fused;
  const int fft_rc = issue_fft(s, first_block, bf, cnt);
  s->mg_route_now = false;
  if (fft_rc) return -1;
  if (fused) {
    TimedRegion tr(s, TC_BCAST, s->s_fft);
    mgpu_signal_ready_kernel<<<1, 32, 0, s->s_fft>>>(s->d_mg_peer_flag_ptrs, G, me, p, seq);
    K9_CHECK(cudaGetLastError() == cudaSuccess, "signal kernel launch failed");
  } else if (G > 1 && s->mg_transport == KA9Q_MGPU_P2P && !pull && s->s_mgx) {
    // The exchange runs on a stream of its own behind the transform: the FFT stream goes straight on to the next batch's
    // transform instead of idling while the arcs travel (at 4 GPUs the channel kernels fill every SM, the transform only
    // gets the SMs as they retire, and transform + exchange in one stream took longer than the channel kernels).
    K9_CUDA(cudaEventRecord(s->e_mg_fft, s->s_fft));
    K9_CUDA(cudaStreamWaitEvent(s->s_mgx, s->e_mg_fft, 0));
    TimedRegion tr(s, TC_BCAST, s->s_mgx);
    if (exchange_p2p(s, nblocks, p, seq, s->s_mgx)) return -1;
    s->pub_stream = s->s_mgx;
  } else if (G > 1) {
    TimedRegion tr(s, TC_BCAST, s->s_fft);
    if (s->mg_transport == KA9Q_MGPU_NCCL) {
      if (exchange_nccl(s, nblocks, p)) return -1;
    } else if (pull) {
      if (exchange_pull(s, nblocks, p, seq)) return -1;
    } else {
      if (exchange_p2p(s, nblocks, p, seq, s->s_fft)) return -1;
    }
  }
  if (publish_spectrum(s)) return -1;
  if (G > 1 && s->mg_transport == KA9Q_MGPU_P2P && !pull) {
    // the channel stream waits for every producer's arcs (issue_channels makes s_comp wait for e_spec_ready first)
    s->mg_wait_ready = seq;
  }
  if (issue_channels(s, nblocks)) return -1;
  if (G > 1 && s->mg_transport == KA9Q_MGPU_P2P && !pull) {
    // on the signal stream, behind the channel kernels: the channel stream itself goes straight on to the next batch
    K9_CUDA(cudaEventRecord(s->e_mg_chan, s->s_comp));
    K9_CUDA(cudaStreamWaitEvent(s->s_mgsig, s->e_mg_chan, 0));
    mgpu_signal_free_kernel<<<1, 32, 0, s->s_mgsig>>>(s->d_mg_peer_flag_ptrs, G, me, p, seq);
    K9_CHECK(cudaGetLastError() == cudaSuccess, "signal kernel launch failed");
    // later FFTs into buffer p also wait for this signal having been sent (keeps seq order on the wire)
    K9_CUDA(cudaEventRecord(s->e_spec_free[p], s->s_mgsig));
  }
  if (resident)
    s->block0 = s->pushed / s->cfg.L;
  else
    s->block0 += nblocks;
  return 0;
}

// P2P transport: has a wait on a peer flag timed out (a peer died)? Valid after ka9q_stream_sync.
int ka9q_stream_mgpu_error(ka9q_stream* s) {
  K9_CHECK(s && s->committed, "bad argument");
  if (!s->d_flags) return 0;
  int e = 0;
  K9_CUDA(cudaMemcpy(&e, &s->d_flags->error, sizeof(int), cudaMemcpyDeviceToHost));
  return e;
}

}  // extern "C"

// called by issue_channels (stream.cu) on the channel stream, after it has been made to wait for e_spec_ready[p]
int mgpu_wait_ready(ka9q_stream* s, int p) {
  if (s->mg_wait_ready <= 0) return 0;
  // the one-warp wait kernel spins on the wait stream from the moment the batch is issued; the channel stream only waits
  // for its event (already complete whenever the exchange finished under the previous batch's channel kernels)
  // (not before this rank's own exchange of the batch has finished: the peers' is then about done as well, so the kernel
  //  does not sit spinning for a whole batch)
  if (cudaStreamWaitEvent(s->s_mgwait, s->e_spec_ready[p], 0) != cudaSuccess) return -1;
  {
    TimedRegion tr(s, TC_WAIT, s->s_mgwait);
    mgpu_wait_kernel<<<1, 32, 0, s->s_mgwait>>>(s->d_flags, 0, p, s->mg_nranks, s->mg_rank, s->mg_wait_ready);
  }
  s->mg_wait_ready = 0;
  if (cudaGetLastError() != cudaSuccess) return -1;
  if (cudaEventRecord(s->e_mg_ready, s->s_mgwait) != cudaSuccess) return -1;
  return cudaStreamWaitEvent(s->s_comp, s->e_mg_ready, 0) == cudaSuccess ? 0 : -1;
}
