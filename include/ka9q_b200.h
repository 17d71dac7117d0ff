/* ka9q_b200.h — C ABI of libka9q_b200.so: the B200 (sm_100a) receive-DSP hot path of ka9q-radio.
 *
 * Two layers, both plain C (pointers and sizes only; no CUDA or torch types):
 *
 *  A. DROP-IN layer — the reference's own symbols with the reference's own struct layouts, so `radio`,
 *     `packet`, `modulate` and `hackrf` can link this library in place of filter.o / osc.o / decimate.o:
 *        filter.h:81-92   create_filter_input/output, execute_filter_input/output, delete_*, set_filter,
 *                         window_filter, window_rfilter, make_kaiser, noise_gain, Kaiser_beta
 *        osc.h:21-24      set_osc, step_osc, renorm_osc, is_phasor_init
 *        decimate.h:10-11 hb15_block, hb3_block
 *     FFTs, bin selection/response multiply, filter design and the half-band FIRs run on the GPU; the host
 *     buffers the callers poke directly (in->input.c, in->fdomain, out->output.c, out->response ...) are kept
 *     coherent mirrors. The oscillator is host-side scalar state by nature (one complex-double multiply per call).
 *
 *  B. BATCH layer (ka9q_*) — the channelizer the drop-in layer cannot express: ONE forward FFT per I/Q stream
 *     block shared by thousands of channels, each channel = bin rotation + response multiply + 2048-point
 *     inverse FFT + overlap discard + fused demodulator (FM / AM / linear) + int16 PCM. It replaces, per channel,
 *     one whole `radio` process: proc_samples (radio.c:41-150) -> execute_filter_input (filter.c:146) ->
 *     execute_filter_output (filter.c:175) -> demod_fm / demod_am / demod_linear (fm.c:21, am.c:15, linear.c:21)
 *     -> scaleclip (audio.c:22-28). Mode rows (modes.txt -> struct modetab, radio.h:33-48) and RTP I/O stay in C
 *     on the host and feed ka9q_chan_params / consume the PCM rows.
 *
 * All functions return 0 on success and a negative value on error unless stated otherwise;
 * ka9q_last_error() returns a thread-local description. There is NO CPU fallback: without a usable CUDA device
 * every compute entry point fails.
 */
#ifndef KA9Q_B200_H
#define KA9Q_B200_H 1

#include <pthread.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#define KA9Q_CFLOAT void /* complex float * in C */
#else
#include <complex.h>
#define KA9Q_CFLOAT complex float
#endif

/* ====================================================================================================
 * A. Drop-in layer
 * ==================================================================================================== */

/* reference filter.h:17-22 */
enum filtertype {
  NONE,
  COMPLEX,
  CROSS_CONJ, /* COMPLEX output with the cross-conjugation used for ISB (filter.c:239-249) */
  REAL,
};

/* reference filter.h:25-28 */
union rc {
  float *r;
  KA9Q_CFLOAT *c;
};

/* Binary-compatible with reference filter.h:54-66 (callers read/write fields directly: radio.c:89,139-140;
 * fm.c:131,141,160; radio.c:388-396 reads fdomain/ilen/impulse_length). fwd_plan is an opaque handle here. */
struct filter_in {
  enum filtertype in_type;       /* REAL or COMPLEX */
  unsigned int ilen;             /* L: user samples per block */
  unsigned int impulse_length;   /* M: impulse response length */
  KA9Q_CFLOAT *fdomain;          /* spectrum mirror: N bins (COMPLEX) or N/2+1 (REAL) */
  union rc input_buffer;         /* N = L + M - 1 samples */
  union rc input;                /* input_buffer + (M-1): the caller writes L new samples here */
  void *fwd_plan;                /* opaque (device context) */
  unsigned int blocknum;         /* incremented by execute_filter_input */
  pthread_mutex_t filter_mutex;
  pthread_cond_t filter_cond;
};

/* Binary-compatible with reference filter.h:67-80 (fm.c:93-94,129,169-170; am.c:55-56; linear.c:117-120,140-148,
 * 211-213,254,280,286,295,299; radio_status.c:171). */
struct filter_out {
  struct filter_in *master;
  enum filtertype out_type;
  KA9Q_CFLOAT *response;         /* owned by the slave once passed in (filter.c:271, :539-543) */
  pthread_mutex_t response_mutex;
  KA9Q_CFLOAT *f_fdomain;
  float noise_gain;
  union rc output_buffer;        /* N/decimate samples */
  union rc output;               /* output_buffer + N/decimate - olen */
  void *rev_plan;                /* opaque (device context) */
  unsigned int decimate;
  unsigned int olen;
  unsigned int blocknum;
};

/* filter.h:81-92 */
int window_filter(int L, int M, KA9Q_CFLOAT *response, float beta);
int window_rfilter(int L, int M, KA9Q_CFLOAT *response, float beta);
struct filter_in *create_filter_input(unsigned int L, unsigned int M, enum filtertype in_type);
struct filter_out *create_filter_output(struct filter_in *master, KA9Q_CFLOAT *response, unsigned int decimate,
                                        enum filtertype out_type);
int execute_filter_input(struct filter_in *);
int execute_filter_output(struct filter_out *);
int delete_filter_input(struct filter_in *);
int delete_filter_output(struct filter_out *);
int make_kaiser(float *window, unsigned int M, float beta);
int set_filter(struct filter_out *, float low, float high, float kaiser_beta);
float noise_gain(struct filter_out const *);
extern float Kaiser_beta; /* filter.c:279 */

/* Allocation helpers for `response` buffers handed to create_filter_output (the reference uses
 * fftwf_alloc_complex / fftwf_free; a drop-in build maps those two names onto these). */
void *ka9q_alloc(size_t bytes);
void ka9q_free(void *p);

/* reference osc.h:9-24 */
#ifndef __cplusplus
struct osc {
  double freq;
  double rate;
  complex double phasor;
  complex double phasor_step;
  complex double phasor_step_step;
  pthread_mutex_t mutex;
  int steps;
};
void set_osc(struct osc *osc, double f, double r);
complex double step_osc(struct osc *osc);
void renorm_osc(struct osc *osc);
int is_phasor_init(const complex double x);
#endif

/* set_osc(f, r) on a fresh oscillator, then n step_osc() results as interleaved (re, im) doubles (tests, diagnostics) */
int ka9q_osc_run(double f, double r, long n, double *out);

/* reference decimate.h:4-11 */
struct hb15_state {
  float coeffs[4];
  float even_samples[4];
  float odd_samples[4];
  float old_odd_samples[4];
};
void hb15_block(struct hb15_state *state, float *output, float *input, int cnt);
void hb3_block(float *state, float *output, float *input, int cnt);

/* ====================================================================================================
 * B. Batch layer
 * ==================================================================================================== */

typedef struct ka9q_stream ka9q_stream;

/* enum demod_type (radio.h:20-24) */
#define KA9Q_LINEAR_DEMOD 0
#define KA9Q_AM_DEMOD 1
#define KA9Q_FM_DEMOD 2
/* option flags of a mode row (modes.c:104-120) */
#define KA9Q_FLAG_ISB 1
#define KA9Q_FLAG_FLAT 2
#define KA9Q_FLAG_PLL 4
#define KA9Q_FLAG_SQUARE 8
/* wire formats of the I/Q input (multicast.h:19-24; radio.c:60-68,111-119) */
#define KA9Q_IQ_S16 1
#define KA9Q_IQ_S8 2

/* One receive channel = the per-channel part of struct demod (radio.h:64-193) after set_mode (radio.c:322-374). */
typedef struct ka9q_chan_params {
  int demod_type;        /* KA9Q_*_DEMOD */
  int flags;             /* KA9Q_FLAG_* */
  int channels;          /* 1 = mono, 2 = stereo I/Q (linear only) */
  int reserved;
  long long bin;         /* carrier on the N-point grid: f_c = bin * samprate / N (second LO = -f_c, radio.c:217) */
  float low, high;       /* filter edges, Hz (demod->filter.low/high) */
  float kaiser_beta;     /* demod->filter.kaiser_beta (main.c:115: 3.0) */
  float shift;           /* post-detection shift, Hz (demod->tune.shift; linear only) */
  float attack_rate;     /* dB/s, unused by the reference demodulators (am.c:28, linear.c:34-36) */
  float recovery_rate;   /* dB/s */
  float hangtime;        /* s */
  float headroom;        /* amplitude ratio; NAN -> pow(10,-15/20) (main.c:117) */
} ka9q_chan_params;

/* Per-channel, per-block status: the demod->sig.* scalars the reference publishes (radio.h:155-166). */
typedef struct ka9q_chan_status {
  float bb_power;
  float snr;
  float foffset;
  float pdeviation;
  float agc_gain;
  int squelch_open;
  float reserved[2];
} ka9q_chan_status;

typedef struct ka9q_stream_config {
  int device;            /* CUDA device ordinal */
  int samprate;          /* input sample rate, Hz */
  int L, M;              /* block and impulse lengths at the input rate (main.c:113-114) */
  int decimate;          /* samprate / output rate (radio_status.c:266); N/decimate must be 2048 */
  int iq_format;         /* KA9Q_IQ_S16 or KA9Q_IQ_S8 */
  float gain_factor;     /* demod->sdr.gain_factor (radio.c:122) */
  int max_blocks;        /* blocks processed per ka9q_stream_process call (>=1) */
  int capture_filter_output; /* nonzero: keep raw filter output for ka9q_stream_get_filter_output (tests) */
} ka9q_stream_config;

/* pinned (page-locked) host memory for I/Q and PCM buffers */
void *ka9q_host_alloc(size_t bytes);
void ka9q_host_free(void *p);

const char *ka9q_last_error(void);
const char *ka9q_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int ka9q_device_count(void);

int ka9q_stream_create(ka9q_stream **out, const ka9q_stream_config *cfg);
int ka9q_stream_destroy(ka9q_stream *s);
/* Add a channel; returns its index (>=0) or a negative error. Channels can only be added before commit. */
int ka9q_stream_add_channel(ka9q_stream *s, const ka9q_chan_params *p);
/* Designs all channel filters on the device (set_filter, filter.c:500-546), uploads parameters, allocates buffers. */
int ka9q_stream_commit(ka9q_stream *s);
/* Re-design one channel's filter after commit (the UI path: display.c:163-177). */
int ka9q_stream_set_filter(ka9q_stream *s, int chan, float low, float high, float kaiser_beta);
/* Off-grid carriers (the reference's second LO is any double: radio.c:217,299). The carrier of channel
 * `chan` sits at (bin + fine_bins) * samprate / N, |fine_bins| <= 0.5. The grid part stays a bin rotation; the fraction
 * becomes a phase ramp on the channel's impulse response (mixing before a filter h = filtering with h[m] e^{-j2 pi d m} and
 * mixing afterwards) plus a rotation of the kept samples at the output rate (FM: ahead of the discriminator; linear:
 * folded into the shift oscillator, radio.c:313; AM: the envelope does not see it). Call between add_channel and commit.
 * ISB and pll / square channels are refused (grid only). ka9q_stream_split_carrier does the arithmetic:
 * carrier_hz (from the first LO, either sign) -> nearest bin and the fraction left over. */
int ka9q_stream_set_fine_lo(ka9q_stream *s, int chan, double fine_bins);
int ka9q_stream_split_carrier(const ka9q_stream *s, double carrier_hz, long long *bin, double *fine_bins);

/* PL-tone analyser (pltask, fm.c:189-285) for every de-emphasised FM channel: a /32 REAL slave of the audio master filter
 * feeding a 16384-point transform every 0.34 s. Enable before commit; the tone frequency (demod->sig.plfreq: 0 until the
 * first analysis, NAN when no tone stands out) is ka9q_chan_status.reserved[1] of FM channels. */
int ka9q_stream_enable_pl(ka9q_stream *s, int enable);
/* Noise-density estimate compute_n0 (radio.c:383-425; fm.c:78-82, am.c:46-49, linear.c:123-126) for every channel and
 * block, computed once per stream on the device (csrc/n0.cu). Enable before commit; rows are fetched like the PCM. */
int ka9q_stream_enable_n0(ka9q_stream *s, int enable);
int ka9q_stream_fetch_n0(ka9q_stream *s, int nblocks, float *raw /* compute_n0() per block */,
                         float *smooth /* demod->sig.n0 after each block */);
int ka9q_stream_num_channels(const ka9q_stream *s);
/* int16 units per block row of the PCM output, and a channel's offset / channel count inside the row */
long long ka9q_stream_pcm_stride(const ka9q_stream *s);
int ka9q_stream_pcm_offset(const ka9q_stream *s, int chan);
int ka9q_stream_olen(const ka9q_stream *s);
/* blocks of the stream consumed so far (the next streaming compute starts at this block) */
long long ka9q_stream_blocks_done(const ka9q_stream *s);
int ka9q_stream_fft_size(const ka9q_stream *s);
int ka9q_stream_launches_per_call(const ka9q_stream *s);

/* End-to-end call. iq: nblocks*L interleaved I/Q samples in HOST memory (int16 or int8 per iq_format).
 * pcm: HOST buffer of nblocks * pcm_stride int16; status: HOST buffer of nblocks * nchan entries or NULL.
 * Copies H2D, runs forward FFT + all channels, copies D2H, returns when the results are in pcm/status. */
int ka9q_stream_process(ka9q_stream *s, const void *iq, int nblocks, int16_t *pcm, ka9q_chan_status *status);

/* Split form for pipelining and device-resident benchmarking:
 *   push     : H2D copy of nblocks*L samples into the device ring (async on the stream's copy stream)
 *   compute  : forward FFT + channel kernels for the nblocks most recently pushed (async)
 *   fetch    : D2H copy of PCM/status of the last compute (async)
 *   sync     : wait for everything issued so far
 * ka9q_stream_compute_resident re-runs compute on the samples already in the ring (inputs resident in HBM). */
int ka9q_stream_push(ka9q_stream *s, const void *iq, int nblocks);
int ka9q_stream_compute(ka9q_stream *s, int nblocks);
int ka9q_stream_compute_resident(ka9q_stream *s, int nblocks);
int ka9q_stream_fetch(ka9q_stream *s, int nblocks, int16_t *pcm, ka9q_chan_status *status);
int ka9q_stream_sync(ka9q_stream *s);
/* wait only until the H2D copies of push / push_at have finished reading their host buffers (e.g. before ka9q_rx_consume) */
int ka9q_stream_sync_input(ka9q_stream *s);
/* wait only for the D2H copies issued by ka9q_stream_fetch (PCM/status are double-buffered on the device, so the next
 * batch may already be computing): the steady-state loop is push(k+2); compute(k+1); fetch(k); wait_fetch() */
int ka9q_stream_wait_fetch(ka9q_stream *s);
/* wait for the fetch issued `batches_ago` fetches ago (0 = the latest, 1 = the one before): lets the host queue the
 * copy-out of batch k behind that of batch k-1 and only then wait for k-1, so the D2H link never idles on a host round
 * trip: push(k+1); compute(k); fetch(k); wait_fetched(1) -> batch k-1 is in host memory */
int ka9q_stream_wait_fetched(ka9q_stream *s, int batches_ago);
/* Time of the device work of the last compute call in milliseconds (CUDA events on the compute stream), and of
 * the dominant (channel) kernels alone. Valid after ka9q_stream_sync. */
int ka9q_stream_last_timing(ka9q_stream *s, float *total_ms, float *fft_ms, float *chan_ms);

/* Timed region for benchmarks: CUDA events on the compute stream around the region, plus one event pair around every
 * forward FFT and every channel-kernel launch inside it. class_ms[5]/class_launches[5] = {forward FFT, FM kernel,
 * AM kernel, linear kernel, NCCL spectrum broadcast}. timer_stop synchronises. */
/* 1 (default): the forward FFT of batch k+1 overlaps the channel kernels of batch k; 0: serialised (per-kernel timing) */
int ka9q_stream_set_overlap(ka9q_stream *s, int enable);
int ka9q_stream_timer_start(ka9q_stream *s);
/* as timer_start, but only the outer event pair is recorded (no per-kernel events inside the timed steps) */
int ka9q_stream_timer_start_plain(ka9q_stream *s);
int ka9q_stream_timer_stop(ka9q_stream *s, float *ms_total, float *class_ms, int *class_launches);
/* After timer_stop: the bracketed regions in issue order as (class, start, end), ms from timer_start; class 5 = waiting
 * for peer flags (multi-GPU). Returns the number of regions written (<= max_regions). */
int ka9q_stream_timer_timeline(ka9q_stream *s, int max_regions, int *cls, float *start_ms, float *end_ms);

/* Multi-GPU: channels are sharded by the caller (each rank adds only its own channels); the rank that owns the
 * I/Q input runs the forward FFT and broadcasts the spectrum. The caller supplies the broadcast as a callback
 * (NCCL in bench.py through torch.distributed, or ka9q_nccl_* below). When root < 0 every rank computes its own FFT. */
int ka9q_stream_spectrum_ptr(ka9q_stream *s, void **dev_ptr, long long *bytes_per_block);
int ka9q_stream_compute_fft_only(ka9q_stream *s, int nblocks);
int ka9q_stream_compute_channels_only(ka9q_stream *s, int nblocks);
/* forward FFT of blocks [first, first+count) of the resident batch only (FFT sharded by block across ranks) */
int ka9q_stream_compute_fft_blocks(ka9q_stream *s, int nblocks, int first, int count);
/* NCCL spectrum broadcast inside the library (libnccl.so.2 is dlopen'ed on first use). */
int ka9q_nccl_unique_id(void *id128);
int ka9q_stream_nccl_init(ka9q_stream *s, const void *id128, int rank, int nranks);
int ka9q_stream_nccl_broadcast_spectrum(ka9q_stream *s, int nblocks, int root);
/* every rank transformed its nblocks/nranks blocks (compute_fft_blocks); gather all nblocks spectra on every rank */
int ka9q_stream_nccl_allgather_spectrum(ka9q_stream *s, int nblocks);

/* Channel-sharded multi-GPU: one process per GPU, each holding a frequency-contiguous share of the channels.
 * Per batch the forward FFT is sharded by block (rank q transforms blocks [q*nb/G, (q+1)*nb/G)), every producer sends each
 * peer only the arc of the spectrum that peer's channels read, then every rank runs its own channel kernels. Transports:
 * P2P = a copy kernel of this library stores into the peers' spectrum buffers over NVLink (peer memory mapped with CUDA
 * IPC) and hands over with sequence flags in device memory; NCCL = grouped ncclSend / ncclRecv (cross-check).
 * The host plumbing (torch.distributed, MPI, a pipe ...) only all-gathers three small blobs at set-up time:
 *   ka9q_stream_needed_bins -> (lo, len) of this rank's arc;  ka9q_stream_mgpu_export -> 128-byte IPC handle blob. */
#define KA9Q_MGPU_NCCL 1
#define KA9Q_MGPU_P2P 2
int ka9q_stream_needed_bins(ka9q_stream *s, long long *lo, long long *len);
int ka9q_stream_mgpu_export(ka9q_stream *s, void *blob128);
int ka9q_stream_mgpu_setup(ka9q_stream *s, int transport, int rank, int nranks, const long long *lo_all,
                           const long long *len_all, const void *blobs /* nranks x 128 bytes, P2P only */);
/* sample range (absolute stream positions) this rank must hold to transform its blocks of the batch at first_block */
int ka9q_stream_mgpu_input_range(ka9q_stream *s, long long first_block, int nblocks, long long *first_sample,
                                 long long *nsamples);
/* H2D copy of stream samples [first_sample, first_sample + nsamples) to their place in the device ring */
int ka9q_stream_push_at(ka9q_stream *s, const void *iq, long long first_sample, long long nsamples);
/* one batch: FFT of this rank's blocks, exchange, channel kernels. resident != 0 re-runs the last pushed batch. */
int ka9q_stream_mgpu_compute(ka9q_stream *s, int nblocks, int resident);
/* nonzero if a wait on a peer timed out (P2P transport; valid after ka9q_stream_sync) */
int ka9q_stream_mgpu_error(ka9q_stream *s);

/* Introspection for parity tests (device -> host copies, synchronous). */
int ka9q_stream_get_response(ka9q_stream *s, int chan, KA9Q_CFLOAT *out2048, float *noise_gain);
int ka9q_stream_get_filter_output(ka9q_stream *s, int chan, int nblocks, KA9Q_CFLOAT *out /* nblocks*olen */);
int ka9q_stream_get_spectrum(ka9q_stream *s, int block, KA9Q_CFLOAT *outN);
int ka9q_stream_get_if_energy(ka9q_stream *s, int nblocks, float *energy /* sum |x|^2 per block */);

/* Generic batched complex FFT on host buffers (cross-checks and tests): sign -1 forward, +1 backward. */
int ka9q_fft_c2c(int device, int n, int batch, int sign, const KA9Q_CFLOAT *in, KA9Q_CFLOAT *out);
/* How n would be factorised into passes: fills sizes[0..3], returns number of passes or -1 if unsupported. */
int ka9q_fft_plan_describe(int n, int *sizes);

/* Half-band decimator cascade on the device (decimate.c:44-162 driven as hackrf.c:297-318):
 * `stages` x (15-tap half-band /2), highest rate first, separately per plane; state16 holds stages x hb15_state. */
int ka9q_hb15_cascade(int device, int stages, struct hb15_state *states, const float *in, int n_in, float *out);

/* Front-end decimator service: the sample path of the reference's `hackrf` daemon on the device —
 * rx_callback (hackrf.c:129-196: int8 ingest, DC removal, I/Q gain and phase correction, estimates advanced once per USB
 * transfer) and process (hackrf.c:198-345: Fs/4 rotation, cascade of 15-tap half-band decimators on both planes,
 * x 0.5^stages, (short)round(32767 s)). */
typedef struct ka9q_frontend ka9q_frontend;
typedef struct ka9q_frontend_config {
  int device;
  int out_samprate;      /* Out_samprate (hackrf.c:62); the A/D runs at decimate * out_samprate (hackrf.c:463) */
  int decimate;          /* Decimate: a power of two, 2..64 (hackrf.c:63,464) */
  int offset;            /* Offset: 1 = tuner set high by Fs/4, undone by a rotation per sample (hackrf.c:68,271-291) */
  int callback_samples;  /* complex samples per USB transfer = per update of the estimates (hackrf.c:132,182) */
  float dc_alpha;        /* DC_alpha (hackrf.c:74: 1e-7) */
  float power_alpha;     /* Power_alpha (hackrf.c:75: 1.0) */
} ka9q_frontend_config;
typedef struct ka9q_frontend_status {  /* HackCD.* (hackrf.c:36-58) */
  float dc_i, dc_q, imbalance, sinphi, in_power;
  long long clips, samples;
} ka9q_frontend_status;
int ka9q_frontend_create(ka9q_frontend **out, const ka9q_frontend_config *cfg);
int ka9q_frontend_destroy(ka9q_frontend *f);
/* iq8: HOST int8 I/Q, nsamples complex samples (a whole number of callback blocks); out: HOST int16 I/Q, nsamples/decimate */
int ka9q_frontend_process(ka9q_frontend *f, const void *iq8, long long nsamples, int16_t *out);
/* same, the decimated stream goes device-to-device into the channelizer's ring (as ka9q_stream_push would from the host) */
int ka9q_frontend_process_to_stream(ka9q_frontend *f, const void *iq8, long long nsamples, ka9q_stream *s);
/* re-run the device work on the batch left resident by the last process call (benchmarks); ms = device time */
int ka9q_frontend_rerun_resident(ka9q_frontend *f, long long nsamples, float *ms);
int ka9q_frontend_set_estimates(ka9q_frontend *f, float dc_i, float dc_q, float imbalance, float sinphi);
int ka9q_frontend_get_status(ka9q_frontend *f, ka9q_frontend_status *out);
/* append nsamples complex samples that already live in DEVICE memory (the stream's iq_format) to the stream's ring */
int ka9q_stream_push_device(ka9q_stream *s, const void *d_iq, long long nsamples);

/* ---------------------------------------------------------------------------------------------------------------
 * Wire-format glue either side of the path (host C, no GPU). What `radio` does between its sockets and
 * the DSP: I/Q datagram -> sample stream with the reference's sequence / timestamp repair, and PCM rows -> RTP packets.
 * The sockets themselves stay the caller's.
 * ------------------------------------------------------------------------------------------------------------- */
#define KA9Q_RTP_MIN_SIZE 12  /* multicast.h:15 */
#define KA9Q_IQ_PT 97         /* multicast.h:19: raw I/Q, 16 bit */
#define KA9Q_IQ_PT8 98        /* multicast.h:20: raw I/Q, 8 bit */
#define KA9Q_PCM_MONO_PT 11   /* multicast.h:22 */
#define KA9Q_PCM_STEREO_PT 10 /* multicast.h:23 */
#define KA9Q_PCM_BUFSIZE 480  /* audio.c:19: int16 words per PCM packet */

/* struct rtp_state (multicast.h:41-50), same fields and order */
typedef struct ka9q_rtp_state {
  uint32_t ssrc;
  int init;
  uint16_t seq;
  uint32_t timestamp;
  long long packets;
  long long bytes;
  long long drops;
  long long dupes;
} ka9q_rtp_state;

/* Receive side: replaces rtp_recv's parsing (main.c:313-344) and the packet head of proc_samples (radio.c:60-100). */
typedef struct ka9q_ingest {
  ka9q_rtp_state rtp; /* demod->input.rtp */
  int iq_format;      /* KA9Q_IQ_S16 or KA9Q_IQ_S8: the stream's sample format; datagrams of the other type are ignored */
  long long samples;  /* demod->input.samples: reset on SSRC change, counts zero-filled samples too */
  long long zero_filled;
  long long ignored;  /* too short, wrong payload type, duplicate, old, or a jump of more than 192000 samples */
} ka9q_ingest;
void ka9q_ingest_init(ka9q_ingest *g, int iq_format);
/* One received UDP payload: RTP header (CSRCs, extension and padding handled as ntoh_rtp / main.c:322-326 do), the
 * 24-byte legacy status header (main.c:340), then I/Q in host byte order. Appends to dst, in the stream's sample
 * format, first `time_step` zero samples if the timestamp jumped (lost packets, radio.c:81-100) and then the payload.
 * Returns the number of complex samples appended (>= 0), -1 if the datagram was ignored, -2 if dst has no room
 * (`room` complex samples; the RTP state is then left untouched so the call can be repeated). */
long long ka9q_ingest_datagram(ka9q_ingest *g, const void *datagram, int size, void *dst, long long room);
/* rtp_process (multicast.c:305-340) on already-parsed header fields. */
int ka9q_rtp_process(ka9q_rtp_state *state, uint32_t ssrc, uint16_t seq, uint32_t timestamp, int sampcnt);

/* Send side: send_mono_output / send_stereo_output (audio.c:32-132) for PCM that is already int16 (the channel kernels
 * apply scaleclip, audio.c:22-28). */
typedef struct ka9q_pcm_out {
  ka9q_rtp_state rtp; /* demod->output.rtp: ssrc, seq, timestamp, packets, bytes */
  int silent;         /* demod->output.silent */
} ka9q_pcm_out;
typedef int (*ka9q_emit_fn)(void *user, const void *packet, int len); /* return < 0 to stop (send() failed) */
/* Packetise `frames` frames of `channels` (1 or 2) interleaved host-order int16: chunks of at most 480 words, big
 * endian, 12-byte RTP header (PT 11 mono / 10 stereo); an all-zero chunk is not sent but still advances the timestamp,
 * and the first packet after silence carries the marker bit. Returns the number of packets emitted, -1 on bad
 * arguments. */
int ka9q_pcm_packetise(ka9q_pcm_out *out, const int16_t *pcm, int frames, int channels, ka9q_emit_fn emit, void *user);

/* Receive-side host plumbing (main.c:288-365 rtp_recv + the queue side of proc_samples, radio.c:51-100): a receive thread
 * inserts datagrams into a queue sorted by RTP sequence number (main.c:347-361), an ingest thread pops it in order, repairs
 * sequence / timestamp gaps with zeros (ka9q_ingest_datagram) and appends to a page-locked ring of whole 20 ms blocks that
 * ka9q_stream_push reads directly. ka9q_rx_inject / ka9q_rx_drain are the same two steps without threads or sockets. */
typedef struct ka9q_rx ka9q_rx;
typedef struct ka9q_rx_stats {
  long long datagrams_queued, inserted_out_of_order, samples, zero_filled, ignored, rtp_drops, rtp_dupes;
  int pinned;
} ka9q_rx_stats;
ka9q_rx *ka9q_rx_create(int iq_format, long long block_samples, int ring_blocks);
void ka9q_rx_destroy(ka9q_rx *rx);
int ka9q_rx_inject(ka9q_rx *rx, const void *datagram, int size);
long long ka9q_rx_drain(ka9q_rx *rx);
int ka9q_rx_start(ka9q_rx *rx, int fd /* bound, multicast-joined UDP socket owned by the caller */);
int ka9q_rx_stop(ka9q_rx *rx);
const void *ka9q_rx_peek_blocks(ka9q_rx *rx, int nblocks, int wait_ms);
int ka9q_rx_consume(ka9q_rx *rx, int nblocks);
long long ka9q_rx_blocks_ready(ka9q_rx *rx);
void ka9q_rx_get_stats(ka9q_rx *rx, ka9q_rx_stats *st);
/* Egress: one block row of ka9q_stream_fetch, every channel packetised as audio.c:32-132 (ka9q_pcm_packetise) and handed to
 * the kernel `batch` packets at a time with sendmmsg (the reference sends one packet per send(), audio.c:73,122). */
int ka9q_pcm_send_block(int fd, ka9q_pcm_out *outs, const int16_t *pcm_row, const int *offs, const int *channels, int nchan,
                        int frames, int batch);

/* Status side: the "signals" and demodulator section of the TLV status list that `radio` multicasts
 * (radio_status.c:171-203, encodings status.c:31-96: type byte, length byte, big-endian value with leading zero bytes
 * suppressed, EOL = 0 terminates), from one ka9q_chan_status row. Fields this library does not compute (NOISE_DENSITY,
 * PL_TONE, the PLL group) are left out; the delta compression of radio_status.c:compact_packet is the sender's business.
 * demod_type: 0 linear, 1 AM, 2 FM (radio.h:26-30). Returns the number of bytes written, -1 if `room` is too small. */
int ka9q_status_encode_signals(const ka9q_chan_status *st, int demod_type, int isb, float if_power, float noise_bandwidth,
                               int output_channels, unsigned char *buf, int room);

#ifdef __cplusplus
}
#endif
#endif /* KA9Q_B200_H */
