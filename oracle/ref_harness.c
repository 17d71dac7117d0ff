/* Harness around the VERBATIM reference sources (TEST INFRASTRUCTURE ONLY — never linked into the product).
 *
 * This file is compiled together with the reference's own, unmodified C files where they lie under
 * /root/reference (filter.c osc.c dsp.c decimate.c fm.c am.c linear.c radio.c modes.c multicast.c
 * status.c misc.c) by oracle/Makefile into oracle/_ref/libka9q_ref.so. It contains no DSP of its own:
 * it only (a) provides the globals and output stubs that main.c/audio.c would provide, (b) feeds
 * int16 I/Q packets into demod.input.queue exactly as rtp_recv does (main.c:349-362) and paces the
 * producer because filter.c has no back-pressure (filter.c:195-199), and (c) captures PCM after the
 * reference's float->int16 rule (audio.c:22-28) plus the per-block demod->sig.* scalars.
 */
#define _GNU_SOURCE 1
#include <assert.h>
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <semaphore.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <sys/resource.h>
#include <unistd.h>
#include <fftw3.h>
#undef I

#include "misc.h"
#include "dsp.h"
#include "osc.h"
#include "filter.h"
#include "radio.h"
#include "decimate.h"

/* ---- globals normally defined in main.c / multicast.c users ---- */
char Libdir[PATH_MAX] = "/root/reference";
int Mcast_ttl = 0;
int Verbose = 0;
int Tunestep = 0;
int SDR_correct = 0;

size_t strlcpy(char *dst, const char *src, size_t size) {
  size_t const n = strlen(src);
  if (size) {
    size_t const c = n >= size ? size - 1 : n;
    memcpy(dst, src, c);
    dst[c] = '\0';
  }
  return n;
}

/* ---- per-chain capture context, found from the demod pointer ---- */
struct ref_block_status {
  float bb_power, snr, foffset, pdeviation, if_power, n0, agc_gain, cphase;
  int pll_lock;
  int channels;
  float plfreq;   /* demod->sig.plfreq as pltask (fm.c:189-285) left it when the block's PCM was sent */
  int pad;
};

struct capture {
  struct demod *demod;
  int16_t *pcm;
  long pcm_cap, pcm_len;
  float *filt;      /* complex filter output per block (olen complex each), may be NULL */
  long filt_cap_blocks;
  struct ref_block_status *st;
  int st_cap;
  int nblocks;
  sem_t sem;
  struct capture *next;
};
static pthread_mutex_t Cap_mutex = PTHREAD_MUTEX_INITIALIZER;
static struct capture *Captures;

static struct capture *find_capture(struct demod const *d) {
  pthread_mutex_lock(&Cap_mutex);
  struct capture *c = Captures;
  while (c && c->demod != d)
    c = c->next;
  pthread_mutex_unlock(&Cap_mutex);
  return c;
}

/* float -> int16 exactly as audio.c:22-28 (scaleclip is static there, so it is restated here) */
static short clip16(float const x) {
  if (x >= 1.0)
    return SHRT_MAX;
  else if (x <= -1.0)
    return SHRT_MIN;
  return (short)(SHRT_MAX * x);
}

static void capture_block(struct demod *demod, const float *buffer, int nfloats, int channels) {
  struct capture *c = find_capture(demod);
  if (!c)
    return;
  for (int i = 0; i < nfloats; i++) {
    if (c->pcm && c->pcm_len < c->pcm_cap)
      c->pcm[c->pcm_len++] = clip16(buffer[i]);
  }
  struct filter_out *fo = demod->filter.out;
  if (c->filt && fo && c->nblocks < c->filt_cap_blocks && fo->out_type != REAL)
    memcpy(c->filt + 2 * (size_t)fo->olen * c->nblocks, fo->output.c, sizeof(complex float) * fo->olen);
  if (c->st && c->nblocks < c->st_cap) {
    struct ref_block_status *s = &c->st[c->nblocks];
    s->bb_power = demod->sig.bb_power;
    s->snr = demod->sig.snr;
    s->foffset = demod->sig.foffset;
    s->pdeviation = demod->sig.pdeviation;
    s->if_power = demod->sig.if_power;
    s->n0 = demod->sig.n0;
    s->agc_gain = demod->agc.gain;
    s->cphase = demod->sig.cphase;
    s->pll_lock = demod->sig.pll_lock;
    s->plfreq = demod->sig.plfreq;
    s->channels = channels;
  }
  c->nblocks++;
  sem_post(&c->sem);
}

/* stubs for audio.c:32 and audio.c:82 */
int send_stereo_output(struct demod *const demod, float const *buffer, int size) {
  capture_block(demod, buffer, 2 * size, 2);
  return 0;
}
int send_mono_output(struct demod *const demod, float const *buffer, int size) {
  capture_block(demod, buffer, size, 1);
  return 0;
}

/* ---- mode table helpers ---- */
int ref_modes_clear(void) {
  Nmodes = 0;
  memset(Modes, 0, sizeof(struct modetab) * 256);
  return 0;
}
int ref_modes_load(const char *dir) {
  ref_modes_clear();
  strlcpy(Libdir, dir, sizeof(Libdir));
  return readmodes("modes.txt");
}
int ref_modes_count(void) { return Nmodes; }
/* flags: bit0 isb, bit1 flat, bit2 pll, bit3 square */
int ref_modes_add(const char *name, int demod_type, float low, float high, float shift, float attack, float recovery,
                  float hang, int channels, int flags) {
  if (Nmodes >= 256)
    return -1;
  struct modetab *m = &Modes[Nmodes++];
  memset(m, 0, sizeof(*m));
  strlcpy(m->name, name, sizeof(m->name));
  m->demod_type = demod_type;
  m->low = low;
  m->high = high;
  m->shift = shift;
  m->attack_rate = attack;
  m->recovery_rate = recovery;
  m->hangtime = hang;
  m->channels = channels;
  m->isb = flags & 1;
  m->flat = (flags >> 1) & 1;
  m->pll = (flags >> 2) & 1;
  m->square = (flags >> 3) & 1;
  return 0;
}
int ref_modes_get(int i, char *name, int *demod_type, float *vals /*low high shift attack recovery hang*/, int *channels,
                  int *flags) {
  if (i < 0 || i >= Nmodes)
    return -1;
  struct modetab const *m = &Modes[i];
  strcpy(name, m->name);
  *demod_type = m->demod_type;
  vals[0] = m->low; vals[1] = m->high; vals[2] = m->shift;
  vals[3] = m->attack_rate; vals[4] = m->recovery_rate; vals[5] = m->hangtime;
  *channels = m->channels;
  *flags = (m->isb ? 1 : 0) | (m->flat ? 2 : 0) | (m->pll ? 4 : 0) | (m->square ? 8 : 0);
  return 0;
}

/* ---- full receive chain: proc_samples -> filter -> demod_* (radio.c:41, fm.c:21, am.c:15, linear.c:21) ---- */
struct ref_chain_args {
  const char *mode;
  int samprate;
  int L, M, decimate;
  double carrier_hz;   /* carrier position in the IF; second LO = -carrier_hz (radio.c:217) */
  double lo_cycles;    /* if not NAN: overrides the second LO frequency, cycles/sample */
  float low, high;     /* NAN -> mode table defaults (radio.c:346-354) */
  double shift;        /* NAN -> mode table default */
  float kaiser_beta;
  float gain_factor;
  float headroom;      /* NAN -> pow(10,-15/20) as main.c:117 */
  int pkt_samples;
  int pkt_type;        /* IQ_PT (97) or IQ_PT8 (98) */
  int channels;        /* 0 -> mode default */
};

static void *dummy_thread(void *arg) { return arg; }

static void push_packet(struct demod *demod, struct packet *pkt) {
  /* append at tail, as main.c:349-362 does for in-order packets */
  pkt->next = NULL;
  pthread_mutex_lock(&demod->input.qmutex);
  struct packet **pp = &demod->input.queue;
  while (*pp)
    pp = &(*pp)->next;
  *pp = pkt;
  pthread_cond_signal(&demod->input.qcond);
  pthread_mutex_unlock(&demod->input.qmutex);
}

/* iq: interleaved samples (int16 for IQ_PT, int8 for IQ_PT8). drop[p]!=0 -> packet p is lost in transit.
 * Returns number of blocks captured, <0 on error. */
int ref_chain_run(const struct ref_chain_args *a, const void *iq, long nsamples, const uint8_t *drop, int16_t *pcm,
                  long pcm_cap, long *pcm_len, float *filt, long filt_cap_blocks, struct ref_block_status *st, int st_cap) {
  if (!a || !iq || a->L <= 0 || a->M <= 0 || a->decimate <= 0)
    return -1;
  int const N = a->L + a->M - 1;
  {
    /* compute_n0 puts float[N] on the demod thread's stack (radio.c:390) */
    size_t need = (size_t)N * sizeof(float) + (8u << 20);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    size_t cur = 0;
    pthread_getattr_default_np(&attr);
    pthread_attr_getstacksize(&attr, &cur);
    if (cur < need) {
      pthread_attr_setstacksize(&attr, need);
      pthread_setattr_default_np(&attr);
    }
    pthread_attr_destroy(&attr);
  }
  struct demod *demod = calloc(1, sizeof(*demod));
  struct capture *cap = calloc(1, sizeof(*cap));
  cap->demod = demod;
  cap->pcm = pcm; cap->pcm_cap = pcm_cap;
  cap->filt = filt; cap->filt_cap_blocks = filt_cap_blocks;
  cap->st = st; cap->st_cap = st_cap;
  sem_init(&cap->sem, 0, 0);
  pthread_mutex_lock(&Cap_mutex);
  cap->next = Captures;
  Captures = cap;
  pthread_mutex_unlock(&Cap_mutex);

  /* defaults as main.c:107-127 */
  demod->output.samprate = a->samprate / a->decimate;
  demod->input.samprate = a->samprate;
  demod->sdr.status.samprate = a->samprate;
  demod->filter.L = a->L;
  demod->filter.M = a->M;
  demod->filter.decimate = a->decimate;
  demod->filter.interpolate = 1;
  demod->filter.kaiser_beta = a->kaiser_beta;
  demod->agc.headroom = isnan(a->headroom) ? pow(10., -15. / 20) : a->headroom;
  demod->sdr.gain_factor = a->gain_factor;
  demod->sdr.imbalance = 1;
  demod->filter.low = a->low;
  demod->filter.high = a->high;
  demod->tune.shift = a->shift;
  demod->tune.lock = 1;
  demod->input.source_address.ss_family = -1;
  double const LO1 = 100e6;
  demod->sdr.status.frequency = LO1;
  demod->sdr.min_IF = -0.5f * a->samprate;
  demod->sdr.max_IF = +0.5f * a->samprate;
  demod->tune.freq = LO1 + a->carrier_hz;

  pthread_mutex_init(&demod->sdr.status_mutex, NULL);
  pthread_cond_init(&demod->sdr.status_cond, NULL);
  pthread_mutex_init(&demod->doppler.mutex, NULL);
  pthread_mutex_init(&demod->shift.mutex, NULL);
  pthread_mutex_init(&demod->second_LO.mutex, NULL);
  pthread_mutex_init(&demod->input.qmutex, NULL);
  pthread_cond_init(&demod->input.qcond, NULL);

  demod->filter.in = create_filter_input(a->L, a->M, COMPLEX);
  pthread_t proc_thread;
  pthread_create(&proc_thread, NULL, proc_samples, demod);
  /* set_mode joins demod_thread first (radio.c:337): give it something joinable */
  pthread_create(&demod->demod_thread, NULL, dummy_thread, NULL);
  int const use_defaults = (isnan(a->low) || isnan(a->high)) ? 1 : 0;
  if (set_mode(demod, a->mode, use_defaults) != 0) {
    fprintf(stderr, "ref_chain_run: unknown mode %s\n", a->mode);
    return -2;
  }
  if (!use_defaults) {
    /* set_mode(defaults=0) keeps our edges; shift NAN -> table default (radio.c:355-356) */
  }
  if (a->channels)
    demod->output.channels = a->channels;
  if (!isnan(a->lo_cycles))
    set_osc(&demod->second_LO, a->lo_cycles, 0.0);

  /* feed packets */
  int const bytes_per_samp = (a->pkt_type == IQ_PT8) ? 2 : 4;
  int const pkt_samples = a->pkt_samples > 0 ? a->pkt_samples : 1024;
  long fed_since_block = 0; /* samples (delivered or zero-filled) since last full block boundary */
  long expected_blocks = 0;
  long pos = 0;
  uint16_t seq = 0;
  long pktno = 0;
  long consumed = 0; /* samples that proc_samples will have accounted for, incl. zero fill */
  while (pos < nsamples) {
    int n = (int)((nsamples - pos) < pkt_samples ? (nsamples - pos) : pkt_samples);
    int const lost = drop ? drop[pktno] : 0;
    if (!lost) {
      struct packet *pkt = calloc(1, sizeof(*pkt));
      pkt->rtp.version = 2;
      pkt->rtp.type = a->pkt_type ? a->pkt_type : IQ_PT;
      pkt->rtp.seq = seq;
      pkt->rtp.timestamp = (uint32_t)pos;
      pkt->rtp.ssrc = 1;
      pkt->data = pkt->content;
      pkt->len = n * bytes_per_samp;
      memcpy(pkt->content, (const uint8_t *)iq + (size_t)pos * bytes_per_samp, (size_t)pkt->len);
      /* everything up to pos+n is now accounted for (lost samples are zero-filled on the next delivery) */
      long const newly = (pos + n) - consumed;
      consumed = pos + n;
      push_packet(demod, pkt);
      fed_since_block += newly;
      while (fed_since_block >= a->L) {
        fed_since_block -= a->L;
        expected_blocks++;
        sem_wait(&cap->sem); /* pace: wait until the demodulator has emitted this block */
      }
    }
    seq++;
    pktno++;
    pos += n;
  }
  int const nblocks = cap->nblocks;
  if (pcm_len)
    *pcm_len = cap->pcm_len;

  /* shut down: stop capturing, ask the demod thread to exit, flush one more block of zeros through */
  pthread_mutex_lock(&Cap_mutex);
  for (struct capture **pp = &Captures; *pp; pp = &(*pp)->next) {
    if (*pp == cap) {
      *pp = cap->next;
      break;
    }
  }
  pthread_mutex_unlock(&Cap_mutex);
  /* demod_fm joins its PL-tone thread on exit (fm.c:176), and that thread only re-checks `terminate` after
   * another audio block (fm.c:238-239). So `terminate` must be raised while the demodulator is parked in
   * execute_filter_output, and one more block must follow; give it time to get there. */
  usleep(50000);
  demod->terminate = 1;
  {
    long remaining = a->L - fed_since_block;
    long ts = consumed;
    while (remaining > 0) {
      int n = (int)(remaining < pkt_samples ? remaining : pkt_samples);
      struct packet *pkt = calloc(1, sizeof(*pkt));
      pkt->rtp.version = 2;
      pkt->rtp.type = a->pkt_type ? a->pkt_type : IQ_PT;
      pkt->rtp.seq = seq++;
      pkt->rtp.timestamp = (uint32_t)ts;
      pkt->rtp.ssrc = 1;
      pkt->data = pkt->content;
      pkt->len = n * bytes_per_samp;
      push_packet(demod, pkt);
      ts += n;
      remaining -= n;
    }
  }
  pthread_join(demod->demod_thread, NULL);
  /* proc_samples never returns (radio.c:50); it is parked in pthread_cond_wait, a cancellation point */
  {
    /* wait until the queue has drained so no packet is leaked mid-processing */
    for (;;) {
      pthread_mutex_lock(&demod->input.qmutex);
      int const empty = demod->input.queue == NULL;
      pthread_mutex_unlock(&demod->input.qmutex);
      if (empty)
        break;
      sched_yield();
    }
  }
  pthread_cancel(proc_thread);
  pthread_join(proc_thread, NULL);
  delete_filter_input(demod->filter.in);
  sem_destroy(&cap->sem);
  free(cap);
  free(demod);
  (void)expected_blocks;
  return nblocks;
}

/* ---- filter-only path: execute_filter_input / execute_filter_output (filter.c:146,175), single thread ---- */
/* in_type/out_type use enum filtertype values (filter.h:17-22). `in` holds nblocks*L samples (complex
 * interleaved when in_type==COMPLEX, real otherwise); `out` receives nblocks*olen samples of out_type.
 * If response_in != NULL it is installed as the slave's response (copied; N_dec complex, or N_dec/2+1 when
 * out_type==REAL) instead of calling set_filter. fdomain_out (optional) receives the master's spectrum of the
 * LAST block. */
int ref_filter_run(int L, int M, int decimate, int in_type, int out_type, float low, float high, float beta,
                   const float *response_in, const float *in, int nblocks, float *out, float *response_out,
                   float *noise_gain_out, float *fdomain_out) {
  struct filter_in *master = create_filter_input(L, M, in_type);
  if (!master)
    return -1;
  int const N = L + M - 1;
  int const N_dec = N / decimate;
  int const rbins = (out_type == REAL) ? N_dec / 2 + 1 : N_dec;
  complex float *resp = NULL;
  if (response_in) {
    resp = fftwf_alloc_complex(N_dec);
    memset(resp, 0, sizeof(complex float) * N_dec);
    memcpy(resp, response_in, sizeof(complex float) * rbins);
  }
  struct filter_out *slave = create_filter_output(master, resp, decimate, out_type);
  if (!slave)
    return -1;
  if (!response_in) {
    if (set_filter(slave, low, high, beta) != 0)
      return -2;
  }
  if (response_out) /* set_filter always designs all N_dec bins (filter.c:523); a caller-supplied REAL response has N_dec/2+1 */
    memcpy(response_out, slave->response, sizeof(complex float) * (response_in ? rbins : N_dec));
  if (noise_gain_out)
    *noise_gain_out = slave->noise_gain;
  int const olen = slave->olen;
  for (int b = 0; b < nblocks; b++) {
    if (in_type == REAL)
      memcpy(master->input.r, in + (size_t)b * L, sizeof(float) * L);
    else
      memcpy(master->input.c, in + 2 * (size_t)b * L, sizeof(complex float) * L);
    execute_filter_input(master);
    execute_filter_output(slave);
    if (out_type == REAL)
      memcpy(out + (size_t)b * olen, slave->output.r, sizeof(float) * olen);
    else
      memcpy(out + 2 * (size_t)b * olen, slave->output.c, sizeof(complex float) * olen);
  }
  if (fdomain_out) {
    int const fb = (in_type == REAL) ? N / 2 + 1 : N;
    memcpy(fdomain_out, master->fdomain, sizeof(complex float) * fb);
  }
  delete_filter_output(slave);
  delete_filter_input(master);
  return olen;
}

/* ---- oscillator (osc.c:22-59) ---- */
/* Runs set_osc(f, r) on a zeroed struct osc, then nsteps of step_osc; out gets 2*nsteps doubles (re, im). */
int ref_osc_run(double f, double r, long nsteps, double *out) {
  struct osc o;
  memset(&o, 0, sizeof(o));
  pthread_mutex_init(&o.mutex, NULL);
  set_osc(&o, f, r);
  for (long i = 0; i < nsteps; i++) {
    complex double const v = step_osc(&o);
    out[2 * i] = creal(v);
    out[2 * i + 1] = cimag(v);
  }
  return 0;
}

/* ---- half-band decimators (decimate.c:44/111, :88/148), state carried across calls by the caller ---- */
int ref_hb15(float *state16 /* coeffs[4], even[4], odd[4], old_odd[4] */, float *out, float *in, int cnt) {
  struct hb15_state st;
  memcpy(&st, state16, sizeof(st));
  hb15_block(&st, out, in, cnt);
  memcpy(state16, &st, sizeof(st));
  return 0;
}
int ref_hb3(float *state, float *out, float *in, int cnt) {
  hb3_block(state, out, in, cnt);
  return 0;
}

/* ---- misc direct wrappers ---- */
int ref_make_kaiser(float *w, unsigned M, float beta) { return make_kaiser(w, M, beta); }
int ref_window_filter(int L, int M, float *response, float beta) { return window_filter(L, M, (complex float *)response, beta); }
int ref_window_rfilter(int L, int M, float *response, float beta) { return window_rfilter(L, M, (complex float *)response, beta); }
const char *ref_build_info(void) {
  return "verbatim reference sources from /root/reference, flags per reference Makefile:2 "
         "(-O3 -march=native -std=gnu11 -pthread -funsafe-math-optimizations -DNDEBUG=1)";
}

/* ---- wire-format glue: the reference's own functions around thin restatements of their callers ---- */
#include <fcntl.h>
#include <sys/socket.h>
#include "multicast.h"

/* One datagram through rtp_recv's parsing (main.c:313-344, real ntoh_rtp) and the packet head of proc_samples
 * (radio.c:60-100, real rtp_process). Appends raw samples (4 or 2 bytes each, by payload type), zero bytes for lost
 * samples. Returns complex samples appended, -1 if the datagram is ignored or dropped. */
long long ref_glue_ingest(struct rtp_state *state, long long *samples, const unsigned char *datagram, int size,
                          unsigned char *dst) {
  if (size < RTP_MIN_SIZE) return -1;
  unsigned char *content = malloc(size + 64);
  memcpy(content, datagram, size);
  struct rtp_header rtp;
  memset(&rtp, 0, sizeof(rtp));
  unsigned char *dp = ntoh_rtp(&rtp, content);
  size -= (dp - content);
  if (rtp.pad) {
    size -= dp[size - 1];
    rtp.pad = 0;
  }
  if (rtp.type != IQ_PT && rtp.type != IQ_PT8) {
    free(content);
    return -1;
  }
  dp += 24;
  size -= 24;
  const int bps = rtp.type == IQ_PT ? 4 : 2;
  const int sampcount = size / bps;
  if (rtp.ssrc != state->ssrc) *samples = 0;
  const int time_step = rtp_process(state, &rtp, sampcount);
  if (time_step < 0 || time_step > 192000) {
    free(content);
    return -1;
  }
  if (time_step > 0) {
    *samples += time_step;
    memset(dst, 0, (size_t)time_step * bps);
    dst += (size_t)time_step * bps;
  }
  *samples += sampcount;
  memcpy(dst, dp, (size_t)sampcount * bps);
  free(content);
  return (long long)time_step + sampcount;
}

/* The real send_mono_output / send_stereo_output (audio.c, compiled with renamed entry points: this harness owns the
 * original names to capture PCM) writing into a UNIX datagram socket pair; the packets are read back verbatim.
 * state5: ssrc, timestamp, seq, silent, (out) packets. Returns the number of packets, lens[i] their sizes,
 * out = packets back to back. Keep calls short (a few packets): the socket queue is small. */
/* Toolchain quirk of this image: with /opt/gcc, <complex.h> (pulled in by radio.h) resolves to libstdc++'s C wrapper whose
 * c++config.h #undefs min/max, so audio.c's min() macro (misc.h:12) is gone by the time audio.c:44,:95 use it and the
 * compiler emits a call to a function `min`. Both call sites pass ints. */
#undef min
int min(int a, int b) { return a < b ? a : b; }
int refaudio_send_mono_output(struct demod *, const float *, int);
int refaudio_send_stereo_output(struct demod *, const float *, int);
int ref_glue_send(int channels, long long *state5, const float *buf, int frames, unsigned char *out, int out_cap,
                  int *lens, int max_packets) {
  int sv[2];
  if (socketpair(AF_UNIX, SOCK_DGRAM, 0, sv) != 0) return -1;
  fcntl(sv[1], F_SETFL, O_NONBLOCK);
  struct demod *demod = calloc(1, sizeof(*demod));
  demod->output.fd = sv[0];
  demod->output.rtp.ssrc = (uint32_t)state5[0];
  demod->output.rtp.timestamp = (uint32_t)state5[1];
  demod->output.rtp.seq = (uint16_t)state5[2];
  demod->output.silent = (int)state5[3];
  if (channels == 2)
    refaudio_send_stereo_output(demod, buf, frames);
  else
    refaudio_send_mono_output(demod, buf, frames);
  int n = 0, used = 0;
  while (n < max_packets) {
    const int r = recv(sv[1], out + used, out_cap - used, 0);
    if (r <= 0) break;
    lens[n++] = r;
    used += r;
  }
  state5[1] = demod->output.rtp.timestamp;
  state5[2] = demod->output.rtp.seq;
  state5[3] = demod->output.silent;
  state5[4] = demod->output.rtp.packets;
  close(sv[0]);
  close(sv[1]);
  free(demod);
  return n;
}

/* status TLV: the reference's own encoders (status.c) in the order radio_status.c:171-205 uses them, for
 * the fields the product computes. types[] carries the enum status_type values (the caller reads them from status.h). */
#include "status.h"
int ref_glue_status(int demod_type, int isb, float noise_bw, float if_power, float bb_power, float gain, float pdev,
                    float foffset, float snr, int channels, unsigned char *out) {
  unsigned char *bp = out;
  encode_float(&bp, NOISE_BANDWIDTH, noise_bw);
  encode_float(&bp, IF_POWER, if_power);
  encode_float(&bp, BASEBAND_POWER, bb_power);
  encode_byte(&bp, DEMOD_MODE, demod_type);
  switch (demod_type) {
  case AM_DEMOD:
    encode_float(&bp, DEMOD_GAIN, gain);
    break;
  case FM_DEMOD:
    encode_float(&bp, PEAK_DEVIATION, pdev);
    encode_float(&bp, FREQ_OFFSET, foffset);
    encode_float(&bp, DEMOD_SNR, snr);
    break;
  default:
    encode_float(&bp, DEMOD_GAIN, gain);
    encode_int32(&bp, INDEPENDENT_SIDEBAND, isb);
    break;
  }
  encode_int32(&bp, OUTPUT_CHANNELS, channels);
  encode_eol(&bp);
  return (int)(bp - out);
}
