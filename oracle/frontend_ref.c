/* TEST INFRASTRUCTURE ONLY: CPU restatement of the sample path of the reference's `hackrf` daemon, which cannot be
 * compiled here (hackrf.c needs libhackrf / libusb headers). Every step cites the line of /root/reference/hackrf.c it
 * follows; the half-band decimator itself is the VERBATIM reference (decimate.c, linked into this library).
 *
 *   rx_callback   hackrf.c:129-196   int8 ingest, clip count, DC removal, I/Q gain + phase correction, estimate updates
 *   process       hackrf.c:198-345   Fs/4 rotation, hb15 cascade on both planes, x Filter_atten, (short)round(32767 s)
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "decimate.h"

struct fe_ref {
  int adc_samprate, log_decimate, offset, callback_samples;
  float dc_alpha, power_alpha;
  /* HackCD.* (hackrf.c:36-58), zero-initialised global (hackrf.c:81) */
  complex float DC;
  float sinphi, imbalance, in_power;
  long long clips;
  /* hackrf.c:121-124 */
  float gain_q, gain_i, secphi, tanphi;
  int rotate_phase; /* hackrf.c:208 */
  struct hb15_state st_real[8], st_imag[8];
};

struct fe_ref *ref_frontend_create(int out_samprate, int decimate, int offset, int callback_samples, float dc_alpha,
                                   float power_alpha) {
  struct fe_ref *f = calloc(1, sizeof(*f));
  f->adc_samprate = decimate * out_samprate;                 /* hackrf.c:463 */
  f->log_decimate = (int)round(log2(decimate));              /* hackrf.c:464 */
  f->offset = offset;
  f->callback_samples = callback_samples;
  f->dc_alpha = dc_alpha;
  f->power_alpha = power_alpha;
  f->gain_q = f->gain_i = f->secphi = 1;
  f->tanphi = 0;
  for (int i = 0; i < f->log_decimate; i++) {                /* hackrf.c:229-237 */
    f->st_real[i].coeffs[3] = f->st_imag[i].coeffs[3] = 490. / 802;
    f->st_real[i].coeffs[2] = f->st_imag[i].coeffs[2] = -116. / 802;
    f->st_real[i].coeffs[1] = f->st_imag[i].coeffs[1] = 33. / 802;
    f->st_real[i].coeffs[0] = f->st_imag[i].coeffs[0] = -6. / 802;
  }
  return f;
}
void ref_frontend_destroy(struct fe_ref *f) { free(f); }

void ref_frontend_set_estimates(struct fe_ref *f, float dc_i, float dc_q, float imbalance, float sinphi) {
  f->DC = CMPLXF(dc_i, dc_q);
  f->imbalance = imbalance;
  f->sinphi = sinphi;
  f->gain_q = sqrtf(0.5 * (1 + f->imbalance));              /* hackrf.c:190-193 */
  f->gain_i = sqrtf(0.5 * (1 + 1. / f->imbalance));
  f->secphi = 1 / sqrtf(1 - f->sinphi * f->sinphi);
  f->tanphi = f->sinphi * f->secphi;
}

void ref_frontend_status(const struct fe_ref *f, float *out7) {
  out7[0] = crealf(f->DC);
  out7[1] = cimagf(f->DC);
  out7[2] = f->imbalance;
  out7[3] = f->sinphi;
  out7[4] = f->in_power;
  out7[5] = (float)f->clips;
}

/* nsamples complex int8 samples (a whole number of callback blocks) -> nsamples / decimate int16 I/Q pairs */
int ref_frontend_process(struct fe_ref *f, const signed char *iq8, long nsamples, short *out) {
  float const SCALE8 = 1. / 127; /* radio.c:39 / hackrf.c:155 */
  int const Decimate = 1 << f->log_decimate;
  if (nsamples % f->callback_samples || nsamples % Decimate) return -1;
  float *wr = malloc(sizeof(float) * nsamples), *wi = malloc(sizeof(float) * nsamples);
  float rate_factor = 1. / (f->adc_samprate * f->power_alpha); /* hackrf.c:138 */
  long pos = 0;
  for (long blk = 0; blk < nsamples / f->callback_samples; blk++) {
    int const samples = f->callback_samples;
    const signed char *dp = iq8 + 2 * blk * (long)samples;
    complex float samp_sum = 0;
    float i_energy = 0, q_energy = 0, dotprod = 0;
    for (int n = 0; n < samples; n++) { /* hackrf.c:140-178 */
      int isamp_i = *dp++;
      int isamp_q = *dp++;
      if (isamp_q == -128) { f->clips++; isamp_q = -127; }
      if (isamp_i == -128) { f->clips++; isamp_i = -127; }
      complex float samp = CMPLXF(isamp_i, isamp_q) * SCALE8;
      samp_sum += samp;
      samp -= f->DC;
      i_energy += crealf(samp) * crealf(samp);
      q_energy += cimagf(samp) * cimagf(samp);
      __real__ samp *= f->gain_i;
      __imag__ samp *= f->gain_q;
      dotprod += crealf(samp) * cimagf(samp);
      __imag__ samp = f->secphi * cimagf(samp) - f->tanphi * crealf(samp);
      /* process(), hackrf.c:264-291: rotation as the sample leaves the buffer */
      float samp_i = crealf(samp), samp_q = cimagf(samp);
      switch (f->rotate_phase) {
      default:
      case 0: wr[pos] = samp_i;  wi[pos] = samp_q;  break;
      case 1: wr[pos] = -samp_q; wi[pos] = samp_i;  break;
      case 2: wr[pos] = -samp_i; wi[pos] = -samp_q; break;
      case 3: wr[pos] = samp_q;  wi[pos] = -samp_i; break;
      }
      pos++;
      f->rotate_phase += f->offset;
      f->rotate_phase &= 3;
    }
    /* hackrf.c:182-194 */
    f->DC += f->dc_alpha * (samp_sum - samples * f->DC);
    float block_energy = 0.5 * (i_energy + q_energy);
    if (block_energy > 0) {
      f->in_power = block_energy / samples;
      f->imbalance += rate_factor * samples * ((i_energy / q_energy) - f->imbalance);
      float dpn = dotprod / block_energy;
      f->sinphi += rate_factor * samples * (dpn - f->sinphi);
      f->gain_q = sqrtf(0.5 * (1 + f->imbalance));
      f->gain_i = sqrtf(0.5 * (1 + 1. / f->imbalance));
      f->secphi = 1 / sqrtf(1 - f->sinphi * f->sinphi);
      f->tanphi = f->sinphi * f->secphi;
    }
  }
  /* hackrf.c:297-328: stage Log_decimate-1 first ... stage 0 last, in place, each plane with its own states (all stages are
     hb15: stage_threshold = 8 > Log_decimate, hackrf.c:76); the streaming filter gives the same samples for any chunking */
  long n = nsamples;
  for (int j = f->log_decimate - 1; j >= 0; j--) {
    hb15_block(&f->st_real[j], wr, wr, (int)(n / 2));
    hb15_block(&f->st_imag[j], wi, wi, (int)(n / 2));
    n /= 2;
  }
  float const Filter_atten = powf(.5, f->log_decimate); /* hackrf.c:469 */
  for (long j = 0; j < n; j++) {
    float s = wr[j] * Filter_atten;
    out[2 * j] = (short)round(32767 * s); /* hackrf.c:309 */
    s = wi[j] * Filter_atten;
    out[2 * j + 1] = (short)round(32767 * s); /* hackrf.c:327 */
  }
  free(wr);
  free(wi);
  return (int)n;
}
