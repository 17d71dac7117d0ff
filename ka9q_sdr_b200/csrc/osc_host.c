/* Drop-in oscillator (include/ka9q_b200.h section A; reference osc.h:9-24, osc.c:14-59).
 *
 * Host-side scalar state by nature: one complex-double recurrence step per call, used by callers that keep a
 * time-domain NCO (packet.c, modulate.c, the PLL in linear.c). The channelizer itself does NOT use it — it replaces
 * the per-sample second-LO multiply (radio.c:132) by a bin rotation plus a per-block phase and
 * computes the post-detection shift phasor in closed form on the device (chan_kernels.cu, linear_kernel).
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include "../../include/ka9q_b200.h"

enum { RENORM_INTERVAL = 16384 }; /* osc.c:11 */

static inline double complex unit_phasor(double cycles) {
  /* cos + j sin of 2*pi*cycles, evaluated as sincos(x*M_PI) like dsp.h:49 / dsp.c:36-40 */
  double s, c;
  sincos(2 * cycles * M_PI, &s, &c);
  return CMPLX(c, s);
}

int is_phasor_init(const complex double x) {
  double const n = creal(x) * creal(x) + cimag(x) * cimag(x);
  return !(isnan(creal(x)) || isnan(cimag(x)) || n < 0.9);
}

void set_osc(struct osc *o, double f, double r) {
  pthread_mutex_lock(&o->mutex);
  if (!is_phasor_init(o->phasor)) { /* keep phase continuity across retunes (osc.c:24-27) */
    o->phasor = 1;
    o->steps = 0;
  }
  o->freq = f;
  o->rate = r;
  o->phasor_step = unit_phasor(o->freq);
  o->phasor_step_step = (o->rate != 0) ? unit_phasor(o->rate) : 1;
  pthread_mutex_unlock(&o->mutex);
}

complex double step_osc(struct osc *o) {
  complex double const out = o->phasor;
  if (o->freq != 0) {
    o->phasor *= o->phasor_step;
    if (o->rate != 0)
      o->phasor_step *= o->phasor_step_step;
  }
  if (++o->steps == RENORM_INTERVAL)
    renorm_osc(o);
  return out;
}

void renorm_osc(struct osc *o) {
  o->steps = 0;
  o->phasor /= cabs(o->phasor);
  if (o->rate != 0)
    o->phasor_step /= cabs(o->phasor_step);
}

/* Test/diagnostic helper: set_osc(f, r) on a fresh oscillator, then n step_osc() results as (re, im) doubles. */
int ka9q_osc_run(double f, double r, long n, double *out) {
  struct osc o = {0};
  if (!out || n < 0)
    return -1;
  pthread_mutex_init(&o.mutex, NULL);
  set_osc(&o, f, r);
  for (long i = 0; i < n; i++) {
    complex double const v = step_osc(&o);
    out[2 * i] = creal(v);
    out[2 * i + 1] = cimag(v);
  }
  pthread_mutex_destroy(&o.mutex);
  return 0;
}
