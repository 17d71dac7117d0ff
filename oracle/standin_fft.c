/* Stand-in CPU FFT for the oracle (TEST INFRASTRUCTURE ONLY — never linked into the product).
 * See standin_fft.h. Double-precision Stockham autosort, decimation in frequency:
 *   y[q + s*(R*p + j)] = w_n^(p*j) * sum_r x[q + s*(p + m*r)] * W_R^(r*j),  m = n/R,
 * then n <- n/R, s <- s*R, buffers swapped. After the last pass the data is in natural order.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "standin_fft.h"

#define MAXFACT 64

struct sfft_plan {
  int n;
  int nfact;
  int fact[MAXFACT];
  double complex *tw;  /* tw[k] = exp(-2*pi*i*k/n), k in [0,n) */
};

static void make_twiddles(double complex *tw, int n) {
  /* Use octant symmetry so every entry comes from sincos of an argument in [0, pi/4]. */
  for (int k = 0; k < n; k++) {
    /* angle = 2*pi*k/n; reduce k to first octant */
    long long k8 = (long long)k * 8;
    int oct = (int)(k8 / n);            /* 0..7 */
    long long rem = k8 - (long long)oct * n; /* in [0,n): angle within octant = (pi/4)*rem/n */
    double c, s;
    double a;
    int odd = oct & 1;
    if (!odd)
      a = (M_PI / 4.0) * (double)rem / (double)n;
    else
      a = (M_PI / 4.0) * (double)(n - rem) / (double)n;
    s = sin(a);
    c = cos(a);
    double cr, si; /* cos, sin of the full angle */
    switch (oct) {
    default:
    case 0: cr = c;  si = s;  break;
    case 1: cr = s;  si = c;  break;
    case 2: cr = -s; si = c;  break;
    case 3: cr = -c; si = s;  break;
    case 4: cr = -c; si = -s; break;
    case 5: cr = -s; si = -c; break;
    case 6: cr = s;  si = -c; break;
    case 7: cr = c;  si = -s; break;
    }
    tw[k] = CMPLX(cr, -si); /* exp(-i*angle) */
  }
}

sfft_plan *sfft_create(int n) {
  if (n < 1)
    return NULL;
  sfft_plan *p = calloc(1, sizeof(*p));
  if (!p)
    return NULL;
  p->n = n;
  int m = n;
  while (m % 4 == 0) { p->fact[p->nfact++] = 4; m /= 4; }
  while (m % 2 == 0) { p->fact[p->nfact++] = 2; m /= 2; }
  while (m % 3 == 0) { p->fact[p->nfact++] = 3; m /= 3; }
  while (m % 5 == 0) { p->fact[p->nfact++] = 5; m /= 5; }
  for (int f = 7; m > 1; f += 2) {
    while (m % f == 0) {
      if (p->nfact == MAXFACT) { free(p); return NULL; }
      p->fact[p->nfact++] = f;
      m /= f;
    }
  }
  p->tw = malloc(sizeof(double complex) * (size_t)n);
  if (!p->tw) { free(p); return NULL; }
  make_twiddles(p->tw, n);
  return p;
}

void sfft_destroy(sfft_plan *p) {
  if (!p)
    return;
  free(p->tw);
  free(p);
}

int sfft_size(const sfft_plan *p) { return p ? p->n : 0; }

/* twiddle lookup with sign: W(k) = exp(sign*2*pi*i*k/N), k taken mod N */
static inline double complex TW(const sfft_plan *p, long long k, int sign) {
  double complex w = p->tw[k % p->n];
  return sign < 0 ? w : conj(w);
}

static void pass_generic(const sfft_plan *pl, int R, int n, int s, const double complex *x, double complex *y, int sign) {
  int const m = n / R;
  int const N = pl->n;
  long long const step = N / n;  /* w_n = W_N^step */
  long long const rstep = N / R; /* W_R = W_N^rstep */
  double complex v[R], o[R];
  for (int p = 0; p < m; p++) {
    for (int q = 0; q < s; q++) {
      for (int r = 0; r < R; r++)
        v[r] = x[q + (size_t)s * (p + (size_t)m * r)];
      for (int j = 0; j < R; j++) {
        double complex acc = 0;
        for (int r = 0; r < R; r++)
          acc += v[r] * TW(pl, rstep * (((long long)r * j) % R), sign);
        o[j] = acc;
      }
      for (int j = 0; j < R; j++)
        y[q + (size_t)s * ((size_t)R * p + j)] = o[j] * TW(pl, step * (long long)p * j, sign);
    }
  }
}

static void pass2(const sfft_plan *pl, int n, int s, const double complex *x, double complex *y, int sign) {
  int const m = n / 2;
  long long const step = pl->n / n;
  for (int p = 0; p < m; p++) {
    double complex const w1 = TW(pl, step * p, sign);
    for (int q = 0; q < s; q++) {
      double complex const a = x[q + (size_t)s * p];
      double complex const b = x[q + (size_t)s * (p + (size_t)m)];
      y[q + (size_t)s * (2 * (size_t)p)] = a + b;
      y[q + (size_t)s * (2 * (size_t)p + 1)] = (a - b) * w1;
    }
  }
}

static inline double complex mulj(double complex a, int sign) {
  /* multiply by exp(sign*i*pi/2) = sign*i */
  return sign < 0 ? CMPLX(cimag(a), -creal(a)) : CMPLX(-cimag(a), creal(a));
}

static void pass4(const sfft_plan *pl, int n, int s, const double complex *x, double complex *y, int sign) {
  int const m = n / 4;
  long long const step = pl->n / n;
  for (int p = 0; p < m; p++) {
    double complex const w1 = TW(pl, step * p, sign);
    double complex const w2 = TW(pl, step * 2 * p, sign);
    double complex const w3 = TW(pl, step * 3 * p, sign);
    for (int q = 0; q < s; q++) {
      double complex const a = x[q + (size_t)s * p];
      double complex const b = x[q + (size_t)s * (p + (size_t)m)];
      double complex const c = x[q + (size_t)s * (p + 2 * (size_t)m)];
      double complex const d = x[q + (size_t)s * (p + 3 * (size_t)m)];
      double complex const apc = a + c, amc = a - c, bpd = b + d;
      double complex const jbmd = mulj(b - d, sign);
      y[q + (size_t)s * (4 * (size_t)p)] = apc + bpd;
      y[q + (size_t)s * (4 * (size_t)p + 1)] = (amc + jbmd) * w1;
      y[q + (size_t)s * (4 * (size_t)p + 2)] = (apc - bpd) * w2;
      y[q + (size_t)s * (4 * (size_t)p + 3)] = (amc - jbmd) * w3;
    }
  }
}

void sfft_exec(const sfft_plan *pl, const double complex *in, double complex *out, int sign) {
  int const N = pl->n;
  if (N == 1) {
    out[0] = in[0];
    return;
  }
  double complex *bufa = malloc(sizeof(double complex) * (size_t)N);
  double complex *bufb = malloc(sizeof(double complex) * (size_t)N);
  memcpy(bufa, in, sizeof(double complex) * (size_t)N);
  double complex *x = bufa, *y = bufb;
  int n = N, s = 1;
  for (int f = 0; f < pl->nfact; f++) {
    int const R = pl->fact[f];
    switch (R) {
    case 2: pass2(pl, n, s, x, y, sign); break;
    case 4: pass4(pl, n, s, x, y, sign); break;
    default: pass_generic(pl, R, n, s, x, y, sign); break;
    }
    n /= R;
    s *= R;
    double complex *t = x; x = y; y = t;
  }
  memcpy(out, x, sizeof(double complex) * (size_t)N);
  free(bufa);
  free(bufb);
}
