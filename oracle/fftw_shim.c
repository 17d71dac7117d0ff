/* fftwf_* API implemented for the oracle (TEST INFRASTRUCTURE ONLY — never linked into the product).
 *
 * The reference calls FFTW3 single precision (reference Makefile:65; INSTALLING.md:13, unpinned).
 * FFTW is not installed in this image, so the verbatim reference objects are linked against this
 * shim instead. Three backends, chosen by ka9q_oracle_set_fft_backend() or $KA9Q_ORACLE_FFT:
 *   "standin" (default)  double-precision mixed-radix FFT (standin_fft.c), result rounded to float.
 *                        Used for parity: it is *more* accurate than fp32 FFTW, so it is a
 *                        stricter reference for the 1e-5 relative-RMS criterion.
 *   "mkl"                Intel MKL DFTI single precision, dlopen'ed from torch's libtorch_cpu.so
 *                        (path in $KA9Q_ORACLE_MKL_LIB). Used for the CPU *timing* baseline: a
 *                        production-quality fp32 FFT comparable to FFTW.
 *   "fftw"               a system libfftw3f.so.3 if dlopen finds one (the true FFTW path).
 *   "auto"               fftw, else mkl, else standin.
 * Transform conventions (FFTW manual): unnormalised; forward sign -1; r2c writes n/2+1 bins;
 * c2r takes n/2+1 Hermitian bins (imaginary parts of DC and Nyquist ignored).
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shim/fftw3.h"
#include "standin_fft.h"

enum kind { K_C2C, K_R2C, K_C2R };
enum backend { B_UNSET = 0, B_STANDIN, B_MKL, B_FFTW };

struct ka9q_shim_plan {
  enum kind kind;
  enum backend backend;
  int n;
  int sign;
  void *in;
  void *out;
  sfft_plan *sp;        /* standin */
  void *mkl;            /* DFTI descriptor */
  void *fftw;           /* real fftwf_plan */
};

/* ---------------- backend selection ---------------- */
static pthread_mutex_t Backend_mutex = PTHREAD_MUTEX_INITIALIZER;
static enum backend Backend = B_UNSET;

/* MKL DFTI entry points (public MKL API; enum values from mkl_dfti.h) */
enum {
  DFTI_FORWARD_SCALE = 4, DFTI_BACKWARD_SCALE = 5, DFTI_CONJUGATE_EVEN_STORAGE = 10, DFTI_PLACEMENT = 11,
  DFTI_PACKED_FORMAT = 21, DFTI_THREAD_LIMIT = 27,
  DFTI_COMPLEX = 32, DFTI_REAL = 33, DFTI_SINGLE = 35, DFTI_COMPLEX_COMPLEX = 39, DFTI_NOT_INPLACE = 44,
  DFTI_CCE_FORMAT = 57
};
static long (*p_DftiCreateDescriptor_s_1d)(void **, int, long);
static long (*p_DftiSetValue)(void *, int, ...);
static long (*p_DftiCommitDescriptor)(void *);
static long (*p_DftiComputeForward)(void *, void *, ...);
static long (*p_DftiComputeBackward)(void *, void *, ...);
static long (*p_DftiFreeDescriptor)(void **);

/* real FFTW entry points */
static void *(*p_fftw_plan_dft_1d)(int, void *, void *, int, unsigned);
static void *(*p_fftw_plan_dft_r2c_1d)(int, float *, void *, unsigned);
static void *(*p_fftw_plan_dft_c2r_1d)(int, void *, float *, unsigned);
static void (*p_fftw_execute)(void *);
static void (*p_fftw_destroy_plan)(void *);
static void (*p_fftw_make_planner_thread_safe)(void);

static int load_mkl(void) {
  static int tried = 0, ok = 0;
  if (tried)
    return ok;
  tried = 1;
  const char *path = getenv("KA9Q_ORACLE_MKL_LIB");
  void *h = NULL;
  if (path && *path)
    h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h)
    h = dlopen("libtorch_cpu.so", RTLD_NOW | RTLD_LOCAL);
  if (!h)
    h = dlopen("libmkl_rt.so", RTLD_NOW | RTLD_LOCAL);
  if (!h)
    return 0;
  p_DftiCreateDescriptor_s_1d = dlsym(h, "DftiCreateDescriptor_s_1d");
  p_DftiSetValue = dlsym(h, "DftiSetValue");
  p_DftiCommitDescriptor = dlsym(h, "DftiCommitDescriptor");
  p_DftiComputeForward = dlsym(h, "DftiComputeForward");
  p_DftiComputeBackward = dlsym(h, "DftiComputeBackward");
  p_DftiFreeDescriptor = dlsym(h, "DftiFreeDescriptor");
  ok = p_DftiCreateDescriptor_s_1d && p_DftiSetValue && p_DftiCommitDescriptor && p_DftiComputeForward &&
       p_DftiComputeBackward && p_DftiFreeDescriptor;
  return ok;
}

static int load_fftw(void) {
  static int tried = 0, ok = 0;
  if (tried)
    return ok;
  tried = 1;
  void *h = dlopen("libfftw3f.so.3", RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
  if (!h)
    return 0;
  p_fftw_plan_dft_1d = dlsym(h, "fftwf_plan_dft_1d");
  p_fftw_plan_dft_r2c_1d = dlsym(h, "fftwf_plan_dft_r2c_1d");
  p_fftw_plan_dft_c2r_1d = dlsym(h, "fftwf_plan_dft_c2r_1d");
  p_fftw_execute = dlsym(h, "fftwf_execute");
  p_fftw_destroy_plan = dlsym(h, "fftwf_destroy_plan");
  p_fftw_make_planner_thread_safe = dlsym(h, "fftwf_make_planner_thread_safe");
  ok = p_fftw_plan_dft_1d && p_fftw_plan_dft_r2c_1d && p_fftw_plan_dft_c2r_1d && p_fftw_execute && p_fftw_destroy_plan;
  if (ok && p_fftw_make_planner_thread_safe)
    p_fftw_make_planner_thread_safe();
  return ok;
}

/* Returns 0 on success, -1 if the requested backend is unavailable (backend left unchanged). */
int ka9q_oracle_set_fft_backend(const char *name) {
  int r = 0;
  pthread_mutex_lock(&Backend_mutex);
  if (!name || !*name || strcmp(name, "standin") == 0) {
    Backend = B_STANDIN;
  } else if (strcmp(name, "mkl") == 0) {
    if (load_mkl()) Backend = B_MKL; else r = -1;
  } else if (strcmp(name, "fftw") == 0) {
    if (load_fftw()) Backend = B_FFTW; else r = -1;
  } else if (strcmp(name, "auto") == 0) {
    if (load_fftw()) Backend = B_FFTW;
    else if (load_mkl()) Backend = B_MKL;
    else Backend = B_STANDIN;
  } else {
    r = -1;
  }
  pthread_mutex_unlock(&Backend_mutex);
  return r;
}

static enum backend current_backend(void) {
  if (Backend == B_UNSET) {
    const char *e = getenv("KA9Q_ORACLE_FFT");
    if (ka9q_oracle_set_fft_backend(e) != 0)
      ka9q_oracle_set_fft_backend("standin");
  }
  return Backend;
}

const char *ka9q_oracle_fft_backend(void) {
  switch (current_backend()) {
  case B_MKL: return "mkl-dfti-fp32";
  case B_FFTW: return "fftw3f";
  default: return "standin-fp64-mixed-radix";
  }
}

/* ---------------- allocation ---------------- */
void *fftwf_malloc(size_t n) {
  void *p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64) != 0)
    return NULL;
  /* zero-filled: linear.c:90-91 transforms its 64K-sample carrier-search buffer before it has been filled once
     (fft_samples > fftsize/2), i.e. it reads memory FFTW's allocator never initialised; zeros make the oracle deterministic */
  memset(p, 0, n ? n : 64);
  return p;
}
fftwf_complex *fftwf_alloc_complex(size_t n) { return fftwf_malloc(n * 2 * sizeof(float)); }
float *fftwf_alloc_real(size_t n) { return fftwf_malloc(n * sizeof(float)); }
void fftwf_free(void *p) { free(p); }

/* ---------------- planning ---------------- */
static fftwf_plan make_plan(enum kind kind, int n, void *in, void *out, int sign) {
  struct ka9q_shim_plan *p = calloc(1, sizeof(*p));
  if (!p)
    return NULL;
  p->kind = kind;
  p->n = n;
  p->sign = sign;
  p->in = in;
  p->out = out;
  p->backend = current_backend();
  switch (p->backend) {
  case B_FFTW:
    if (kind == K_C2C) p->fftw = p_fftw_plan_dft_1d(n, in, out, sign, FFTW_ESTIMATE);
    else if (kind == K_R2C) p->fftw = p_fftw_plan_dft_r2c_1d(n, in, out, FFTW_ESTIMATE);
    else p->fftw = p_fftw_plan_dft_c2r_1d(n, in, out, FFTW_ESTIMATE);
    if (p->fftw)
      return p;
    p->backend = B_STANDIN;
    break;
  case B_MKL: {
    long st = p_DftiCreateDescriptor_s_1d(&p->mkl, kind == K_C2C ? DFTI_COMPLEX : DFTI_REAL, (long)n);
    if (st == 0 && in != out) st = p_DftiSetValue(p->mkl, DFTI_PLACEMENT, DFTI_NOT_INPLACE);
    if (st == 0 && kind != K_C2C) st = p_DftiSetValue(p->mkl, DFTI_CONJUGATE_EVEN_STORAGE, DFTI_COMPLEX_COMPLEX);
    if (st == 0 && kind != K_C2C) st = p_DftiSetValue(p->mkl, DFTI_PACKED_FORMAT, DFTI_CCE_FORMAT);
    if (st == 0) st = p_DftiSetValue(p->mkl, DFTI_THREAD_LIMIT, 1);
    if (st == 0) st = p_DftiCommitDescriptor(p->mkl);
    if (st == 0)
      return p;
    if (p->mkl) p_DftiFreeDescriptor(&p->mkl);
    p->mkl = NULL;
    p->backend = B_STANDIN;
    break;
  }
  default:
    break;
  }
  p->sp = sfft_create(n);
  if (!p->sp) {
    free(p);
    return NULL;
  }
  return p;
}

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
  (void)flags;
  return make_plan(K_C2C, n, in, out, sign);
}
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags) {
  (void)flags;
  return make_plan(K_R2C, n, in, out, FFTW_FORWARD);
}
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags) {
  (void)flags;
  return make_plan(K_C2R, n, in, out, FFTW_BACKWARD);
}

void fftwf_destroy_plan(fftwf_plan p) {
  if (!p)
    return;
  if (p->sp) sfft_destroy(p->sp);
  if (p->mkl) p_DftiFreeDescriptor(&p->mkl);
  if (p->fftw) p_fftw_destroy_plan(p->fftw);
  free(p);
}

/* ---------------- execution ---------------- */
static void exec_standin(const struct ka9q_shim_plan *p) {
  int const n = p->n;
  double complex *a = malloc(sizeof(double complex) * (size_t)n);
  double complex *b = malloc(sizeof(double complex) * (size_t)n);
  switch (p->kind) {
  case K_C2C: {
    float const *in = p->in;
    float *out = p->out;
    for (int i = 0; i < n; i++)
      a[i] = CMPLX((double)in[2 * i], (double)in[2 * i + 1]);
    sfft_exec(p->sp, a, b, p->sign);
    for (int i = 0; i < n; i++) {
      out[2 * i] = (float)creal(b[i]);
      out[2 * i + 1] = (float)cimag(b[i]);
    }
    break;
  }
  case K_R2C: {
    float const *in = p->in;
    float *out = p->out;
    for (int i = 0; i < n; i++)
      a[i] = (double)in[i];
    sfft_exec(p->sp, a, b, -1);
    for (int i = 0; i <= n / 2; i++) {
      out[2 * i] = (float)creal(b[i]);
      out[2 * i + 1] = (float)cimag(b[i]);
    }
    break;
  }
  case K_C2R: {
    float const *in = p->in;
    float *out = p->out;
    /* Hermitian extension; imaginary parts of DC (and Nyquist for even n) are ignored */
    a[0] = (double)in[0];
    for (int i = 1; i <= n / 2; i++) {
      double complex v = CMPLX((double)in[2 * i], (double)in[2 * i + 1]);
      if ((n % 2 == 0) && i == n / 2) {
        a[i] = creal(v);
      } else {
        a[i] = v;
        a[n - i] = conj(v);
      }
    }
    sfft_exec(p->sp, a, b, +1);
    for (int i = 0; i < n; i++)
      out[i] = (float)creal(b[i]);
    break;
  }
  }
  free(a);
  free(b);
}

void fftwf_execute(const fftwf_plan p) {
  if (!p)
    return;
  switch (p->backend) {
  case B_FFTW:
    p_fftw_execute(p->fftw);
    return;
  case B_MKL:
    if (p->in == p->out) {
      if (p->kind == K_C2C && p->sign > 0) p_DftiComputeBackward(p->mkl, p->in);
      else if (p->kind == K_C2R) p_DftiComputeBackward(p->mkl, p->in);
      else p_DftiComputeForward(p->mkl, p->in);
    } else {
      if (p->kind == K_C2C && p->sign > 0) p_DftiComputeBackward(p->mkl, p->in, p->out);
      else if (p->kind == K_C2R) p_DftiComputeBackward(p->mkl, p->in, p->out);
      else p_DftiComputeForward(p->mkl, p->in, p->out);
    }
    return;
  default:
    exec_standin(p);
    return;
  }
}

int fftwf_import_system_wisdom(void) { return 0; }
void fftwf_make_planner_thread_safe(void) {}
int fftwf_init_threads(void) { return 1; }
void fftwf_plan_with_nthreads(int nthreads) { (void)nthreads; }
