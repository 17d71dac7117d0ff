/* Minimal FFTW3 single-precision API shim (test infrastructure, not product code).
 *
 * FFTW3 is a third-party dependency of the reference (link line: reference Makefile:65,
 * version unpinned per INSTALLING.md:13) and is NOT installed in this image. This header
 * declares only the public-API subset the reference's hot-path files call (41 call sites in
 * filter.c, 9 in fm.c, 8 in linear.c, 4 in main.c) so those files compile UNMODIFIED; the
 * implementation is oracle/fftw_shim.c. Transform definitions follow the published FFTW
 * manual: forward = sum x[n] exp(-2 pi i k n / N), backward = exp(+...), both unnormalised;
 * r2c writes N/2+1 bins; c2r assumes Hermitian input and may destroy it.
 */
#ifndef KA9Q_ORACLE_FFTW3_SHIM_H
#define KA9Q_ORACLE_FFTW3_SHIM_H 1
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(_Complex_I) && defined(complex) && defined(I)
typedef float _Complex fftwf_complex;
#else
typedef float fftwf_complex[2];
#endif

typedef struct ka9q_shim_plan *fftwf_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_complex *fftwf_alloc_complex(size_t n);
float *fftwf_alloc_real(size_t n);
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

int fftwf_import_system_wisdom(void);
void fftwf_make_planner_thread_safe(void);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int nthreads);

#ifdef __cplusplus
}
#endif
#endif
