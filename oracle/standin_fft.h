/* Stand-in CPU FFT for the oracle (TEST INFRASTRUCTURE ONLY — never linked into the product).
 *
 * Double-precision mixed-radix Stockham autosort FFT for any N (radices 4,2,3,5 native,
 * other prime factors by a direct O(R^2) butterfly). It stands in for FFTW3, which the
 * reference links (reference Makefile:65) but which is not installed in this image.
 * Definition (FFTW manual): out[k] = sum_n in[n] * exp(sign * 2*pi*i*k*n/N), unnormalised.
 */
#ifndef KA9Q_ORACLE_STANDIN_FFT_H
#define KA9Q_ORACLE_STANDIN_FFT_H 1
#include <complex.h>

typedef struct sfft_plan sfft_plan;

sfft_plan *sfft_create(int n);
void sfft_destroy(sfft_plan *p);
int sfft_size(const sfft_plan *p);
/* in and out must not alias; work is scratch of n elements (may be NULL -> allocated per call). */
void sfft_exec(const sfft_plan *p, const double complex *in, double complex *out, int sign);

#endif
