/*
 * fftw3.h -- TEST-INFRASTRUCTURE stand-in for the subset of FFTW3 (single
 * precision) that the reference's src/fsk.c calls.  FFTW itself is a
 * third-party dependency of the reference (configure.ac:16 "fftw3f", version
 * un-pinned) and is absent from /root/reference and from this image.
 *
 * Call sites served (reference src/fsk.c): fftwf_malloc :73,:75; fftwf_free
 * :86-87,:100-101; fftwf_plan_many_dft_r2c :78-82; fftwf_execute :157,:552;
 * fftwf_destroy_plan :102.
 *
 * This file is written from the published FFTW API (function names and
 * argument meaning only); the transform behind it (fftw3_shim.c) is an
 * independent mixed-radix Cooley-Tukey implementation.  It is used ONLY to
 * compile the unmodified reference sources into oracle/_ref/ as the parity
 * checker / CPU baseline.  Nothing in the product path includes it.
 */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct oracle_fftwf_plan_s *fftwf_plan;

#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftwf_malloc(size_t n);
void  fftwf_free(void *p);

/* Only rank==1, howmany==1, unit strides are supported (that is all
 * src/fsk.c:78-82 asks for: pa_nchannels is fixed at 1). */
fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int *n, int howmany,
	float *in, const int *inembed, int istride, int idist,
	fftwf_complex *out, const int *onembed, int ostride, int odist,
	unsigned flags);

void fftwf_execute(const fftwf_plan plan);
void fftwf_destroy_plan(fftwf_plan plan);

#ifdef __cplusplus
}
#endif
#endif
