/*
 * fftw3_shim.c -- TEST-INFRASTRUCTURE real-to-complex FFT behind the
 * fftw3.h stand-in (see that header for why it exists and which reference
 * call sites it serves).  Independent implementation: single-precision
 * arithmetic with double-precision-derived twiddles (the accuracy class of
 * FFTW's float codelets), decimation-in-time mixed radix (2,3,4,5 + generic
 * odd radix), even sizes via the half-length complex transform.
 *
 * Only oracle/_ref/ artefacts link this.  Never part of the product path.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "fftw3.h"

typedef struct { float r, i; } cpx;

struct oracle_fftwf_plan_s {
    int n;          /* real transform length */
    int nc;         /* complex transform length (n/2 if n even, else n) */
    int even;
    float *in;
    fftwf_complex *out;
    int nfac;
    int fac[64];    /* pairs (radix, remaining) */
    cpx *tw;        /* nc twiddles exp(-2 pi i k / nc) */
    cpx *rtw;       /* n/2+1 twiddles exp(-2 pi i k / n) for the real split */
    cpx *work;      /* nc */
    cpx *scratch;   /* generic-radix scratch, max radix */
    cpx *cin;       /* nc, used when n is odd */
};

/* every buffer gets its own cache lines (128-byte aligned, padded to a multiple of 128):
 * the multi-threaded CPU baseline runs one plan per thread and must not false-share */
static void *alloc_lines(size_t n)
{
    void *p = NULL;
    n = (n + 127) & ~(size_t)127;
    if (posix_memalign(&p, 128, n ? n : 128) != 0)
	return NULL;
    return p;
}

void *fftwf_malloc(size_t n) { return alloc_lines(n); }

void fftwf_free(void *p) { free(p); }

static void factorize(int n, int *fac, int *nfac)
{
    int p = 4, k = 0;
    double floor_sqrt = floor(sqrt((double)n));
    do {
	while (n % p) {
	    switch (p) {
		case 4: p = 2; break;
		case 2: p = 3; break;
		default: p += 2; break;
	    }
	    if (p > floor_sqrt)
		p = n;
	}
	n /= p;
	fac[k++] = p;
	fac[k++] = n;
    } while (n > 1);
    *nfac = k / 2;
}

static inline cpx cmul(cpx a, cpx b)
{
    cpx c = { a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r };
    return c;
}
static inline cpx cadd(cpx a, cpx b) { cpx c = { a.r + b.r, a.i + b.i }; return c; }
static inline cpx csub(cpx a, cpx b) { cpx c = { a.r - b.r, a.i - b.i }; return c; }

static void bfly2(cpx *out, size_t fstride, const cpx *tw, int m)
{
    cpx *o2 = out + m;
    for (int k = 0; k < m; k++) {
	cpx t = cmul(o2[k], tw[k * fstride]);
	o2[k] = csub(out[k], t);
	out[k] = cadd(out[k], t);
    }
}

static void bfly3(cpx *out, size_t fstride, const cpx *tw, int m, int nc)
{
    const float s3 = tw[fstride * m].i;	/* -sin(2pi/3) */
    (void)nc;
    for (int k = 0; k < m; k++) {
	cpx a = out[k];
	cpx b = cmul(out[k + m], tw[k * fstride]);
	cpx c = cmul(out[k + 2 * m], tw[2 * k * fstride]);
	cpx s = cadd(b, c), d = csub(b, c);
	cpx h = { a.r - 0.5f * s.r, a.i - 0.5f * s.i };
	cpx e = { d.r * s3, d.i * s3 };
	out[k] = cadd(a, s);
	out[k + m].r = h.r - e.i;  out[k + m].i = h.i + e.r;
	out[k + 2 * m].r = h.r + e.i;  out[k + 2 * m].i = h.i - e.r;
    }
}

static void bfly4(cpx *out, size_t fstride, const cpx *tw, int m)
{
    for (int k = 0; k < m; k++) {
	cpx a = out[k];
	cpx b = cmul(out[k + m], tw[k * fstride]);
	cpx c = cmul(out[k + 2 * m], tw[2 * k * fstride]);
	cpx d = cmul(out[k + 3 * m], tw[3 * k * fstride]);
	cpx s0 = cadd(a, c), s1 = csub(a, c);
	cpx s2 = cadd(b, d), s3 = csub(b, d);
	out[k] = cadd(s0, s2);
	out[k + 2 * m] = csub(s0, s2);
	/* forward transform: multiply s3 by -i */
	out[k + m].r = s1.r + s3.i;  out[k + m].i = s1.i - s3.r;
	out[k + 3 * m].r = s1.r - s3.i;  out[k + 3 * m].i = s1.i + s3.r;
    }
}

static void bfly5(cpx *out, size_t fstride, const cpx *tw, int m)
{
    const cpx ya = tw[fstride * m], yb = tw[fstride * 2 * m];
    for (int k = 0; k < m; k++) {
	cpx x0 = out[k];
	cpx x1 = cmul(out[k + m], tw[k * fstride]);
	cpx x2 = cmul(out[k + 2 * m], tw[2 * k * fstride]);
	cpx x3 = cmul(out[k + 3 * m], tw[3 * k * fstride]);
	cpx x4 = cmul(out[k + 4 * m], tw[4 * k * fstride]);
	cpx s7 = cadd(x1, x4), s10 = csub(x1, x4);
	cpx s8 = cadd(x2, x3), s9 = csub(x2, x3);
	out[k].r = x0.r + s7.r + s8.r;
	out[k].i = x0.i + s7.i + s8.i;
	cpx s5 = { x0.r + s7.r * ya.r + s8.r * yb.r, x0.i + s7.i * ya.r + s8.i * yb.r };
	cpx s6 = { s10.i * ya.i + s9.i * yb.i, -s10.r * ya.i - s9.r * yb.i };
	out[k + m] = csub(s5, s6);
	out[k + 4 * m] = cadd(s5, s6);
	cpx s11 = { x0.r + s7.r * yb.r + s8.r * ya.r, x0.i + s7.i * yb.r + s8.i * ya.r };
	cpx s12 = { -s10.i * yb.i + s9.i * ya.i, s10.r * yb.i - s9.r * ya.i };
	out[k + 2 * m] = cadd(s11, s12);
	out[k + 3 * m] = csub(s11, s12);
    }
}

static void bfly_generic(cpx *out, size_t fstride, const cpx *tw, int m, int p,
	int nc, cpx *scratch)
{
    for (int u = 0; u < m; u++) {
	int k = u;
	for (int q = 0; q < p; q++) {
	    scratch[q] = out[k];
	    k += m;
	}
	k = u;
	for (int q = 0; q < p; q++) {
	    size_t twidx = 0;
	    cpx acc = scratch[0];
	    for (int j = 1; j < p; j++) {
		twidx += fstride * (size_t)k;
		if (twidx >= (size_t)nc)
		    twidx -= nc;
		acc = cadd(acc, cmul(scratch[j], tw[twidx]));
	    }
	    out[k] = acc;
	    k += m;
	}
    }
}

static void fft_work(const struct oracle_fftwf_plan_s *pl, cpx *out,
	const cpx *in, size_t fstride, const int *fac)
{
    const int p = fac[0], m = fac[1];
    cpx *o = out, *oend = out + (size_t)p * m;
    if (m == 1) {
	do {
	    *o++ = *in;
	    in += fstride;
	} while (o != oend);
    } else {
	do {
	    fft_work(pl, o, in, fstride * p, fac + 2);
	    in += fstride;
	    o += m;
	} while (o != oend);
    }
    switch (p) {
	case 2: bfly2(out, fstride, pl->tw, m); break;
	case 3: bfly3(out, fstride, pl->tw, m, pl->nc); break;
	case 4: bfly4(out, fstride, pl->tw, m); break;
	case 5: bfly5(out, fstride, pl->tw, m); break;
	default: bfly_generic(out, fstride, pl->tw, m, p, pl->nc, pl->scratch); break;
    }
}

fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int *n, int howmany,
	float *in, const int *inembed, int istride, int idist,
	fftwf_complex *out, const int *onembed, int ostride, int odist,
	unsigned flags)
{
    (void)inembed; (void)onembed; (void)idist; (void)odist; (void)flags;
    if (rank != 1 || howmany != 1 || istride != 1 || ostride != 1 || n[0] < 1)
	return NULL;
    struct oracle_fftwf_plan_s *pl = alloc_lines(sizeof(*pl));
    if (!pl)
	return NULL;
    memset(pl, 0, sizeof(*pl));
    pl->n = n[0];
    pl->even = (pl->n % 2 == 0) && pl->n >= 2;
    pl->nc = pl->even ? pl->n / 2 : pl->n;
    pl->in = in;
    pl->out = out;
    factorize(pl->nc, pl->fac, &pl->nfac);
    int maxp = 1;
    for (int i = 0; i < pl->nfac; i++)
	if (pl->fac[2 * i] > maxp)
	    maxp = pl->fac[2 * i];
    pl->tw = alloc_lines(sizeof(cpx) * (size_t)pl->nc);
    pl->work = alloc_lines(sizeof(cpx) * (size_t)pl->nc);
    pl->scratch = alloc_lines(sizeof(cpx) * (size_t)maxp);
    pl->cin = alloc_lines(sizeof(cpx) * (size_t)pl->nc);
    pl->rtw = alloc_lines(sizeof(cpx) * (size_t)(pl->n / 2 + 1));
    if (!pl->tw || !pl->work || !pl->scratch || !pl->cin || !pl->rtw) {
	fftwf_destroy_plan(pl);
	return NULL;
    }
    for (int k = 0; k < pl->nc; k++) {
	double ph = -2.0 * M_PI * (double)k / (double)pl->nc;
	pl->tw[k].r = (float)cos(ph);
	pl->tw[k].i = (float)sin(ph);
    }
    for (int k = 0; k <= pl->n / 2; k++) {
	double ph = -2.0 * M_PI * (double)k / (double)pl->n;
	pl->rtw[k].r = (float)cos(ph);
	pl->rtw[k].i = (float)sin(ph);
    }
    return pl;
}

void fftwf_execute(const fftwf_plan pl)
{
    const int n = pl->n, nc = pl->nc;
    cpx *out = (cpx *)pl->out;
    if (!pl->even) {
	for (int j = 0; j < n; j++) {
	    pl->cin[j].r = pl->in[j];
	    pl->cin[j].i = 0.0f;
	}
	fft_work(pl, pl->work, pl->cin, 1, pl->fac);
	for (int k = 0; k <= n / 2; k++)
	    out[k] = pl->work[k];
	return;
    }
    /* even n: z[j] = x[2j] + i x[2j+1] is the input array reinterpreted */
    fft_work(pl, pl->work, (const cpx *)pl->in, 1, pl->fac);
    const cpx *Z = pl->work;
    out[0].r = Z[0].r + Z[0].i;
    out[0].i = 0.0f;
    out[nc].r = Z[0].r - Z[0].i;
    out[nc].i = 0.0f;
    for (int k = 1; k < nc; k++) {
	cpx a = Z[k];
	cpx b = { Z[nc - k].r, -Z[nc - k].i };
	cpx e = { 0.5f * (a.r + b.r), 0.5f * (a.i + b.i) };	/* even part */
	cpx d = { 0.5f * (a.r - b.r), 0.5f * (a.i - b.i) };
	/* odd part = d / i = (d.i, -d.r) */
	cpx o = { d.i, -d.r };
	cpx t = cmul(o, pl->rtw[k]);
	out[k] = cadd(e, t);
    }
}

void fftwf_destroy_plan(fftwf_plan pl)
{
    if (!pl)
	return;
    free(pl->tw);
    free(pl->work);
    free(pl->scratch);
    free(pl->cin);
    free(pl->rtw);
    free(pl);
}
