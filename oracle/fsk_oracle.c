/*
 * fsk_oracle.c -- TEST INFRASTRUCTURE (see fsk_oracle.h).  CPU restatement of
 * the reference FSK demodulation hot path; the parity checker, never the
 * product.  Citations are path:line under /root/reference/.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <errno.h>
#include <pthread.h>
#include <sched.h>
#include <time.h>

#include "fsk_oracle.h"

/* ------------------------------------------------------------------------ */
/* plan (src/fsk.c:33-95)                                                   */
/* ------------------------------------------------------------------------ */

int orc_plan_init(orc_plan *p, float sample_rate, float f_mark, float f_space,
	float filter_bw)
{
    memset(p, 0, sizeof(*p));
    p->sample_rate = sample_rate;
    p->f_mark = f_mark;
    p->f_space = f_space;
    p->band_width = filter_bw;				/* :50 */
    float half_bw = p->band_width / 2.0f;		/* :52 */
    p->fftsize = (sample_rate + half_bw) / p->band_width;	/* :53 float -> int */
    p->nbands = p->fftsize / 2 + 1;			/* :54 */
    p->b_mark = (f_mark + half_bw) / p->band_width;	/* :56 float -> unsigned */
    p->b_space = (f_space + half_bw) / p->band_width;	/* :57 */
    if (p->b_mark >= p->nbands || p->b_space >= p->nbands) {	/* :58-64 */
	errno = EINVAL;
	return -1;
    }
    return 0;
}

void orc_plan_free(orc_plan *p)
{
    free(p->tw);
    p->tw = NULL;
    p->tw_n = 0;
}

/* exp(-2 pi i k n / fftsize) for both tone bins, argument reduced exactly in
 * integers so the table is accurate for any window length. */
static int plan_twiddles(orc_plan *p, unsigned n)
{
    if (p->tw && p->tw_n >= n)
	return 0;
    free(p->tw);
    p->tw = malloc(sizeof(double) * 4 * (size_t)(n ? n : 1));
    if (!p->tw)
	return -1;
    const unsigned long long F = (unsigned long long)p->fftsize;
    for (unsigned i = 0; i < n; i++) {
	double am = 2.0 * M_PI * (double)(((unsigned long long)p->b_mark * i) % F) / (double)F;
	double as = 2.0 * M_PI * (double)(((unsigned long long)p->b_space * i) % F) / (double)F;
	p->tw[4 * i + 0] = cos(am);
	p->tw[4 * i + 1] = sin(am);
	p->tw[4 * i + 2] = cos(as);
	p->tw[4 * i + 3] = sin(as);
    }
    p->tw_n = n;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* bit analyzer (src/fsk.c:107-174)                                         */
/* ------------------------------------------------------------------------ */

void orc_bit_mags(orc_plan *p, const float *samples, unsigned bit_nsamples,
	float *mag_mark, float *mag_space)
{
    /* The reference copies bit_nsamples samples into the zeroed fftin (:130),
     * runs the whole r2c FFT (:157) and reads bins b_mark, b_space (:158-159).
     * Samples beyond fftsize could not be copied there; window lengths are
     * always <= fftsize in the reference (band_width <= data_rate, :960). */
    plan_twiddles(p, bit_nsamples);
    double rm = 0, im = 0, rs = 0, is = 0;
    const double *tw = p->tw;
    for (unsigned i = 0; i < bit_nsamples; i++) {
	double x = samples[i];
	rm += x * tw[4 * i + 0];
	im -= x * tw[4 * i + 1];
	rs += x * tw[4 * i + 2];
	is -= x * tw[4 * i + 3];
    }
    float magscalar = 2.0f / (float)bit_nsamples;		/* :132 */
    *mag_mark  = hypotf((float)rm, (float)im) * magscalar;	/* :108-113, :158 */
    *mag_space = hypotf((float)rs, (float)is) * magscalar;	/* :159 */
}

/* ------------------------------------------------------------------------ */
/* frame analyzer (src/fsk.c:178-446, CONFIDENCE_ALGO 6)                    */
/* ------------------------------------------------------------------------ */

static void bit_decide(orc_plan *p, const float *s, unsigned n,
	unsigned *val, float *sig, float *noise)
{
    float mm, ms;
    orc_bit_mags(p, s, n, &mm, &ms);
    if (mm > ms) {		/* :161 strict: tie -> space/0 */
	*val = 1; *sig = mm; *noise = ms;
    } else {
	*val = 0; *sig = ms; *noise = mm;
    }
}

float orc_frame_analyze(orc_plan *p, const float *samples, float samples_per_bit,
	int n_bits, const char *expect_bits,
	unsigned long long *bits_out, float *ampl_out,
	float *o_sig, float *o_noise, unsigned *o_val)
{
    unsigned bit_nsamples = (float)(samples_per_bit + 0.5f);	/* :183 */
    unsigned val[64];
    float sig[64], noise[64];

    /* pass 1: the required ('0'/'1') bits; first mismatch rejects (:199-226) */
    for (int b = 0; b < n_bits; b++) {
	if (expect_bits[b] == 'd')
	    continue;
	unsigned begin = (float)(samples_per_bit * b + 0.5f);	/* :204 */
	bit_decide(p, samples + begin, bit_nsamples, &val[b], &sig[b], &noise[b]);
	if ((unsigned)(expect_bits[b] - '0') != val[b])
	    return 0.0f;					/* :211-212 */
    }
    /* pass 2: the don't-care bits (:246-261) */
    for (int b = 0; b < n_bits; b++) {
	if (expect_bits[b] != 'd')
	    continue;
	unsigned begin = (float)(samples_per_bit * b + 0.5f);	/* :249 */
	bit_decide(p, samples + begin, bit_nsamples, &val[b], &sig[b], &noise[b]);
    }

    /* :271-289 */
    float total_sig = 0.0f, total_noise = 0.0f;
    float avg_mark = 0.0f, avg_space = 0.0f;
    unsigned n_mark = 0, n_space = 0;
    for (int b = 0; b < n_bits; b++) {
	total_sig += sig[b];
	if (noise[b] > FLT_EPSILON)			/* :279 */
	    total_noise += noise[b];
	if (val[b] == 1) { avg_mark += sig[b]; n_mark++; }
	else             { avg_space += sig[b]; n_space++; }
    }
    float snr = total_sig / total_noise;		/* :292, may be inf */
    float avg_bit_sig = total_sig / n_bits;		/* :295 */
    if (n_mark)  avg_mark /= n_mark;			/* :298-301 */
    if (n_space) avg_space /= n_space;

    float divergence = 0.0f;				/* :305-313 */
    for (int b = 0; b < n_bits; b++) {
	float other = val[b] ? avg_mark : avg_space;
	divergence += fabsf(sig[b] - other) / other;
    }
    divergence *= 2;
    divergence /= n_bits;

    float confidence = snr * (1.0f - divergence);	/* :336 */
    *ampl_out = avg_bit_sig;				/* :342 */

    unsigned long long bits = 0;			/* :439-441 LSB first */
    for (int b = 0; b < n_bits; b++)
	bits |= (unsigned long long)val[b] << b;
    *bits_out = bits;

    if (o_sig)   memcpy(o_sig, sig, sizeof(float) * n_bits);
    if (o_noise) memcpy(o_noise, noise, sizeof(float) * n_bits);
    if (o_val)   memcpy(o_val, val, sizeof(unsigned) * n_bits);
    return confidence;
}

/* ------------------------------------------------------------------------ */
/* frame search (src/fsk.c:449-538)                                         */
/* ------------------------------------------------------------------------ */

float orc_find_frame(orc_plan *p, const float *samples, unsigned frame_nsamples,
	unsigned try_first_sample, unsigned try_max_nsamples,
	unsigned try_step_nsamples, float limit, const char *expect,
	unsigned long long *bits_out, float *ampl_out, unsigned *frame_start_out)
{
    int expect_n_bits = (int)strlen(expect);			/* :461 */
    float samples_per_bit = (float)frame_nsamples / expect_n_bits;	/* :465 */

    unsigned best_t = 0;
    float best_c = 0.0f, best_a = 0.0f;
    unsigned long long best_bits = 0;

    for (int j = 0; ; j++) {					/* :477-502 */
	int up = (j % 2) ? 1 : -1;
	int t = (int)try_first_sample + up * ((j + 1) / 2) * (int)try_step_nsamples;
	if (t >= (int)try_max_nsamples)
	    break;
	if (t < 0)
	    continue;
	float a = 0.0f;
	unsigned long long bits = 0;
	float c = orc_frame_analyze(p, samples + t, samples_per_bit,
		expect_n_bits, expect, &bits, &a, NULL, NULL, NULL);
	if (best_c < c) {
	    best_t = t; best_c = c; best_a = a; best_bits = bits;
	    if (best_c >= limit)
		break;
	}
    }
    *bits_out = best_bits;					/* :504-506 */
    *ampl_out = best_a;
    *frame_start_out = best_t;
    return best_c;
}

static float default_find_frame(void *ctx, const float *samples,
	unsigned frame_nsamples, unsigned try_first, unsigned try_max,
	unsigned try_step, float limit, const char *expect,
	unsigned long long *bits, float *ampl, unsigned *frame_start)
{
    return orc_find_frame((orc_plan *)ctx, samples, frame_nsamples, try_first,
	    try_max, try_step, limit, expect, bits, ampl, frame_start);
}

/* ------------------------------------------------------------------------ */
/* rx loop set-up (src/minimodem.c:442-487, :1037-1131)                     */
/* ------------------------------------------------------------------------ */

int orc_build_expect_bits_string(char *out, int nstartbits, int n_data_bits,
	float nstopbits, int invert_start_stop, int use_expect_bits,
	unsigned long long expect_bits)
{
    char start_v = invert_start_stop ? '1' : '0';
    char stop_v  = invert_start_stop ? '0' : '1';
    int j = 0;
    if (nstopbits != 0.0f)
	out[j++] = stop_v;		/* the previous frame's stop bit */
    for (int i = 0; i < nstartbits; i++)
	out[j++] = start_v;
    for (int i = 0; i < n_data_bits; i++, j++)
	out[j] = use_expect_bits ? (char)(((expect_bits >> i) & 1) + '0') : 'd';
    if (nstopbits != 0.0f)
	out[j++] = stop_v;
    out[j] = 0;
    return j;
}

void orc_rx_derive(const orc_rx_config *cfg, orc_rx_derived *d)
{
    memset(d, 0, sizeof(*d));
    unsigned int sample_rate = (unsigned int)cfg->sample_rate;
    d->nsamples_per_bit = sample_rate / cfg->data_rate;			/* :1037 */
    d->frame_n_bits = cfg->n_data_bits + cfg->nstartbits + cfg->nstopbits;	/* :943 */

    unsigned nbits = 1 + cfg->nstartbits + cfg->n_data_bits + 1;	/* :1056-1060 */
    size_t sz = ceilf(d->nsamples_per_bit) * (nbits + 1);		/* :1063 */
    sz *= 2;
    if (sz < sample_rate / 12)						/* :1068 */
	sz = sample_rate / 12;
    d->samplebuf_size = sz;

    const float overscan = 0.5f;					/* :1091 */
    d->nsamples_overscan = d->nsamples_per_bit * overscan + 0.5f;	/* :1105 */
    if (overscan > 0.0f && d->nsamples_overscan == 0)
	d->nsamples_overscan = 1;

    float frame_n_bits = d->frame_n_bits;				/* :1112 */
    d->frame_nsamples = d->nsamples_per_bit * frame_n_bits + 0.5f;	/* :1113 */

    if (cfg->expect_data_string) {
	strncpy(d->expect_data, cfg->expect_data_string, sizeof(d->expect_data) - 1);
	d->expect_n_bits = strlen(d->expect_data);
    } else {
	d->expect_n_bits = orc_build_expect_bits_string(d->expect_data,
		cfg->nstartbits, cfg->n_data_bits, cfg->nstopbits,
		cfg->invert_start_stop, 0, 0);				/* :1118 */
    }
    if (cfg->do_rx_sync && (long long)cfg->sync_byte >= 0)		/* :1123 */
	orc_build_expect_bits_string(d->expect_sync, cfg->nstartbits,
		cfg->n_data_bits, cfg->nstopbits, cfg->invert_start_stop,
		1, cfg->sync_byte);
    else
	strcpy(d->expect_sync, d->expect_data);				/* :1127 */

    d->expect_nsamples = d->nsamples_per_bit * d->expect_n_bits;	/* :1131 truncating */
}

unsigned long long orc_rx_databits(const orc_rx_config *cfg, unsigned long long bits)
{
    if (cfg->nstopbits != 0.0f)		/* :1415 chop the prev_stop bit */
	bits >>= 1;
    /* bit_window(bits, nstartbits, n_data_bits), src/databits.h:35-46 */
    unsigned long long mask = (1ULL << cfg->n_data_bits) - 1;
    if (mask == 0)
	bits = bits >> cfg->nstartbits;
    else
	bits = (bits >> cfg->nstartbits) & mask;
    if (cfg->msb_first) {		/* bit_reverse, src/databits.h:21-33: 32-bit accumulator */
	unsigned int out = 0;
	unsigned n = cfg->n_data_bits;
	unsigned long long v = bits;
	while (n--) {
	    out = (out << 1) | (v & 1);
	    v >>= 1;
	}
	bits = out;
    }
    return bits;
}

/* ------------------------------------------------------------------------ */
/* rx loop (src/minimodem.c:1137-1463)                                      */
/* ------------------------------------------------------------------------ */

#define PUSH(res, arr, n, cap, val) do { \
    if ((res)->n == (res)->cap) { \
	(res)->cap = (res)->cap ? (res)->cap * 2 : 256; \
	(res)->arr = realloc((res)->arr, sizeof(*(res)->arr) * (res)->cap); \
    } \
    (res)->arr[(res)->n++] = (val); } while (0)

void orc_rx_result_free(orc_rx_result *res)
{
    free(res->frames);
    free(res->reports);
    free(res->calls);
    memset(res, 0, sizeof(*res));
}

struct rx_src {
    const float *samples;
    size_t nsamples;
    size_t rd;			/* literal mode: next unread sample */
    float noise_add;		/* the constant --Xrxnoise adds to each sample */
    int has_noise;
};

int orc_rx_run(const orc_rx_config *cfg, const float *samples, size_t nsamples,
	int mode, float rxnoise, int rx_one, int want_calls,
	orc_find_frame_fn find_frame, void *ctx, orc_rx_result *res)
{
    orc_rx_derived d;
    orc_plan plan;
    int own_plan = 0;
    /* res must be zero-initialised before its first use; buffers are reused */
    res->nframes = res->nreports = res->ncalls = 0;
    res->n_find_frame_calls = 0;
    orc_rx_derive(cfg, &d);
    if (d.expect_n_bits == 0 || d.expect_n_bits > 64)
	return -1;
    if (!find_frame) {
	if (orc_plan_init(&plan, cfg->sample_rate, cfg->f_mark, cfg->f_space,
		    cfg->band_width) != 0)
	    return -1;
	own_plan = 1;
	find_frame = default_find_frame;
	ctx = &plan;
    }

    struct rx_src src = { samples, nsamples, 0, 0.0f, rxnoise != 0.0f };
    if (src.has_noise) {
	/* src/simpleaudio-sndfile.c:64-70: f = rxnoise*2; x += (rand()/RAND_MAX - 0.5f)*f
	 * with rand()/RAND_MAX an INTEGER division (== 0 unless rand()==RAND_MAX). */
	float f = rxnoise * 2;
	src.noise_add = (0 - 0.5f) * f;
    }

    const size_t S = d.samplebuf_size;
    /* widest read the callee can make past samplebuf[0] */
    const size_t touch_max = (size_t)(d.nsamples_per_bit + d.nsamples_overscan) + 2
	    + d.expect_nsamples + (size_t)d.nsamples_per_bit + 2;
    /* per-thread scratch, kept across calls (page faults are expensive on VMs) */
    static __thread float *scratch;
    static __thread size_t scratch_cap;
    const size_t want_floats = mode == ORC_RX_LITERAL ? S + touch_max : 2 * touch_max + 8;
    if (scratch_cap < want_floats) {
	free(scratch);
	scratch = malloc(want_floats * sizeof(float));
	scratch_cap = scratch ? want_floats : 0;
	if (!scratch)
	    return -1;
    }
    memset(scratch, 0, want_floats * sizeof(float));	/* :1071 (malloc there) */
    float *ring = mode == ORC_RX_LITERAL ? scratch : NULL;
    float *tail = mode == ORC_RX_LITERAL ? NULL : scratch;
    size_t samples_nvalid = 0;		/* literal */
    unsigned long long pos = 0;		/* absolute index of samplebuf[0] */

    int carrier = 0;					/* :1081-1088 */
    float confidence_total = 0, amplitude_total = 0;
    unsigned nframes_decoded = 0;
    size_t carrier_nsamples = 0;
    unsigned noconfidence = 0;
    unsigned advance = 0;
    float track_amplitude = 0.0f, peak_confidence = 0.0f;	/* :1132-1133 */

    for (;;) {
	const float *buf;
	size_t nvalid;
	if (mode == ORC_RX_LITERAL) {
	    if (advance == S) {				/* :1146-1149 */
		samples_nvalid = 0;
		pos += advance;
		advance = 0;
	    }
	    if (advance) {				/* :1150-1156 */
		if (advance > samples_nvalid)
		    break;
		memmove(ring, ring + advance, (S - advance) * sizeof(float));
		samples_nvalid -= advance;
		pos += advance;
	    }
	    if (samples_nvalid < S / 2) {		/* :1158-1174 */
		size_t want = S / 2;
		size_t r = src.nsamples - src.rd;
		if (r > want) r = want;
		memcpy(ring + samples_nvalid, src.samples + src.rd, r * sizeof(float));
		if (src.has_noise)
		    for (size_t i = 0; i < want; i++)
			ring[samples_nvalid + i] += src.noise_add;
		src.rd += r;
		samples_nvalid += r;
	    }
	    nvalid = samples_nvalid;
	    buf = ring;
	} else {
	    /* flat restatement: nvalid is "everything that remains"; the
	     * reference's nvalid differs from it only while >= S/2, where no
	     * test below can tell the difference (rx loop). */
	    size_t remaining = src.nsamples - (size_t)pos;
	    if (advance) {
		if (advance > remaining)
		    break;
		pos += advance;
		remaining -= advance;
	    }
	    nvalid = remaining;
	    buf = NULL;
	}
	if (nvalid == 0)				/* :1176 */
	    break;
	if (nvalid < d.expect_nsamples)			/* :1229 */
	    break;

	unsigned try_max;				/* :1236-1241 */
	if (carrier)
	    try_max = d.nsamples_per_bit * 0.75f + 0.5f;
	else
	    try_max = d.nsamples_per_bit;
	try_max += d.nsamples_overscan;
	unsigned try_step = try_max / 3;		/* :1248-1251 */
	if (try_step == 0)
	    try_step = 1;

	if (mode != ORC_RX_LITERAL) {
	    /* samples past the end read as zero (batched-API semantic) */
	    size_t need = (size_t)try_max + d.expect_nsamples + (size_t)d.nsamples_per_bit + 2;
	    if (need > 2 * touch_max) need = 2 * touch_max;
	    if (nvalid >= need && !src.has_noise) {
		buf = src.samples + pos;
	    } else {
		size_t have = nvalid < need ? nvalid : need;
		memset(tail, 0, (2 * touch_max + 8) * sizeof(float));
		for (size_t i = 0; i < have; i++)
		    tail[i] = src.samples[pos + i] + (src.has_noise ? src.noise_add : 0.0f);
		buf = tail;
	    }
	}

	float confidence, amplitude = 0.0f;
	unsigned long long bits = 0;
	unsigned frame_start_sample = 0;
	float limit = cfg->confidence_search_limit;	/* :1262 */
	unsigned try_first = carrier ? d.nsamples_overscan : 0;	/* :1263 */
	const char *expect = carrier ? d.expect_data : d.expect_sync;	/* :1270 */

	confidence = find_frame(ctx, buf, d.expect_nsamples, try_first, try_max,
		try_step, limit, expect, &bits, &amplitude, &frame_start_sample);
	res->n_find_frame_calls++;
	if (want_calls) {
	    orc_rx_call c = { d.expect_nsamples, try_first, try_max, try_step, limit,
		!carrier, confidence, bits, amplitude, frame_start_sample, pos };
	    PUSH(res, calls, ncalls, cap_calls, c);
	}

	int do_refine_frame = 0;
	if (confidence < peak_confidence * 0.75f) {	/* :1278-1282 */
	    do_refine_frame = 1;
	    peak_confidence = 0;
	}
	if (amplitude < track_amplitude * 0.25f)	/* :1286 */
	    confidence = 0;

	if (confidence <= cfg->confidence_threshold) {	/* :1292 */
	    if (++noconfidence > 20) {			/* :1295 FSK_MAX_NOCONFIDENCE_BITS */
		if (carrier) {
		    orc_rx_report rp = { nframes_decoded, carrier_nsamples,
			confidence_total, amplitude_total, (unsigned)res->nframes };
		    PUSH(res, reports, nreports, cap_reports, rp);
		    carrier = 0;
		    carrier_nsamples = 0;
		    confidence_total = 0;
		    amplitude_total = 0;
		    nframes_decoded = 0;
		    track_amplitude = 0.0f;
		    if (rx_one)
			break;
		}
	    }
	    advance = try_max;				/* :1318 */
	    continue;
	}

	carrier_nsamples += d.frame_nsamples;		/* :1324 */
	unsigned acquired = 0;
	if (carrier) {
	    carrier_nsamples += frame_start_sample;	/* :1329-1330 */
	    carrier_nsamples -= d.nsamples_overscan;
	} else {
	    carrier = 1;				/* :1350-1353 */
	    acquired = 1;
	    do_refine_frame = 1;
	}

	if (do_refine_frame) {				/* :1357-1389 */
	    if (confidence < INFINITY && try_step > 1) {
		try_step = try_max / 8;
		if (try_step == 0)
		    try_step = 1;
		float confidence2, amplitude2 = 0.0f;
		unsigned long long bits2 = 0;
		unsigned fss2 = 0;
		/* the string choice re-evaluates `carrier`, which is 1 by now */
		confidence2 = find_frame(ctx, buf, d.expect_nsamples, try_first,
			try_max, try_step, INFINITY,
			carrier ? d.expect_data : d.expect_sync,
			&bits2, &amplitude2, &fss2);
		res->n_find_frame_calls++;
		if (want_calls) {
		    orc_rx_call c = { d.expect_nsamples, try_first, try_max, try_step,
			INFINITY, !carrier, confidence2, bits2, amplitude2, fss2, pos };
		    PUSH(res, calls, ncalls, cap_calls, c);
		}
		if (confidence2 > confidence) {
		    bits = bits2;
		    amplitude = amplitude2;
		    frame_start_sample = fss2;
		}
	    }
	}

	track_amplitude = (track_amplitude + amplitude) / 2;	/* :1391 */
	if (peak_confidence < confidence)
	    peak_confidence = confidence;
	confidence_total += confidence;			/* :1397-1400 */
	amplitude_total += amplitude;
	nframes_decoded++;
	noconfidence = 0;

	advance = frame_start_sample + d.frame_nsamples - d.nsamples_overscan;	/* :1407 */

	orc_rx_frame fr = { bits, confidence, amplitude, frame_start_sample, acquired, pos };
	PUSH(res, frames, nframes, cap_frames, fr);
    }

    if (carrier) {					/* :1469-1474 */
	orc_rx_report rp = { nframes_decoded, carrier_nsamples,
	    confidence_total, amplitude_total, (unsigned)res->nframes };
	PUSH(res, reports, nreports, cap_reports, rp);
    }
    if (own_plan)
	orc_plan_free(&plan);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* multi-threaded driver for the CPU baseline                               */
/* ------------------------------------------------------------------------ */

struct many_job {
    const orc_rx_config *cfg;
    const float *samples;
    size_t nstreams, stride, nsamples;
    int tid, nthreads;
    unsigned *frames_per_stream;
    unsigned long long *bits_xor;
    unsigned long long total;
    orc_plan_new_fn plan_new;
    orc_find_frame_fn find_frame;
    orc_plan_destroy_fn plan_destroy;
};

static void *many_worker(void *arg)
{
    struct many_job *j = arg;
    /* spread the workers over the allowed CPUs at once: short runs otherwise stay
     * packed on one vCPU until the scheduler's load balancer notices */
    cpu_set_t allowed, one;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
	int ncpu = CPU_COUNT(&allowed), want = j->tid % (ncpu ? ncpu : 1), seen = 0;
	for (int c = 0; c < CPU_SETSIZE; c++) {
	    if (!CPU_ISSET(c, &allowed))
		continue;
	    if (seen++ == want) {
		CPU_ZERO(&one);
		CPU_SET(c, &one);
		pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
		break;
	    }
	}
    }
    orc_plan plan;
    void *ctx = &plan;
    orc_find_frame_fn ff = default_find_frame;
    if (j->plan_new) {
	ctx = j->plan_new(j->cfg->sample_rate, j->cfg->f_mark, j->cfg->f_space, j->cfg->band_width);
	ff = j->find_frame;
	if (!ctx)
	    return NULL;
    } else if (orc_plan_init(&plan, j->cfg->sample_rate, j->cfg->f_mark, j->cfg->f_space,
		j->cfg->band_width) != 0)
	return NULL;
    orc_rx_result r;
    memset(&r, 0, sizeof(r));
    for (size_t s = j->tid; s < j->nstreams; s += j->nthreads) {
	orc_rx_run(j->cfg, j->samples + s * j->stride, j->nsamples, ORC_RX_FLAT,
		0.0f, 0, 0, ff, ctx, &r);
	unsigned long long x = 0;
	for (size_t i = 0; i < r.nframes; i++)
	    x ^= r.frames[i].bits * (i + 1);
	if (j->frames_per_stream) j->frames_per_stream[s] = (unsigned)r.nframes;
	if (j->bits_xor) j->bits_xor[s] = x;
	j->total += r.nframes;
    }
    orc_rx_result_free(&r);
    if (j->plan_new)
	j->plan_destroy(ctx);
    else
	orc_plan_free(&plan);
    return NULL;
}

unsigned long long orc_rx_many(const orc_rx_config *cfg, const float *samples,
	size_t nstreams, size_t stride, size_t nsamples, int nthreads,
	orc_plan_new_fn plan_new, orc_find_frame_fn find_frame, orc_plan_destroy_fn plan_destroy,
	unsigned *frames_per_stream, unsigned long long *bits_xor_per_stream)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = calloc(nthreads, sizeof(*th));
    struct many_job *jobs = calloc(nthreads, sizeof(*jobs));
    for (int t = 0; t < nthreads; t++) {
	jobs[t] = (struct many_job){ cfg, samples, nstreams, stride, nsamples, t,
	    nthreads, frames_per_stream, bits_xor_per_stream, 0,
	    plan_new, find_frame, plan_destroy };
	pthread_create(&th[t], NULL, many_worker, &jobs[t]);
    }
    unsigned long long total = 0;
    for (int t = 0; t < nthreads; t++) {
	pthread_join(th[t], NULL);
	total += jobs[t].total;
    }
    free(th);
    free(jobs);
    return total;
}

/* ------------------------------------------------------------------------ */
/* persistent worker pool for the CPU timing arms                             */
/* ------------------------------------------------------------------------ */
/* orc_rx_many starts its threads and builds one plan per thread inside every call, walks the
 * streams strided across threads and runs on pages first touched by whoever filled the buffer:
 * on a 2-socket host the same pass came out anywhere between 4 and 13 Gsamples/s.  The pool
 * fixes what can be fixed from here: workers are created once and pinned one per allowed CPU,
 * each builds its plan once, owns a CONTIGUOUS block of streams and first-touches (copies) that
 * block into pool-owned memory, so a timed pass is nothing but the rx loops; the pass time is
 * taken inside, from the release of the start barrier to the last worker's arrival. */
struct orc_pool;
struct pool_worker {
    struct orc_pool *pool;
    int tid;
    pthread_t th;
    void *ctx;
    orc_plan plan;
    orc_find_frame_fn ff;
    unsigned long long total;
    int ok;
};
struct orc_pool {
    orc_rx_config cfg;
    int nthreads;
    orc_plan_new_fn plan_new;
    orc_find_frame_fn find_frame;
    orc_plan_destroy_fn plan_destroy;
    struct pool_worker *w;
    pthread_barrier_t start, stop;
    int cmd;			/* 1 load, 2 run, 0 exit */
    /* the loaded batch */
    float *buf;
    const float *src;
    size_t nstreams, stride, src_stride, nsamples;
    unsigned *frames_per_stream;
    unsigned long long *bits_xor;
};

/* Pinning order: the allowed CPUs sorted so that one hardware thread of every physical core comes
 * first (sockets alternating), the cores' second hardware threads after them -- N workers then
 * occupy min(N, cores) distinct cores whatever the kernel's CPU numbering is (siblings adjacent,
 * or all first threads before all second ones).  ORC_POOL_PIN=none leaves the scheduler alone,
 * ORC_POOL_PIN=compact takes the allowed CPUs in numeric order. */
static int pool_cpu_order(int *order, int cap)
{
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0)
	return 0;
    struct { int cpu, pkg, core; } c[CPU_SETSIZE];
    int n = 0;
    for (int cpu = 0; cpu < CPU_SETSIZE && n < cap; cpu++) {
	if (!CPU_ISSET(cpu, &allowed))
	    continue;
	char path[128];
	int pkg = 0, core = cpu;
	FILE *f;
	snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/physical_package_id", cpu);
	if ((f = fopen(path, "r"))) { if (fscanf(f, "%d", &pkg) != 1) pkg = 0; fclose(f); }
	snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/core_id", cpu);
	if ((f = fopen(path, "r"))) { if (fscanf(f, "%d", &core) != 1) core = cpu; fclose(f); }
	c[n].cpu = cpu; c[n].pkg = pkg; c[n].core = core;
	n++;
    }
    const char *mode = getenv("ORC_POOL_PIN");
    if (mode && strcmp(mode, "compact") == 0) {
	for (int i = 0; i < n; i++)
	    order[i] = c[i].cpu;
	return n;
    }
    /* rank of a CPU among the hardware threads of its core (0 = first thread) */
    int rank[CPU_SETSIZE];
    for (int i = 0; i < n; i++) {
	rank[i] = 0;
	for (int j = 0; j < i; j++)
	    if (c[j].pkg == c[i].pkg && c[j].core == c[i].core)
		rank[i]++;
    }
    int k = 0;
    for (int r = 0; r < 8 && k < n; r++)
	for (int i = 0; i < n; i++)
	    if (rank[i] == r)
		order[k++] = c[i].cpu;
    return k;
}

static void pool_pin(int tid)
{
    const char *mode = getenv("ORC_POOL_PIN");
    if (mode && strcmp(mode, "none") == 0)
	return;
    int order[CPU_SETSIZE];
    const int norder = pool_cpu_order(order, CPU_SETSIZE);
    if (norder <= 0)
	return;
    cpu_set_t one;
    CPU_ZERO(&one);
    CPU_SET(order[tid % norder], &one);
    pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
}

static void *pool_worker_main(void *arg)
{
    struct pool_worker *w = arg;
    struct orc_pool *p = w->pool;
    pool_pin(w->tid);
    w->ok = 1;
    if (p->plan_new) {
	w->ctx = p->plan_new(p->cfg.sample_rate, p->cfg.f_mark, p->cfg.f_space, p->cfg.band_width);
	w->ff = p->find_frame;
	if (!w->ctx)
	    w->ok = 0;
    } else {
	w->ctx = &w->plan;
	w->ff = default_find_frame;
	if (orc_plan_init(&w->plan, p->cfg.sample_rate, p->cfg.f_mark, p->cfg.f_space, p->cfg.band_width) != 0)
	    w->ok = 0;
    }
    orc_rx_result r;
    memset(&r, 0, sizeof(r));
    for (;;) {
	pthread_barrier_wait(&p->start);
	const int cmd = p->cmd;
	if (cmd == 0)
	    break;
	const size_t s0 = p->nstreams * (size_t)w->tid / (size_t)p->nthreads;
	const size_t s1 = p->nstreams * (size_t)(w->tid + 1) / (size_t)p->nthreads;
	if (cmd == 1) {			/* first touch + copy of this worker's block */
	    for (size_t s = s0; s < s1; s++) {
		float *d = p->buf + s * p->stride;
		memcpy(d, p->src + s * p->src_stride, p->nsamples * sizeof(float));
		if (p->stride > p->nsamples)
		    memset(d + p->nsamples, 0, (p->stride - p->nsamples) * sizeof(float));
	    }
	} else if (cmd == 2 && w->ok) {
	    w->total = 0;
	    for (size_t s = s0; s < s1; s++) {
		orc_rx_run(&p->cfg, p->buf + s * p->stride, p->nsamples, ORC_RX_FLAT, 0.0f, 0, 0,
			w->ff, w->ctx, &r);
		unsigned long long x = 0;
		for (size_t i = 0; i < r.nframes; i++)
		    x ^= r.frames[i].bits * (i + 1);
		if (p->frames_per_stream) p->frames_per_stream[s] = (unsigned)r.nframes;
		if (p->bits_xor) p->bits_xor[s] = x;
		w->total += r.nframes;
	    }
	}
	pthread_barrier_wait(&p->stop);
    }
    orc_rx_result_free(&r);
    if (w->ok) {
	if (p->plan_new)
	    p->plan_destroy(w->ctx);
	else
	    orc_plan_free(&w->plan);
    }
    return NULL;
}

struct orc_pool *orc_pool_new(const orc_rx_config *cfg, int nthreads, orc_plan_new_fn plan_new,
	orc_find_frame_fn find_frame, orc_plan_destroy_fn plan_destroy)
{
    if (nthreads < 1) nthreads = 1;
    struct orc_pool *p = calloc(1, sizeof(*p));
    if (!p)
	return NULL;
    p->cfg = *cfg;
    p->nthreads = nthreads;
    p->plan_new = plan_new;
    p->find_frame = find_frame;
    p->plan_destroy = plan_destroy;
    p->w = calloc((size_t)nthreads, sizeof(*p->w));
    pthread_barrier_init(&p->start, NULL, (unsigned)nthreads + 1);
    pthread_barrier_init(&p->stop, NULL, (unsigned)nthreads + 1);
    for (int t = 0; t < nthreads; t++) {
	p->w[t].pool = p;
	p->w[t].tid = t;
	pthread_create(&p->w[t].th, NULL, pool_worker_main, &p->w[t]);
    }
    return p;
}

/* copies [nstreams][src_stride] floats (nsamples valid per row) into pool memory, every worker
 * touching its own block first */
int orc_pool_load(struct orc_pool *p, const float *samples, size_t nstreams, size_t src_stride,
	size_t nsamples)
{
    free(p->buf);
    p->stride = (nsamples + 15) & ~(size_t)15;
    p->buf = NULL;
    if (posix_memalign((void **)&p->buf, 4096, nstreams * p->stride * sizeof(float) + 4096) != 0)
	return -1;
    p->src = samples;
    p->src_stride = src_stride;
    p->nstreams = nstreams;
    p->nsamples = nsamples;
    p->cmd = 1;
    pthread_barrier_wait(&p->start);
    pthread_barrier_wait(&p->stop);
    p->src = NULL;
    return 0;
}

/* one pass over the loaded batch; returns its wall time in seconds (taken here, around the two
 * barriers) and the number of frames decoded */
double orc_pool_run(struct orc_pool *p, unsigned long long *total, unsigned *frames_per_stream,
	unsigned long long *bits_xor_per_stream)
{
    struct timespec t0, t1;
    p->frames_per_stream = frames_per_stream;
    p->bits_xor = bits_xor_per_stream;
    p->cmd = 2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&p->start);
    pthread_barrier_wait(&p->stop);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    unsigned long long sum = 0;
    int ok = 1;
    for (int t = 0; t < p->nthreads; t++) {
	sum += p->w[t].total;
	ok &= p->w[t].ok;
    }
    if (total)
	*total = sum;
    if (!ok)
	return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void orc_pool_free(struct orc_pool *p)
{
    if (!p)
	return;
    p->cmd = 0;
    pthread_barrier_wait(&p->start);
    for (int t = 0; t < p->nthreads; t++)
	pthread_join(p->w[t].th, NULL);
    pthread_barrier_destroy(&p->start);
    pthread_barrier_destroy(&p->stop);
    free(p->buf);
    free(p->w);
    free(p);
}

/* ------------------------------------------------------------------------ */
/* tx (src/minimodem.c:81-250, src/simple-tone-generator.c:38-175)          */
/* ------------------------------------------------------------------------ */

struct tonegen {
    unsigned rate;
    float mag;
    unsigned table_len;
    float *table_f;
    short *table_s;
    int s16;
    float cphase;		/* sa_tone_cphase, simple-tone-generator.c:92 */
    float *out;
    size_t n, cap;
};

static void tonegen_init(struct tonegen *g, const orc_tx_config *cfg, float *out, size_t cap)
{
    memset(g, 0, sizeof(*g));
    g->rate = (unsigned)cfg->sample_rate;
    g->mag = cfg->amplitude;
    g->table_len = cfg->sin_table_len;
    g->s16 = cfg->s16;
    g->out = out;
    g->cap = cap;
    if (g->table_len) {					/* simple-tone-generator.c:38-58 */
	g->table_f = malloc(sizeof(float) * g->table_len);
	g->table_s = malloc(sizeof(short) * g->table_len);
	unsigned short mag_s = 32767.0f * g->mag + 0.5f;
	if (g->mag > 1.0f) mag_s = 32767;
	if (mag_s < 1) mag_s = 1;
	for (unsigned i = 0; i < g->table_len; i++)
	    g->table_s[i] = lroundf(mag_s * sinf((float)M_PI * 2 * i / g->table_len));
	for (unsigned i = 0; i < g->table_len; i++)
	    g->table_f[i] = g->mag * sinf((float)M_PI * 2 * i / g->table_len);
    }
}

static void tonegen_free(struct tonegen *g)
{
    free(g->table_f);
    free(g->table_s);
}

/* simpleaudio_tone, simple-tone-generator.c:107-175 (tone_freq != 0 only) */
static void tonegen_tone(struct tonegen *g, float tone_freq, size_t nsamples_dur)
{
    float wave_nsamples = g->rate / tone_freq;			/* :116 */
    for (size_t i = 0; i < nsamples_dur && g->n < g->cap; i++) {
	float turns = (float)i / wave_nsamples + g->cphase;	/* :121 */
	float v;
	if (g->s16) {
	    short sv;
	    if (g->table_s) {					/* :77-84 */
		int t = (float)g->table_len * turns + 0.5f;
		t %= g->table_len;
		sv = g->table_s[t];
	    } else {						/* :145-152 */
		unsigned short mag_s = 32767.0f * g->mag + 0.5f;
		if (g->mag > 1.0f) mag_s = 32767;
		if (mag_s < 1) mag_s = 1;
		sv = lroundf(mag_s * sinf((float)M_PI * 2 * turns));
	    }
	    /* written as PCM16, read back by the rx as float (libsndfile scales
	     * by 1/32768; src/minimodem.c:786-788 forces float reads) */
	    v = (float)sv * (1.0f / 32768.0f);
	} else {
	    if (g->table_f) {					/* :89-96 */
		int t = (float)g->table_len * turns + 0.5f;
		t %= g->table_len;
		v = g->table_f[t];
	    } else {
		v = g->mag * sinf((float)M_PI * 2 * turns);	/* :134 */
	    }
	}
	g->out[g->n++] = v;
    }
    g->cphase = fmodf(g->cphase + (float)nsamples_dur / wave_nsamples, 1.0f);	/* :163-164 */
}

struct tx_lengths { size_t bit, start, stop; };

static struct tx_lengths tx_lengths(const orc_tx_config *cfg)
{
    struct tx_lengths L;
    size_t sample_rate = (size_t)cfg->sample_rate;
    L.bit = sample_rate / cfg->data_rate + 0.5f;	/* src/minimodem.c:132 */
    L.start = L.bit * cfg->nstartbits;			/* :96-97 size_t*float -> size_t */
    L.stop = L.bit * cfg->nstopbits;			/* :110-111 */
    return L;
}

size_t orc_tx_nsamples(const orc_tx_config *cfg, size_t nwords)
{
    struct tx_lengths L = tx_lengths(cfg);
    size_t per_frame = (cfg->nstartbits > 0 ? L.start : 0)
	    + (size_t)cfg->n_data_bits * L.bit + (cfg->nstopbits > 0 ? L.stop : 0);
    if (nwords == 0)
	return 0;
    return (size_t)cfg->leader_bits * L.bit + (cfg->do_tx_sync_bytes + nwords) * per_frame
	    + (size_t)cfg->trailer_bits * L.bit;
}

static void tx_frame(struct tonegen *g, const orc_tx_config *cfg, struct tx_lengths L,
	unsigned bits, int msb_first)
{
    /* fsk_transmit_frame, src/minimodem.c:81-112 */
    if (cfg->nstartbits > 0)
	tonegen_tone(g, cfg->invert_start_stop ? cfg->f_mark : cfg->f_space, L.start);
    for (unsigned i = 0; i < cfg->n_data_bits; i++) {
	unsigned bit = msb_first ? (bits >> (cfg->n_data_bits - i - 1)) & 1 : (bits >> i) & 1;
	tonegen_tone(g, bit == 1 ? cfg->f_mark : cfg->f_space, L.bit);
    }
    if (cfg->nstopbits > 0)
	tonegen_tone(g, cfg->invert_start_stop ? cfg->f_space : cfg->f_mark, L.stop);
}

size_t orc_tx_words(const orc_tx_config *cfg, const unsigned *words, size_t nwords,
	float *out, size_t out_cap)
{
    /* fsk_transmit_stdin, src/minimodem.c:114-250, non-interactive, input never
     * blocks: leader, sync preamble, frames, then the trailer from
     * tx_stop_transmit_sighandler (:59-74). */
    if (nwords == 0)
	return 0;				/* :246-247 nothing transmitted */
    struct tonegen g;
    tonegen_init(&g, cfg, out, out_cap);
    struct tx_lengths L = tx_lengths(cfg);
    for (int j = 0; j < cfg->leader_bits; j++)	/* :211-212 */
	tonegen_tone(&g, cfg->invert_start_stop ? cfg->f_space : cfg->f_mark, L.bit);
    for (unsigned j = 0; j < cfg->do_tx_sync_bytes; j++)	/* :218-221 */
	tx_frame(&g, cfg, L, cfg->sync_byte, 0);
    for (size_t w = 0; w < nwords; w++)		/* :225-228 */
	tx_frame(&g, cfg, L, words[w], cfg->msb_first);
    for (int j = 0; j < cfg->trailer_bits; j++)	/* :65-66 */
	tonegen_tone(&g, cfg->f_mark, L.bit);
    size_t n = g.n;
    tonegen_free(&g);
    return n;
}
