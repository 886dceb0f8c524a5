/*
 * fsk_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement ("port") of the
 * reference's FSK demodulation hot path, used ONLY as the parity checker in
 * tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline leg.
 * The product (minimodem_b200/, include/) never includes, links or calls it.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle_*.py) against
 *   - oracle/_ref/libfsk_ref.so  = the UNMODIFIED reference src/fsk.c compiled
 *     in place with the FFT stand-in (oracle/shim), call by call, and
 *   - tests/golden/ (npz files)  = traces of the unmodified reference CLI
 *     (oracle/_ref/minimodem_ref_trace) running its own tests/NN-name.test vectors.
 *
 * Third-party arithmetic: the reference obtains its two tone bins from FFTW3
 * single precision (fftw3f, version un-pinned: configure.ac:16; absent from
 * /root/reference and this image).  Only two bins of a mathematically defined
 * DFT are consumed (src/fsk.c:157-159), so this restatement evaluates exactly
 * those two bins by direct summation in double precision and rounds to float;
 * any correct r2c FFT agrees with that to float rounding (~1e-7 relative).
 *
 * All reference citations are path:line under /root/reference/.
 */
#ifndef FSK_ORACLE_H
#define FSK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plan: restates src/fsk.c:33-95 (fsk_plan_new) ---------------------- */
typedef struct orc_plan {
    float sample_rate, f_mark, f_space, band_width;
    int fftsize;
    unsigned nbands, b_mark, b_space;
    /* scratch: twiddle cache for the current bit_nsamples */
    unsigned tw_n;
    double *tw;		/* [tw_n][4] = cos_m, sin_m, cos_s, sin_s */
} orc_plan;

/* returns 0 and fills *p, or -1 (EINVAL case of src/fsk.c:58-64) */
int  orc_plan_init(orc_plan *p, float sample_rate, float f_mark, float f_space,
		float filter_bw);
void orc_plan_free(orc_plan *p);

/* src/fsk.c:107-174: one bit window.  mags = {mag_mark, mag_space} */
void orc_bit_mags(orc_plan *p, const float *samples, unsigned bit_nsamples,
		float *mag_mark, float *mag_space);

/* src/fsk.c:178-446 (CONFIDENCE_ALGO 6).  Optional per-bit outputs
 * (sig/noise/value arrays of n_bits entries) may be NULL. */
float orc_frame_analyze(orc_plan *p, const float *samples, float samples_per_bit,
		int n_bits, const char *expect_bits,
		unsigned long long *bits_out, float *ampl_out,
		float *bit_sig, float *bit_noise, unsigned *bit_val);

/* src/fsk.c:449-538 */
float orc_find_frame(orc_plan *p, const float *samples, unsigned frame_nsamples,
		unsigned try_first_sample, unsigned try_max_nsamples,
		unsigned try_step_nsamples, float try_confidence_search_limit,
		const char *expect_bits_string,
		unsigned long long *bits_out, float *ampl_out,
		unsigned *frame_start_out);

/* ---- rx loop: restates src/minimodem.c:1034-1463 ------------------------ */

/* What main() derives from the command line before the loop
 * (src/minimodem.c:819-965); the caller fills this in. */
typedef struct orc_rx_config {
    float sample_rate;		/* :534, :1029 */
    float data_rate;		/* bfsk_data_rate */
    float f_mark, f_space;	/* after defaults :900-934 and inversion :953 */
    float band_width;		/* after clamp :960 */
    unsigned n_data_bits;
    int nstartbits;
    float nstopbits;
    int invert_start_stop;
    int msb_first;
    int do_rx_sync;
    unsigned long long sync_byte;	/* (unsigned long long)-1 = none */
    float confidence_threshold;		/* :513 default 1.5 */
    float confidence_search_limit;	/* :523 default 2.3, sanitised :964 */
    const char *expect_data_string;	/* NULL = build (uic supplies its own :875) */
} orc_rx_config;

/* Everything the loop derives once (src/minimodem.c:1037-1131) */
typedef struct orc_rx_derived {
    float nsamples_per_bit;
    unsigned frame_n_bits;		/* :943 truncating */
    unsigned frame_nsamples;		/* :1113 */
    unsigned expect_n_bits;
    unsigned expect_nsamples;		/* :1131 */
    unsigned nsamples_overscan;		/* :1105-1108 */
    size_t samplebuf_size;		/* :1063-1070 */
    char expect_data[68];
    char expect_sync[68];
} orc_rx_derived;

int orc_build_expect_bits_string(char *out, int nstartbits, int n_data_bits,
		float nstopbits, int invert_start_stop, int use_expect_bits,
		unsigned long long expect_bits);	/* :442-487 */

void orc_rx_derive(const orc_rx_config *cfg, orc_rx_derived *d);

/* One record per frame that passed the squelch (reached :1391). */
typedef struct orc_rx_frame {
    unsigned long long bits;	/* raw fsk_find_frame bits (before :1415 chop) */
    float confidence;		/* coarse confidence (:1265; refine does not replace it) */
    float amplitude;
    unsigned frame_start;	/* within-buffer start (after refine) */
    unsigned acquired;		/* 1 if this frame acquired carrier (:1332-1355) */
    unsigned long long pos;	/* absolute sample index of samplebuf[0] */
} orc_rx_frame;

/* One record per carrier drop / end-of-stream report (report_no_carrier :253) */
typedef struct orc_rx_report {
    unsigned nframes_decoded;
    unsigned long long carrier_nsamples;
    float confidence_total;
    float amplitude_total;
    unsigned after_frame;	/* number of frame records emitted before it */
} orc_rx_report;

/* one record per fsk_find_frame call, for call-by-call comparison */
typedef struct orc_rx_call {
    unsigned frame_nsamples, try_first, try_max, try_step;
    float limit;
    int use_sync_string;
    float confidence;
    unsigned long long bits;
    float ampl;
    unsigned frame_start;
    unsigned long long pos;
} orc_rx_call;

enum { ORC_RX_LITERAL = 0,	/* emulate the sample ring incl. stale tail */
       ORC_RX_FLAT = 1 };	/* flat buffer, zeros past the end (batched-API semantic) */

typedef float (*orc_find_frame_fn)(void *ctx, const float *samples,
	unsigned frame_nsamples, unsigned try_first, unsigned try_max,
	unsigned try_step, float limit, const char *expect,
	unsigned long long *bits, float *ampl, unsigned *frame_start);

typedef struct orc_rx_result {
    orc_rx_frame *frames;   size_t nframes,  cap_frames;
    orc_rx_report *reports; size_t nreports, cap_reports;
    orc_rx_call *calls;     size_t ncalls,   cap_calls;	/* only if want_calls */
    unsigned long long n_find_frame_calls;
} orc_rx_result;

/* Runs the whole rx loop over samples[0..nsamples).  rxnoise reproduces the
 * --Xrxnoise quirk (src/simpleaudio-sndfile.c:64-70: a constant -rxnoise DC
 * offset applied to every *requested* frame of each read).  find_frame/ctx
 * NULL = orc_find_frame on an internal plan.  rx_one stops after the first
 * carrier drop (:1310).  *res must be zero-initialised before its first use
 * (its buffers are reused by later calls).  Returns 0, or -1 on bad config. */
int orc_rx_run(const orc_rx_config *cfg, const float *samples, size_t nsamples,
		int mode, float rxnoise, int rx_one, int want_calls,
		orc_find_frame_fn find_frame, void *ctx,
		orc_rx_result *res);
void orc_rx_result_free(orc_rx_result *res);

/* :1415-1428: frame bits -> data bits handed to the databits decoder */
unsigned long long orc_rx_databits(const orc_rx_config *cfg, unsigned long long bits);

/* Multi-threaded driver for the CPU baseline: demodulates nstreams flat
 * streams (row stride in floats) with nthreads threads, one plan per thread
 * (plans are not re-entrant, src/fsk.h:42-44), streams round-robin.  With
 * plan_new/find_frame/plan_destroy NULL the oracle's own two-bin analyzer is
 * used ("port"); pass the three entry points of oracle/_ref/libfsk_ref.so to
 * time the unmodified reference src/fsk.c behind the same rx loop
 * ("reference").  Returns total frames decoded. */
typedef void *(*orc_plan_new_fn)(float, float, float, float);
typedef void (*orc_plan_destroy_fn)(void *);
unsigned long long orc_rx_many(const orc_rx_config *cfg, const float *samples,
		size_t nstreams, size_t stride, size_t nsamples, int nthreads,
		orc_plan_new_fn plan_new, orc_find_frame_fn find_frame,
		orc_plan_destroy_fn plan_destroy,
		unsigned *frames_per_stream, unsigned long long *bits_xor_per_stream);

/* Persistent worker pool for the CPU timing arms (bench.py): pinned workers created once, one
 * plan per worker built once, contiguous stream blocks copied into pool memory by the worker
 * that will demodulate them (first touch), pass time taken inside around the start/stop barriers. */
typedef struct orc_pool orc_pool;
orc_pool *orc_pool_new(const orc_rx_config *cfg, int nthreads, orc_plan_new_fn plan_new,
		orc_find_frame_fn find_frame, orc_plan_destroy_fn plan_destroy);
int orc_pool_load(orc_pool *p, const float *samples, size_t nstreams, size_t src_stride, size_t nsamples);
double orc_pool_run(orc_pool *p, unsigned long long *total, unsigned *frames_per_stream,
		unsigned long long *bits_xor_per_stream);
void orc_pool_free(orc_pool *p);

/* ---- tx: restates src/minimodem.c:81-250 + src/simple-tone-generator.c -- */
typedef struct orc_tx_config {
    float sample_rate;		/* unsigned in the reference, :534 */
    float data_rate;
    float f_mark, f_space;
    unsigned n_data_bits;
    float nstartbits, nstopbits;
    int invert_start_stop, msb_first;
    unsigned do_tx_sync_bytes;
    unsigned sync_byte;
    int leader_bits, trailer_bits;	/* :51-52, :950 */
    float amplitude;			/* --volume, :536 */
    unsigned sin_table_len;		/* --lut, :538 default 4096; 0 = sinf */
    int s16;				/* 1: S16 samples (read back as v/32768), 0: float */
} orc_tx_config;

/* number of samples orc_tx_words will produce */
size_t orc_tx_nsamples(const orc_tx_config *cfg, size_t nwords);
/* words[] = output of the databits encoder, one per frame.  Returns samples written. */
size_t orc_tx_words(const orc_tx_config *cfg, const unsigned *words, size_t nwords,
		float *out, size_t out_cap);

#ifdef __cplusplus
}
#endif
#endif
