/*
 * fsk_b200_internal.h -- private interface between the C host layer
 * (fsk_b200_host.c) and the CUDA translation unit (fsk_b200_kernels.cu).
 * Not installed; include/fsk_b200.h is the public ABI.
 */
#ifndef FSK_B200_INTERNAL_H
#define FSK_B200_INTERNAL_H

#include "fsk_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Geometry of one frame candidate as the kernels consume it (passed by value
 * as a kernel parameter, so changing it costs nothing). */
typedef struct fsk_b200_geom {
    unsigned int n_bits;
    unsigned int bit_nsamples;
    unsigned int span;			/* bit_begin[n_bits-1] + bit_nsamples */
    unsigned int lanes_per_window;	/* filled in at launch */
    float	mag_scalar;		/* 2.0f / bit_nsamples, src/fsk.c:132 */
    float	eps_unscaled;		/* FLT_EPSILON / mag_scalar: the :279 threshold before scaling */
    float	inv_n_bits;		/* 1.0f / n_bits */
    unsigned int bit_begin[FSK_B200_MAX_BITS];
    unsigned char expect[2][FSK_B200_MAX_BITS];	/* [0]=data [1]=sync; 0,1 or 2 ('d') */
} fsk_b200_geom;

/* rx-loop constants for the rx kernel (by value as well) */
typedef struct fsk_b200_loopc {
    unsigned int frame_nsamples, expect_nsamples, nsamples_overscan;
    unsigned int try_max_nocarrier, try_max_carrier;
    float	confidence_threshold, confidence_search_limit;
} fsk_b200_loopc;

void fsk_b200_set_error(const char *fmt, ...);

/* host-side pure derivations (fsk_b200_host.c) */
int fsk_b200_geom_from(unsigned int frame_nsamples, const char *expect_data,
	const char *expect_sync, fsk_b200_geom *g);

/* CUDA side (fsk_b200_kernels.cu) */
int  fsk_b200_cuda_device_ok(void);
void *fsk_b200_cuda_engine_new(void);
void fsk_b200_cuda_engine_destroy(void *ce);
/* (re)build the twiddle table for (fftsize, b_mark, b_space, bit_nsamples) */
int  fsk_b200_cuda_set_table(void *ce, int fftsize, unsigned int b_mark, unsigned int b_space,
	unsigned int bit_nsamples);
int  fsk_b200_cuda_tune(void *ce, int lanes_per_stream, int warps_per_block, int ring_floats);
int  fsk_b200_cuda_find_frame_batch(void *ce, const fsk_b200_geom *g, const float *samples,
	size_t nstreams, size_t stride, const uint32_t *offset, const uint32_t *nvalid,
	const uint32_t *try_first, const uint32_t *try_max, const uint32_t *try_step,
	const float *limit, const uint8_t *expect_sel, fsk_b200_frame *frames, void *stream);
int  fsk_b200_cuda_rx_batch(void *ce, const fsk_b200_geom *g, const fsk_b200_loopc *lc,
	const float *samples, size_t nstreams, size_t stride, const uint32_t *nsamples,
	uint32_t nsamples_all, fsk_b200_frame *frames, uint32_t max_frames,
	fsk_b200_stream_state *states, void *stream);
int  fsk_b200_cuda_rx_batch_host(void *ce, const fsk_b200_geom *g, const fsk_b200_loopc *lc,
	const float *host_samples, size_t nstreams, size_t stride, uint32_t nsamples_all,
	fsk_b200_frame *host_frames, uint32_t max_frames, fsk_b200_stream_state *host_states);
/* single-stream helpers behind the drop-in API: host buffers in, host results out */
int  fsk_b200_cuda_find_frame_one(void *ce, const fsk_b200_geom *g, const float *host_samples,
	unsigned int nfloats, unsigned int try_first, unsigned int try_max, unsigned int try_step,
	float limit, fsk_b200_frame *out);
int  fsk_b200_cuda_band_mags(void *ce, int fftsize, const float *host_samples,
	unsigned int nsamples, unsigned int nbands, float *host_mags);
int  fsk_b200_cuda_detect_carrier_batch(int fftsize, const float *samples, size_t nstreams, size_t stride,
	const uint32_t *offset, uint32_t nsamples, float min_mag_threshold, int32_t *out_band, void *stream);
int  fsk_b200_cuda_stream_push(float *samples, size_t nstreams, size_t stride, uint32_t *fill,
	fsk_b200_stream_state *states, const float *chunk, size_t chunk_stride, const uint32_t *chunk_len,
	uint32_t chunk_len_all, uint32_t *dropped, void *stream);
int  fsk_b200_cuda_tx_batch(const fsk_b200_tx_config *cfg, const float *sin_table,
	uint32_t table_len, const uint32_t *words, uint32_t nwords, const uint32_t *lead_in,
	float *samples_out, size_t nstreams, size_t stride, uint32_t nsamples_out, void *stream);
int  fsk_b200_cuda_s16_to_f32(const int16_t *src, float *dst, size_t nstreams, size_t stride, void *stream);
int  fsk_b200_cuda_rx_batch_host_s16(void *ce, const fsk_b200_geom *g, const fsk_b200_loopc *lc,
	const int16_t *host_samples, size_t nstreams, size_t stride, uint32_t nsamples_all,
	fsk_b200_frame *host_frames, uint32_t max_frames, fsk_b200_stream_state *host_states);
int  fsk_b200_cuda_decode(int kind, unsigned shift, unsigned n_data_bits, int msb_first, int do_rx_sync,
	unsigned long long sync_byte, const fsk_b200_frame *frames, const fsk_b200_stream_state *states,
	size_t nstreams, uint32_t max_frames, fsk_b200_decoder_state *dstates, uint8_t *out,
	uint32_t out_stride, uint32_t *out_count, void *stream);
unsigned long long fsk_b200_cuda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
