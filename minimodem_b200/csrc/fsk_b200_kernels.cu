/*
 * fsk_b200_kernels.cu -- CUDA side of the B200 FSK engine (sm_100a).
 *
 * Work decomposition (DESIGN.md has the derivation):
 *   - one GROUP of G lanes (G = 2..32, a power of two) owns one audio stream;
 *     a warp therefore runs 32/G streams side by side;
 *   - the stream's samples live in a per-stream shared-memory RING indexed by
 *     the absolute sample index (ring[i & mask]); every input sample is fetched
 *     from HBM exactly once, 16 bytes per lane, and stays there while the
 *     candidate frame positions that cover it are searched;
 *   - inside a frame candidate the lanes of a group split the bit windows
 *     (and, when G > n_bits, the samples of a window) and correlate each window
 *     against the mark and space tones at the FFT-bin centre frequencies
 *     (exp(-2 pi i k n / fftsize), k = b_mark, b_space) -- the two bins the
 *     reference reads out of a full FFT (src/fsk.c:157-159);
 *   - confidence (src/fsk.c:271-342), the zig-zag search with early-out
 *     (src/fsk.c:477-502) and the rx-loop state machine
 *     (src/minimodem.c:1229-1407) run per group in registers, in the
 *     reference's order of floating-point operations.
 *
 * Compiled with -fmad=false: every a*b+c below is either an explicit fmaf()
 * (the correlation sums) or two separately rounded operations, as in the
 * reference's x86-64 build.
 */
#include <cuda_runtime.h>
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsk_b200_internal.h"

#define FSK_FLT_EPSILON 1.1920928955078125e-07f
#define ACC_BLOCK 64u	/* fp32 partial sums are folded into fp64 every ACC_BLOCK terms */

static unsigned long long g_launches;

#define CUDA_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fsk_b200_set_error("%s: %s", #call, cudaGetErrorString(e_)); return -EIO; } } while (0)

/* ------------------------------------------------------------------------ */
/* sample sources                                                           */
/* ------------------------------------------------------------------------ */

/* shared-memory ring addressed by absolute sample index */
struct RingSrc {
    const float *ring;
    unsigned mask;
    __device__ __forceinline__ float operator()(unsigned i) const { return ring[i & mask]; }
};

/* straight from global memory, zero beyond the valid length (windows that do
 * not fit the ring: very low baud rates) */
struct GlobalSrc {
    const float *x;
    unsigned n;
    __device__ __forceinline__ float operator()(unsigned i) const { return i < n ? __ldg(x + i) : 0.0f; }
};

/* ------------------------------------------------------------------------ */
/* frame analysis: src/fsk.c:178-446 for one candidate start                */
/* ------------------------------------------------------------------------ */

template <int G, class Src>
__device__ __forceinline__ float frame_analyze(const Src &src, unsigned t0,
	const fsk_b200_geom &geo, int sel, const float4 *__restrict__ tw, float2 *scr,
	unsigned g, unsigned gmask, unsigned long long &bits_out, float &ampl_out)
{
    const unsigned N = geo.bit_nsamples, nb = geo.n_bits, L = geo.lanes_per_window;
    const unsigned wpp = G / L;			/* windows analysed per pass */
    const unsigned part = g & (L - 1), wslot = g / L;
    bool mismatch = false;

    __syncwarp(gmask);				/* previous readers of scr are done */
    for (unsigned w0 = 0; w0 < nb; w0 += wpp) {
	const unsigned w = w0 + wslot;
	const bool active = w < nb;
	float rm = 0.f, im = 0.f, rs = 0.f, is = 0.f;
	if (active) {
	    const unsigned base = t0 + geo.bit_begin[w];
	    if (N <= ACC_BLOCK * L) {
#pragma unroll 4
		for (unsigned n = part; n < N; n += L) {
		    const float x = src(base + n);
		    const float4 c = tw[n];
		    rm = fmaf(x, c.x, rm);
		    im = fmaf(x, c.y, im);
		    rs = fmaf(x, c.z, rs);
		    is = fmaf(x, c.w, is);
		}
	    } else {
		/* long windows: bounded fp32 partial sums folded into fp64 */
		double drm = 0., dim = 0., drs = 0., dis = 0.;
		for (unsigned n0 = part; n0 < N; n0 += ACC_BLOCK * L) {
		    const unsigned nend = min(N, n0 + ACC_BLOCK * L);
		    float prm = 0.f, pim = 0.f, prs = 0.f, pis = 0.f;
#pragma unroll 4
		    for (unsigned n = n0; n < nend; n += L) {
			const float x = src(base + n);
			const float4 c = tw[n];
			prm = fmaf(x, c.x, prm);
			pim = fmaf(x, c.y, pim);
			prs = fmaf(x, c.z, prs);
			pis = fmaf(x, c.w, pis);
		    }
		    drm += prm; dim += pim; drs += prs; dis += pis;
		}
		rm = (float)drm; im = (float)dim; rs = (float)drs; is = (float)dis;
	    }
	}
	for (unsigned o = L >> 1; o; o >>= 1) {
	    rm += __shfl_xor_sync(gmask, rm, o);
	    im += __shfl_xor_sync(gmask, im, o);
	    rs += __shfl_xor_sync(gmask, rs, o);
	    is += __shfl_xor_sync(gmask, is, o);
	}
	if (active && part == 0) {
	    /* band_mag, src/fsk.c:107-114, then the decision at :158-169 */
	    float mag_mark = sqrtf(rm * rm + im * im) * geo.mag_scalar;
	    float mag_space = sqrtf(rs * rs + is * is) * geo.mag_scalar;
	    /* The reference drops off-tone magnitudes <= FLT_EPSILON from the noise sum
	     * (src/fsk.c:279) so that exactly periodic tones give confidence = inf.  fp32
	     * accumulation is good to ~2e-7 of the signal, not enough to classify a
	     * magnitude that close to FLT_EPSILON: such (rare: synthetic, orthogonal-tone)
	     * windows are re-summed in fp64, where float*float products are exact. */
	    {
		const float lo = fminf(mag_mark, mag_space), hi = fmaxf(mag_mark, mag_space);
		if (lo < FSK_FLT_EPSILON + 2e-6f * hi) {
		    const unsigned base = t0 + geo.bit_begin[w];
		    double drm = 0., dim = 0., drs = 0., dis = 0.;
		    for (unsigned n = 0; n < N; n++) {
			const double x = (double)src(base + n);
			const float4 c = tw[n];
			drm = fma(x, (double)c.x, drm);
			dim = fma(x, (double)c.y, dim);
			drs = fma(x, (double)c.z, drs);
			dis = fma(x, (double)c.w, dis);
		    }
		    const float frm = (float)drm, fim = (float)dim, frs = (float)drs, fis = (float)dis;
		    mag_mark = sqrtf(frm * frm + fim * fim) * geo.mag_scalar;
		    mag_space = sqrtf(frs * frs + fis * fis) * geo.mag_scalar;
		}
	    }
	    const bool one = mag_mark > mag_space;		/* strict: tie -> space */
	    const float sig = one ? mag_mark : mag_space;
	    const float noise = one ? mag_space : mag_mark;
	    /* the bit value rides in the sign of the (non-negative) noise magnitude */
	    scr[w] = make_float2(sig, one ? -noise : noise);
	    const unsigned e = geo.expect[sel][w];
	    if (e != 2u && e != (one ? 1u : 0u))
		mismatch = true;			/* pass 1 reject, src/fsk.c:211-212 */
	}
    }
    __syncwarp(gmask);
    if (__any_sync(gmask, mismatch)) {
	bits_out = 0;
	ampl_out = 0.f;
	return 0.f;
    }

    /* src/fsk.c:271-301, bit index ascending, one rounding per operation */
    float total_sig = 0.f, total_noise = 0.f, avg_mark = 0.f, avg_space = 0.f;
    unsigned n_mark = 0, n_space = 0;
    unsigned long long bits = 0;
    for (unsigned b = 0; b < nb; b++) {
	const float2 v = scr[b];
	const float noise = fabsf(v.y);
	total_sig += v.x;
	if (noise > FSK_FLT_EPSILON)
	    total_noise += noise;
	if (signbit(v.y)) {
	    avg_mark += v.x;
	    n_mark++;
	    bits |= 1ull << b;
	} else {
	    avg_space += v.x;
	    n_space++;
	}
    }
    const float snr = total_sig / total_noise;		/* may be +inf */
    const float avg_bit_sig = total_sig / (float)(int)nb;
    if (n_mark)
	avg_mark = avg_mark / (float)n_mark;
    if (n_space)
	avg_space = avg_space / (float)n_space;

    /* divergence terms (src/fsk.c:305-311) computed by the window owners ... */
    __syncwarp(gmask);
    for (unsigned w0 = 0; w0 < nb; w0 += wpp) {
	const unsigned w = w0 + wslot;
	if (w < nb && part == 0) {
	    const float2 v = scr[w];
	    const float other = signbit(v.y) ? avg_mark : avg_space;
	    scr[w].x = fabsf(v.x - other) / other;
	}
    }
    __syncwarp(gmask);
    /* ... and summed in bit order */
    float divergence = 0.f;
    for (unsigned b = 0; b < nb; b++)
	divergence += scr[b].x;
    divergence *= 2.f;
    divergence = divergence / (float)(int)nb;

    bits_out = bits;
    ampl_out = avg_bit_sig;
    return snr * (1.0f - divergence);			/* src/fsk.c:336 */
}

/* ------------------------------------------------------------------------ */
/* frame search: src/fsk.c:449-538                                          */
/* ------------------------------------------------------------------------ */

template <int G, class Src>
__device__ __forceinline__ float find_frame(const Src &src, unsigned base,
	const fsk_b200_geom &geo, int sel, const float4 *__restrict__ tw, float2 *scr,
	unsigned g, unsigned gmask, unsigned try_first, unsigned try_max, unsigned try_step,
	float limit, unsigned long long &best_bits, float &best_a, unsigned &best_t)
{
    float best_c = 0.f;
    best_t = 0;
    best_a = 0.f;
    best_bits = 0;
    for (int j = 0;; j++) {
	const int up = (j & 1) ? 1 : -1;
	const int t = (int)try_first + up * ((j + 1) / 2) * (int)try_step;
	if (t >= (int)try_max)
	    break;
	if (t < 0)
	    continue;
	unsigned long long bits;
	float a;
	const float c = frame_analyze<G, Src>(src, base + (unsigned)t, geo, sel, tw, scr, g, gmask,
		bits, a);
	if (best_c < c) {			/* NaN and negatives never win */
	    best_t = (unsigned)t;
	    best_c = c;
	    best_a = a;
	    best_bits = bits;
	    if (best_c >= limit)
		break;				/* first to reach the limit wins */
	}
    }
    return best_c;
}

/* ------------------------------------------------------------------------ */
/* ring fill: HBM -> shared memory, 16 bytes per lane, each sample once     */
/* ------------------------------------------------------------------------ */

template <int G>
__device__ __forceinline__ void ring_fill(float *ring, unsigned mask, const float *__restrict__ x,
	unsigned n, unsigned from, unsigned to, unsigned g)
{
    /* from, to multiples of 4; x 16-byte aligned */
    for (unsigned i = from + 4u * g; i < to; i += 4u * G) {
	float4 v;
	if (i + 4u <= n) {
	    v = __ldg(reinterpret_cast<const float4 *>(x + i));
	} else {
	    v.x = (i + 0u < n) ? __ldg(x + i + 0) : 0.f;
	    v.y = (i + 1u < n) ? __ldg(x + i + 1) : 0.f;
	    v.z = (i + 2u < n) ? __ldg(x + i + 2) : 0.f;
	    v.w = (i + 3u < n) ? __ldg(x + i + 3) : 0.f;
	}
	*reinterpret_cast<float4 *>(ring + (i & mask)) = v;
    }
}

__device__ __forceinline__ void store_frame(fsk_b200_frame *f, unsigned long long bits, float conf,
	float ampl, unsigned start)
{
    uint32_t *p = reinterpret_cast<uint32_t *>(f);
    p[0] = (uint32_t)bits;
    p[1] = (uint32_t)(bits >> 32);
    p[2] = __float_as_uint(conf);
    p[3] = __float_as_uint(ampl);
    p[4] = start;
}

/* shared memory carve-up common to both kernels */
struct Smem {
    const float4 *tw;
    float *ring;
    float2 *scr;
};

template <int G>
__device__ __forceinline__ Smem carve(float4 *smem, const fsk_b200_geom &geo,
	const float4 *__restrict__ tw_global, unsigned tw_in_smem, unsigned ring_floats,
	unsigned warps_per_block)
{
    const unsigned N = geo.bit_nsamples;
    Smem s;
    float4 *p = smem;
    if (tw_in_smem) {
	for (unsigned i = threadIdx.x; i < N; i += blockDim.x)
	    p[i] = tw_global[i];
	s.tw = p;
	p += N;
    } else {
	s.tw = tw_global;
    }
    const unsigned spw = 32 / G;
    const unsigned warp = threadIdx.x >> 5, sidx = (threadIdx.x & 31) / G;
    const unsigned slot = warp * spw + sidx;
    float *rings = reinterpret_cast<float *>(p);
    s.ring = rings + (size_t)slot * ring_floats;
    float2 *scrs = reinterpret_cast<float2 *>(rings + (size_t)warps_per_block * spw * ring_floats);
    s.scr = scrs + (size_t)slot * geo.n_bits;
    __syncthreads();
    return s;
}

/* ------------------------------------------------------------------------ */
/* K1: batched fsk_find_frame                                               */
/* ------------------------------------------------------------------------ */

template <int G>
__global__ void __launch_bounds__(256)
k_find_frame(const __grid_constant__ fsk_b200_geom geo, const float4 *__restrict__ tw_global,
	unsigned tw_in_smem, unsigned ring_floats, const float *__restrict__ samples,
	unsigned nstreams, size_t stride, const uint32_t *__restrict__ offset,
	const uint32_t *__restrict__ nvalid, const uint32_t *__restrict__ try_first,
	const uint32_t *__restrict__ try_max, const uint32_t *__restrict__ try_step,
	const float *__restrict__ limit, const uint8_t *__restrict__ expect_sel,
	fsk_b200_frame *__restrict__ frames)
{
    extern __shared__ float4 smem4[];
    const unsigned wpb = blockDim.x >> 5;
    const Smem sm = carve<G>(smem4, geo, tw_global, tw_in_smem, ring_floats, wpb);
    const unsigned lane = threadIdx.x & 31, g = lane % G, sidx = lane / G, spw = 32 / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sidx * G));
    const unsigned warp = threadIdx.x >> 5;
    const unsigned mask = ring_floats ? ring_floats - 1u : 0u;

    for (unsigned s = (blockIdx.x * wpb + warp) * spw + sidx; s < nstreams;
	    s += gridDim.x * wpb * spw) {
	const float *x = samples + (size_t)s * stride;
	const unsigned off = offset ? offset[s] : 0u;
	const unsigned n = nvalid[s];
	const unsigned tmax = try_max[s];
	unsigned tstep = try_step[s];
	if (tstep == 0)
	    tstep = 1;
	const int sel = expect_sel ? (expect_sel[s] ? 1 : 0) : 0;
	unsigned long long bits = 0;
	float ampl = 0.f, conf = 0.f;
	unsigned start = 0;
	if (tmax) {
	    const unsigned need_end = off + tmax - 1u + geo.span;
	    const unsigned from = off & ~3u;
	    const unsigned to = (need_end + 3u) & ~3u;
	    if (ring_floats && to - from <= ring_floats) {
		__syncwarp(gmask);
		ring_fill<G>(sm.ring, mask, x, n, from, to, g);
		__syncwarp(gmask);
		RingSrc src = { sm.ring, mask };
		conf = find_frame<G, RingSrc>(src, off, geo, sel, sm.tw, sm.scr, g, gmask,
			try_first[s], tmax, tstep, limit[s], bits, ampl, start);
	    } else {
		GlobalSrc src = { x, n };
		conf = find_frame<G, GlobalSrc>(src, off, geo, sel, sm.tw, sm.scr, g, gmask,
			try_first[s], tmax, tstep, limit[s], bits, ampl, start);
	    }
	}
	if (g == 0)
	    store_frame(frames + s, bits, conf, ampl, start);
    }
}

/* ------------------------------------------------------------------------ */
/* K2: the rx loop (src/minimodem.c:1137-1463) for whole streams            */
/* ------------------------------------------------------------------------ */

template <int G, bool RING>
__global__ void __launch_bounds__(256)
k_rx(const __grid_constant__ fsk_b200_geom geo, const __grid_constant__ fsk_b200_loopc lc,
	const float4 *__restrict__ tw_global, unsigned tw_in_smem, unsigned ring_floats,
	const float *__restrict__ samples, unsigned nstreams, size_t stride,
	const uint32_t *__restrict__ nsamples, uint32_t nsamples_all,
	fsk_b200_frame *__restrict__ frames, uint32_t max_frames,
	fsk_b200_stream_state *__restrict__ states)
{
    extern __shared__ float4 smem4[];
    const unsigned wpb = blockDim.x >> 5;
    const Smem sm = carve<G>(smem4, geo, tw_global, tw_in_smem, ring_floats, wpb);
    const unsigned lane = threadIdx.x & 31, g = lane % G, sidx = lane / G, spw = 32 / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sidx * G));
    const unsigned warp = threadIdx.x >> 5;
    const unsigned mask = ring_floats ? ring_floats - 1u : 0u;

    for (unsigned s = (blockIdx.x * wpb + warp) * spw + sidx; s < nstreams;
	    s += gridDim.x * wpb * spw) {
	fsk_b200_stream_state st = states[s];
	if (st.done)
	    continue;
	const float *x = samples + (size_t)s * stride;
	const unsigned n = nsamples ? nsamples[s] : nsamples_all;
	fsk_b200_frame *out = frames + (size_t)s * max_frames;

	unsigned pos = (unsigned)st.pos;
	unsigned nframes = st.nframes;
	unsigned carrier = st.carrier, noconfidence = st.noconfidence;
	float track_amplitude = st.track_amplitude, peak_confidence = st.peak_confidence;
	unsigned long long carrier_nsamples = st.carrier_nsamples;
	float confidence_total = st.confidence_total, amplitude_total = st.amplitude_total;
	unsigned nframes_decoded = st.nframes_decoded;
	unsigned done = 0;
	unsigned filled = pos & ~3u;		/* ring holds [filled_lo, filled) */
	__syncwarp(gmask);

	for (;;) {
	    if (pos >= n) { done = 1; break; }			/* :1176 */
	    const unsigned remaining = n - pos;
	    if (remaining < lc.expect_nsamples) { done = 1; break; }	/* :1229 */
	    if (nframes >= max_frames)
		break;						/* output full: resumable */

	    unsigned try_max = carrier ? lc.try_max_carrier : lc.try_max_nocarrier;	/* :1236-1241 */
	    unsigned try_step = try_max / 3u;			/* :1248-1251 */
	    if (try_step == 0)
		try_step = 1;
	    const unsigned try_first = carrier ? lc.nsamples_overscan : 0u;	/* :1263 */
	    const int sel = carrier ? 0 : 1;			/* :1270 data / sync string */

	    unsigned long long bits;
	    float amplitude, confidence;
	    unsigned frame_start;
	    unsigned long long bits2;
	    float amplitude2, confidence2 = 0.f;
	    unsigned frame_start2;
	    bool want_refine;

	    if (RING) {
		const unsigned to = (pos + try_max - 1u + geo.span + 3u) & ~3u;
		if (filled < (pos & ~3u))
		    filled = pos & ~3u;		/* skipped ahead of everything fetched */
		if (to > filled) {
		    ring_fill<G>(sm.ring, mask, x, n, filled, to, g);
		    filled = to;
		}
		__syncwarp(gmask);
	    }
	    const RingSrc rsrc = { sm.ring, mask };
	    const GlobalSrc gsrc = { x, n };

	    if (RING)
		confidence = find_frame<G, RingSrc>(rsrc, pos, geo, sel, sm.tw, sm.scr, g, gmask,
			try_first, try_max, try_step, lc.confidence_search_limit,
			bits, amplitude, frame_start);		/* :1265 */
	    else
		confidence = find_frame<G, GlobalSrc>(gsrc, pos, geo, sel, sm.tw, sm.scr, g, gmask,
			try_first, try_max, try_step, lc.confidence_search_limit,
			bits, amplitude, frame_start);

	    want_refine = false;
	    if (confidence < peak_confidence * 0.75f) {		/* :1278-1282 */
		want_refine = true;
		peak_confidence = 0.f;
	    }
	    if (amplitude < track_amplitude * 0.25f)		/* :1286 */
		confidence = 0.f;

	    unsigned advance;
	    if (confidence <= lc.confidence_threshold) {	/* :1292 */
		if (++noconfidence > 20u) {			/* :1295 */
		    if (carrier) {
			/* report_no_carrier(), :1299-1302, as a record */
			if (g == 0)
			    store_frame(out + nframes, carrier_nsamples, confidence_total,
				    amplitude_total, FSK_B200_FRAME_REPORT);
			nframes++;
			carrier = 0;				/* :1303-1308 */
			carrier_nsamples = 0;
			confidence_total = 0.f;
			amplitude_total = 0.f;
			nframes_decoded = 0;
			track_amplitude = 0.f;
		    }
		}
		advance = try_max;				/* :1318 */
	    } else {
		unsigned acquired = 0;
		carrier_nsamples += lc.frame_nsamples;		/* :1324 */
		if (carrier) {
		    carrier_nsamples += frame_start;		/* :1329-1330: the COARSE start */
		    carrier_nsamples -= lc.nsamples_overscan;
		} else {					/* :1332-1355 */
		    carrier = 1;
		    acquired = FSK_B200_FRAME_ACQUIRED;
		    want_refine = true;
		}
		if (want_refine && confidence < INFINITY && try_step > 1u) {	/* :1357-1389 */
		    try_step = try_max / 8u;
		    if (try_step == 0)
			try_step = 1;
		    /* `carrier` is 1 by now, so the data string is searched (:1378) */
		    if (RING)
			confidence2 = find_frame<G, RingSrc>(rsrc, pos, geo, 0, sm.tw, sm.scr, g,
				gmask, try_first, try_max, try_step, INFINITY,
				bits2, amplitude2, frame_start2);
		    else
			confidence2 = find_frame<G, GlobalSrc>(gsrc, pos, geo, 0, sm.tw, sm.scr, g,
				gmask, try_first, try_max, try_step, INFINITY,
				bits2, amplitude2, frame_start2);
		    if (confidence2 > confidence) {
			bits = bits2;
			amplitude = amplitude2;
			frame_start = frame_start2;
		    }
		}
		track_amplitude = (track_amplitude + amplitude) / 2.f;	/* :1391 */
		if (peak_confidence < confidence)
		    peak_confidence = confidence;
		confidence_total += confidence;			/* :1397-1400 */
		amplitude_total += amplitude;
		nframes_decoded++;
		noconfidence = 0;
		if (g == 0)
		    store_frame(out + nframes, bits, confidence, amplitude, frame_start | acquired);
		nframes++;
		advance = frame_start + lc.frame_nsamples - lc.nsamples_overscan;	/* :1407 */
	    }
	    if (advance > remaining) { done = 1; break; }	/* :1151 */
	    pos += advance;
	    if (RING)
		__syncwarp(gmask);	/* all reads of this window precede the next fill */
	}

	if (g == 0) {
	    st.pos = pos;
	    st.nframes = nframes;
	    st.carrier = carrier;
	    st.noconfidence = noconfidence;
	    st.track_amplitude = track_amplitude;
	    st.peak_confidence = peak_confidence;
	    st.carrier_nsamples = carrier_nsamples;
	    st.confidence_total = confidence_total;
	    st.amplitude_total = amplitude_total;
	    st.nframes_decoded = nframes_decoded;
	    st.done = done;
	    states[s] = st;
	}
	__syncwarp(gmask);
    }
}

/* ------------------------------------------------------------------------ */
/* full-spectrum magnitudes for fsk_detect_carrier (src/fsk.c:543-581)      */
/* ------------------------------------------------------------------------ */

__global__ void k_band_mags(const float *__restrict__ x, unsigned nsamples, int fftsize,
	unsigned nbands, float *__restrict__ mags)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbands)
	return;
    double re = 0., im = 0.;
    const unsigned F = (unsigned)fftsize;
    unsigned r = 0;				/* (k*n) mod F, kept exact in integers */
    for (unsigned n = 0; n < nsamples; n++) {
	float sn, cs;
	sincospif(2.0f * (float)r / (float)F, &sn, &cs);
	re += (double)(x[n] * cs);
	im += (double)(x[n] * sn);
	r += k;
	if (r >= F)
	    r -= F;
    }
    const float magscalar = 1.0f / ((float)nsamples / 2.0f);	/* src/fsk.c:553 */
    const float fr = (float)re, fi = (float)im;
    mags[k] = sqrtf(fr * fr + fi * fi) * magscalar;
}

/* ------------------------------------------------------------------------ */
/* transmitter signal model, one warp per stream                            */
/* (src/minimodem.c:81-250, src/simple-tone-generator.c:107-175)            */
/* ------------------------------------------------------------------------ */

struct TxLens { unsigned bit, start, stop, rate; };

__device__ __forceinline__ void tx_tone(float *o, unsigned &pos, unsigned cap, float &cphase,
	unsigned rate, float freq, unsigned dur, const float *__restrict__ lut, unsigned len,
	unsigned lane)
{
    const float wave = (float)rate / freq;			/* :116 */
    for (unsigned i = lane; i < dur; i += 32) {
	const float turns = (float)i / wave + cphase;		/* :121 */
	int t = (int)((float)len * turns + 0.5f);		/* :91 */
	t %= (int)len;
	if (pos + i < cap)
	    o[pos + i] = lut[t];
    }
    pos += dur;
    cphase = fmodf(cphase + (float)dur / wave, 1.0f);		/* :163 */
}

__device__ __forceinline__ void tx_frame(float *o, unsigned &pos, unsigned cap, float &cphase,
	const fsk_b200_tx_config &c, const TxLens &L, unsigned bits, int msb_first,
	const float *__restrict__ lut, unsigned len, unsigned lane)
{
    if (c.nstartbits > 0)
	tx_tone(o, pos, cap, cphase, L.rate, c.invert_start_stop ? c.f_mark : c.f_space, L.start,
		lut, len, lane);
    for (unsigned i = 0; i < c.n_data_bits; i++) {
	const unsigned bit = msb_first ? (bits >> (c.n_data_bits - i - 1)) & 1u : (bits >> i) & 1u;
	tx_tone(o, pos, cap, cphase, L.rate, bit ? c.f_mark : c.f_space, L.bit, lut, len, lane);
    }
    if (c.nstopbits > 0)
	tx_tone(o, pos, cap, cphase, L.rate, c.invert_start_stop ? c.f_space : c.f_mark, L.stop,
		lut, len, lane);
}

__global__ void k_tx(const __grid_constant__ fsk_b200_tx_config c, TxLens L,
	const float *__restrict__ lut, unsigned len, const uint32_t *__restrict__ words,
	unsigned nwords, const uint32_t *__restrict__ lead_in, float *__restrict__ out,
	unsigned nstreams, size_t stride, unsigned cap)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nstreams)
	return;
    float *o = out + (size_t)s * stride;
    unsigned pos = lead_in ? min(lead_in[s], cap) : 0u;
    for (unsigned i = lane; i < pos; i += 32)
	o[i] = 0.f;
    float cphase = 0.f;
    const uint32_t *w = words + (size_t)s * nwords;
    if (nwords) {
	for (int j = 0; j < c.leader_bits; j++)
	    tx_tone(o, pos, cap, cphase, L.rate, c.invert_start_stop ? c.f_space : c.f_mark, L.bit,
		    lut, len, lane);
	for (unsigned j = 0; j < c.do_tx_sync_bytes; j++)
	    tx_frame(o, pos, cap, cphase, c, L, c.sync_byte, 0, lut, len, lane);
	for (unsigned j = 0; j < nwords; j++)
	    tx_frame(o, pos, cap, cphase, c, L, w[j], c.msb_first, lut, len, lane);
	for (int j = 0; j < c.trailer_bits; j++)
	    tx_tone(o, pos, cap, cphase, L.rate, c.f_mark, L.bit, lut, len, lane);
    }
    for (unsigned i = min(pos, cap) + lane; i < cap; i += 32)
	o[i] = 0.f;
}

/* ======================================================================== */
/* host side of the CUDA translation unit                                   */
/* ======================================================================== */

struct CudaEngine {
    int device;
    int sm_count;
    int smem_optin;
    /* twiddle table */
    float4 *d_tw;
    unsigned tw_cap, tw_n;
    int tw_fftsize;
    unsigned tw_bm, tw_bs;
    /* tuning (0 = automatic) */
    int lanes, wpb, ring;
    /* single-stream staging */
    float *d_one;
    size_t d_one_cap;
    uint32_t *d_args;		/* offset, nvalid, first, max, step, limit(as float) */
    fsk_b200_frame *d_frame;
    float *d_mags;
    size_t d_mags_cap;
    /* host-batch slabs */
    float *d_slab[2];
    fsk_b200_frame *d_slab_frames[2];
    fsk_b200_stream_state *d_slab_states[2];
    size_t slab_streams, slab_stride, slab_max_frames;
    cudaStream_t st[2];
};

extern "C" unsigned long long fsk_b200_cuda_launch_count(void) { return g_launches; }

extern "C" int fsk_b200_cuda_device_ok(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
	(void)cudaGetLastError();
	return 0;
    }
    return 1;
}

extern "C" void *fsk_b200_cuda_engine_new(void)
{
    CudaEngine *ce = (CudaEngine *)calloc(1, sizeof(CudaEngine));
    if (!ce)
	return NULL;
    if (cudaGetDevice(&ce->device) != cudaSuccess
	    || cudaDeviceGetAttribute(&ce->sm_count, cudaDevAttrMultiProcessorCount, ce->device) != cudaSuccess
	    || cudaDeviceGetAttribute(&ce->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ce->device) != cudaSuccess) {
	fsk_b200_set_error("cuda engine: %s", cudaGetErrorString(cudaGetLastError()));
	free(ce);
	return NULL;
    }
    const char *e;
    if ((e = getenv("FSK_B200_LANES"))) ce->lanes = atoi(e);
    if ((e = getenv("FSK_B200_WPB"))) ce->wpb = atoi(e);
    if ((e = getenv("FSK_B200_RING"))) ce->ring = atoi(e);
    return ce;
}

extern "C" void fsk_b200_cuda_engine_destroy(void *p)
{
    CudaEngine *ce = (CudaEngine *)p;
    if (!ce)
	return;
    cudaFree(ce->d_tw);
    cudaFree(ce->d_one);
    cudaFree(ce->d_args);
    cudaFree(ce->d_frame);
    cudaFree(ce->d_mags);
    for (int i = 0; i < 2; i++) {
	cudaFree(ce->d_slab[i]);
	cudaFree(ce->d_slab_frames[i]);
	cudaFree(ce->d_slab_states[i]);
	if (ce->st[i])
	    cudaStreamDestroy(ce->st[i]);
    }
    free(ce);
}

extern "C" int fsk_b200_cuda_tune(void *p, int lanes, int wpb, int ring)
{
    CudaEngine *ce = (CudaEngine *)p;
    if (lanes && (lanes < 2 || lanes > 32 || (lanes & (lanes - 1)))) {
	fsk_b200_set_error("lanes per stream must be a power of two in 2..32");
	return -EINVAL;
    }
    if (ring && (ring < 64 || (ring & (ring - 1)))) {
	fsk_b200_set_error("ring size must be a power of two >= 64");
	return -EINVAL;
    }
    if (wpb < 0 || wpb > 8) {
	fsk_b200_set_error("warps per block must be 1..8");
	return -EINVAL;
    }
    ce->lanes = lanes;
    ce->wpb = wpb;
    ce->ring = ring;
    return 0;
}

/* exp(-2 pi i k n / fftsize) for k = b_mark, b_space; the argument is reduced
 * exactly in integers and evaluated in double before rounding to float */
extern "C" int fsk_b200_cuda_set_table(void *p, int fftsize, unsigned b_mark, unsigned b_space,
	unsigned bit_nsamples)
{
    CudaEngine *ce = (CudaEngine *)p;
    if (ce->d_tw && ce->tw_fftsize == fftsize && ce->tw_bm == b_mark && ce->tw_bs == b_space
	    && ce->tw_n >= bit_nsamples)
	return 0;
    if (fftsize <= 0 || bit_nsamples == 0) {
	fsk_b200_set_error("set_table: bad size");
	return -EINVAL;
    }
    float4 *h = (float4 *)malloc(sizeof(float4) * bit_nsamples);
    if (!h)
	return -ENOMEM;
    const unsigned long long F = (unsigned long long)fftsize;
    for (unsigned n = 0; n < bit_nsamples; n++) {
	const double am = 2.0 * M_PI * (double)(((unsigned long long)b_mark * n) % F) / (double)F;
	const double as = 2.0 * M_PI * (double)(((unsigned long long)b_space * n) % F) / (double)F;
	h[n].x = (float)cos(am);
	h[n].y = (float)-sin(am);
	h[n].z = (float)cos(as);
	h[n].w = (float)-sin(as);
    }
    if (ce->tw_cap < bit_nsamples) {
	cudaFree(ce->d_tw);
	ce->d_tw = NULL;
	ce->tw_cap = 0;
	if (cudaMalloc(&ce->d_tw, sizeof(float4) * bit_nsamples) != cudaSuccess) {
	    fsk_b200_set_error("set_table: %s", cudaGetErrorString(cudaGetLastError()));
	    free(h);
	    return -ENOMEM;
	}
	ce->tw_cap = bit_nsamples;
    }
    /* synchronous copy on the legacy stream: ordered before any later launch */
    cudaError_t err = cudaMemcpy(ce->d_tw, h, sizeof(float4) * bit_nsamples, cudaMemcpyHostToDevice);
    free(h);
    if (err != cudaSuccess) {
	fsk_b200_set_error("set_table: %s", cudaGetErrorString(err));
	return -EIO;
    }
    ce->tw_n = bit_nsamples;
    ce->tw_fftsize = fftsize;
    ce->tw_bm = b_mark;
    ce->tw_bs = b_space;
    return 0;
}

/* launch shape shared by K1 and K2 */
struct Shape {
    int G, wpb, blocks;
    unsigned ring, tw_in_smem;
    size_t smem;
    fsk_b200_geom geo;
};

static unsigned pow2_ceil(unsigned v)
{
    unsigned p = 1;
    while (p < v)
	p <<= 1;
    return p;
}

static int pick_shape(const CudaEngine *ce, const fsk_b200_geom *g, unsigned need_floats,
	size_t nstreams, Shape *sh)
{
    sh->geo = *g;
    const size_t smem_max = (size_t)ce->smem_optin;
    /* twiddle table in shared memory when it is small next to the rings */
    const size_t tw_bytes = (size_t)g->bit_nsamples * sizeof(float4);
    sh->tw_in_smem = tw_bytes <= 24 * 1024;
    const size_t fixed = sh->tw_in_smem ? tw_bytes : 0;

    unsigned ring = ce->ring ? (unsigned)ce->ring : pow2_ceil(need_floats + 8u);
    if (ring < 64)
	ring = 64;
    int G = ce->lanes;
    if (!G) {
	/* aim for >= 512 resident threads per SM given how many rings fit */
	const size_t per_stream = (size_t)ring * 4 + (size_t)g->n_bits * 8;
	size_t streams_per_sm = (smem_max > fixed ? smem_max - fixed : 0) / (per_stream ? per_stream : 1);
	if (streams_per_sm < 1)
	    streams_per_sm = 1;
	G = 4;
	while (G < 32 && streams_per_sm * (size_t)G < 512)
	    G <<= 1;
    }
    int wpb = ce->wpb ? ce->wpb : 2;
    const size_t scr_bytes = (size_t)g->n_bits * sizeof(float2);
    for (;;) {
	const size_t spw = 32 / G;
	const size_t smem = (fixed + (size_t)wpb * spw * ((size_t)ring * 4 + scr_bytes) + 15) & ~(size_t)15;
	if (smem <= smem_max) {
	    sh->smem = smem;
	    break;
	}
	if (wpb > 1) { wpb--; continue; }
	if (G < 32) { G <<= 1; continue; }
	if (ring) { ring = 0; continue; }	/* not even one ring fits: read global memory directly */
	sh->tw_in_smem = 0;
	sh->smem = (scr_bytes + 15) & ~(size_t)15;
	break;
    }
    sh->G = G;
    sh->wpb = wpb;
    sh->ring = ring;
    unsigned L = 1;
    while (L * 2 * g->n_bits <= (unsigned)G)
	L *= 2;
    sh->geo.lanes_per_window = L;
    /* one block per wpb*(32/G) streams: the hardware block scheduler hands out
     * streams as SM resources free up (streams differ in length and work) */
    const size_t streams_per_block = (size_t)wpb * (32 / G);
    size_t blocks = (nstreams + streams_per_block - 1) / streams_per_block;
    if (blocks > 0x7fffffff)
	blocks = 0x7fffffff;
    sh->blocks = (int)(blocks ? blocks : 1);
    return 0;
}

template <int G>
static cudaError_t launch_find(const Shape &sh, const CudaEngine *ce, const float *samples,
	size_t nstreams, size_t stride, const uint32_t *offset, const uint32_t *nvalid,
	const uint32_t *try_first, const uint32_t *try_max, const uint32_t *try_step,
	const float *limit, const uint8_t *expect_sel, fsk_b200_frame *frames, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_find_frame<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    (int)sh.smem);
    if (e != cudaSuccess)
	return e;
    k_find_frame<G><<<sh.blocks, sh.wpb * 32, sh.smem, st>>>(sh.geo, ce->d_tw, sh.tw_in_smem, sh.ring,
	    samples, (unsigned)nstreams, stride, offset, nvalid, try_first, try_max, try_step, limit,
	    expect_sel, frames);
    g_launches++;
    return cudaGetLastError();
}

extern "C" int fsk_b200_cuda_find_frame_batch(void *p, const fsk_b200_geom *g, const float *samples,
	size_t nstreams, size_t stride, const uint32_t *offset, const uint32_t *nvalid,
	const uint32_t *try_first, const uint32_t *try_max, const uint32_t *try_step,
	const float *limit, const uint8_t *expect_sel, fsk_b200_frame *frames, void *stream)
{
    CudaEngine *ce = (CudaEngine *)p;
    if (!ce->d_tw || ce->tw_n < g->bit_nsamples) {
	fsk_b200_set_error("find_frame_batch: twiddle table not set");
	return -EINVAL;
    }
    Shape sh;
    /* the ring is sized for the widest search of the rx loop: 1.5 bits + span */
    const unsigned need = g->span + 2u * g->bit_nsamples + 8u;
    pick_shape(ce, g, need, nstreams, &sh);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#define FF(GG) e = launch_find<GG>(sh, ce, samples, nstreams, stride, offset, nvalid, try_first, \
	try_max, try_step, limit, expect_sel, frames, st)
    switch (sh.G) {
	case 2: FF(2); break;
	case 4: FF(4); break;
	case 8: FF(8); break;
	case 16: FF(16); break;
	default: FF(32); break;
    }
#undef FF
    if (e != cudaSuccess) {
	fsk_b200_set_error("find_frame_batch launch: %s", cudaGetErrorString(e));
	return -EIO;
    }
    return 0;
}

template <int G>
static cudaError_t launch_rx(const Shape &sh, const CudaEngine *ce, const fsk_b200_loopc *lc,
	const float *samples, size_t nstreams, size_t stride, const uint32_t *nsamples,
	uint32_t nsamples_all, fsk_b200_frame *frames, uint32_t max_frames,
	fsk_b200_stream_state *states, cudaStream_t st)
{
    cudaError_t e;
    if (sh.ring) {
	e = cudaFuncSetAttribute(k_rx<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
	if (e != cudaSuccess)
	    return e;
	k_rx<G, true><<<sh.blocks, sh.wpb * 32, sh.smem, st>>>(sh.geo, *lc, ce->d_tw, sh.tw_in_smem,
		sh.ring, samples, (unsigned)nstreams, stride, nsamples, nsamples_all, frames,
		max_frames, states);
    } else {
	e = cudaFuncSetAttribute(k_rx<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
	if (e != cudaSuccess)
	    return e;
	k_rx<G, false><<<sh.blocks, sh.wpb * 32, sh.smem, st>>>(sh.geo, *lc, ce->d_tw, sh.tw_in_smem,
		0u, samples, (unsigned)nstreams, stride, nsamples, nsamples_all, frames,
		max_frames, states);
    }
    g_launches++;
    return cudaGetLastError();
}

extern "C" int fsk_b200_cuda_rx_batch(void *p, const fsk_b200_geom *g, const fsk_b200_loopc *lc,
	const float *samples, size_t nstreams, size_t stride, const uint32_t *nsamples,
	uint32_t nsamples_all, fsk_b200_frame *frames, uint32_t max_frames,
	fsk_b200_stream_state *states, void *stream)
{
    CudaEngine *ce = (CudaEngine *)p;
    if (!ce->d_tw || ce->tw_n < g->bit_nsamples) {
	fsk_b200_set_error("rx_batch: twiddle table not set");
	return -EINVAL;
    }
    Shape sh;
    const unsigned tmax = lc->try_max_nocarrier > lc->try_max_carrier
	? lc->try_max_nocarrier : lc->try_max_carrier;
    pick_shape(ce, g, tmax + g->span + 8u, nstreams, &sh);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#define RX(GG) e = launch_rx<GG>(sh, ce, lc, samples, nstreams, stride, nsamples, nsamples_all, \
	frames, max_frames, states, st)
    switch (sh.G) {
	case 2: RX(2); break;
	case 4: RX(4); break;
	case 8: RX(8); break;
	case 16: RX(16); break;
	default: RX(32); break;
    }
#undef RX
    if (e != cudaSuccess) {
	fsk_b200_set_error("rx_batch launch: %s", cudaGetErrorString(e));
	return -EIO;
    }
    return 0;
}

/* host buffers in, host results out: slabs of streams, copy/compute overlap */
extern "C" int fsk_b200_cuda_rx_batch_host(void *p, const fsk_b200_geom *g, const fsk_b200_loopc *lc,
	const float *host_samples, size_t nstreams, size_t stride, uint32_t nsamples_all,
	fsk_b200_frame *host_frames, uint32_t max_frames, fsk_b200_stream_state *host_states)
{
    CudaEngine *ce = (CudaEngine *)p;
    /* slab = as many streams as fit ~256 MiB of samples */
    size_t slab = ((size_t)256 << 20) / (stride * sizeof(float));
    if (slab < 1) slab = 1;
    if (slab > nstreams) slab = nstreams;
    if (ce->slab_streams < slab || ce->slab_stride != stride || ce->slab_max_frames < max_frames) {
	for (int i = 0; i < 2; i++) {
	    cudaFree(ce->d_slab[i]); ce->d_slab[i] = NULL;
	    cudaFree(ce->d_slab_frames[i]); ce->d_slab_frames[i] = NULL;
	    cudaFree(ce->d_slab_states[i]); ce->d_slab_states[i] = NULL;
	    CUDA_TRY(cudaMalloc(&ce->d_slab[i], slab * stride * sizeof(float)));
	    CUDA_TRY(cudaMalloc(&ce->d_slab_frames[i], slab * (size_t)max_frames * sizeof(fsk_b200_frame)));
	    CUDA_TRY(cudaMalloc(&ce->d_slab_states[i], slab * sizeof(fsk_b200_stream_state)));
	    if (!ce->st[i])
		CUDA_TRY(cudaStreamCreateWithFlags(&ce->st[i], cudaStreamNonBlocking));
	}
	ce->slab_streams = slab;
	ce->slab_stride = stride;
	ce->slab_max_frames = max_frames;
    }
    int k = 0;
    for (size_t s0 = 0; s0 < nstreams; s0 += slab, k ^= 1) {
	const size_t ns = nstreams - s0 < slab ? nstreams - s0 : slab;
	cudaStream_t st = ce->st[k];
	CUDA_TRY(cudaMemcpyAsync(ce->d_slab[k], host_samples + s0 * stride, ns * stride * sizeof(float),
		    cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(ce->d_slab_states[k], host_states + s0, ns * sizeof(fsk_b200_stream_state),
		    cudaMemcpyHostToDevice, st));
	int rc = fsk_b200_cuda_rx_batch(ce, g, lc, ce->d_slab[k], ns, stride, NULL, nsamples_all,
		ce->d_slab_frames[k], max_frames, ce->d_slab_states[k], st);
	if (rc)
	    return rc;
	CUDA_TRY(cudaMemcpyAsync(host_frames + s0 * (size_t)max_frames, ce->d_slab_frames[k],
		    ns * (size_t)max_frames * sizeof(fsk_b200_frame), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(host_states + s0, ce->d_slab_states[k], ns * sizeof(fsk_b200_stream_state),
		    cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(ce->st[0]));
    CUDA_TRY(cudaStreamSynchronize(ce->st[1]));
    return 0;
}

/* the drop-in fsk_find_frame: one stream, host samples */
extern "C" int fsk_b200_cuda_find_frame_one(void *p, const fsk_b200_geom *g, const float *host_samples,
	unsigned nfloats, unsigned try_first, unsigned try_max, unsigned try_step, float limit,
	fsk_b200_frame *out)
{
    CudaEngine *ce = (CudaEngine *)p;
    const size_t cap = ((size_t)nfloats + 7) & ~(size_t)3;
    if (ce->d_one_cap < cap) {
	cudaFree(ce->d_one);
	ce->d_one = NULL;
	CUDA_TRY(cudaMalloc(&ce->d_one, cap * sizeof(float)));
	ce->d_one_cap = cap;
    }
    if (!ce->d_args) {
	CUDA_TRY(cudaMalloc(&ce->d_args, 8 * sizeof(uint32_t)));
	CUDA_TRY(cudaMalloc(&ce->d_frame, sizeof(fsk_b200_frame)));
    }
    uint32_t args[8] = { 0, nfloats, try_first, try_max, try_step, 0, 0, 0 };
    memcpy(&args[5], &limit, sizeof(float));
    CUDA_TRY(cudaMemcpy(ce->d_one, host_samples, (size_t)nfloats * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ce->d_args, args, sizeof(args), cudaMemcpyHostToDevice));
    int rc = fsk_b200_cuda_find_frame_batch(ce, g, ce->d_one, 1, cap, ce->d_args + 0, ce->d_args + 1,
	    ce->d_args + 2, ce->d_args + 3, ce->d_args + 4, (const float *)(ce->d_args + 5), NULL,
	    ce->d_frame, NULL);
    if (rc)
	return rc;
    CUDA_TRY(cudaMemcpy(out, ce->d_frame, sizeof(*out), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fsk_b200_cuda_band_mags(void *p, int fftsize, const float *host_samples,
	unsigned nsamples, unsigned nbands, float *host_mags)
{
    CudaEngine *ce = (CudaEngine *)p;
    if (nsamples == 0) {
	for (unsigned i = 0; i < nbands; i++)
	    host_mags[i] = 0.f;
	return 0;
    }
    if (ce->d_one_cap < nsamples) {
	cudaFree(ce->d_one);
	ce->d_one = NULL;
	CUDA_TRY(cudaMalloc(&ce->d_one, (size_t)nsamples * sizeof(float)));
	ce->d_one_cap = nsamples;
    }
    if (ce->d_mags_cap < nbands) {
	cudaFree(ce->d_mags);
	ce->d_mags = NULL;
	CUDA_TRY(cudaMalloc(&ce->d_mags, (size_t)nbands * sizeof(float)));
	ce->d_mags_cap = nbands;
    }
    CUDA_TRY(cudaMemcpy(ce->d_one, host_samples, (size_t)nsamples * sizeof(float), cudaMemcpyHostToDevice));
    k_band_mags<<<(nbands + 127) / 128, 128>>>(ce->d_one, nsamples, fftsize, nbands, ce->d_mags);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(host_mags, ce->d_mags, (size_t)nbands * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fsk_b200_cuda_tx_batch(const fsk_b200_tx_config *cfg, const float *sin_table,
	uint32_t table_len, const uint32_t *words, uint32_t nwords, const uint32_t *lead_in,
	float *samples_out, size_t nstreams, size_t stride, uint32_t nsamples_out, void *stream)
{
    TxLens L;
    /* src/minimodem.c:131-132, :96-97, :110-111: size_t * float -> size_t */
    const size_t sample_rate = (size_t)cfg->sample_rate;
    const size_t bit = sample_rate / cfg->data_rate + 0.5f;
    L.bit = (unsigned)bit;
    L.start = (unsigned)(size_t)(bit * cfg->nstartbits);
    L.stop = (unsigned)(size_t)(bit * cfg->nstopbits);
    L.rate = (unsigned)sample_rate;
    const unsigned threads = 128;
    const size_t blocks = (nstreams * 32 + threads - 1) / threads;
    k_tx<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*cfg, L, sin_table, table_len, words,
	    nwords, lead_in, samples_out, (unsigned)nstreams, stride, nsamples_out);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
	fsk_b200_set_error("tx_batch launch: %s", cudaGetErrorString(e));
	return -EIO;
    }
    return 0;
}
