/*
 * tests/emu/selftest.cpp -- TEST INFRASTRUCTURE: the emulator checked against what CUDA defines,
 * on kernels small enough to verify by hand.  `selftest <case>`; cases that must be caught end in
 * abort() with a "simt_emu:" message (tests/test_emu_parity.py checks both kinds).
 */
#include "cuda_runtime.h"

#include <cstdio>
#include <cstring>

static int failures;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "selftest: %s:%d: %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

/* every group of 8 lanes sums its lane ids by butterfly; groups take different numbers of rounds */
__global__ void k_groups(unsigned *out)
{
    const unsigned lane = threadIdx.x & 31, g = lane & 7, grp = lane >> 3;
    const unsigned mask = 0xffu << (8 * grp);
    unsigned acc = 0;
    for (unsigned round = 0; round <= grp + blockIdx.x; round++) {	/* divergent trip counts between groups */
	unsigned v = threadIdx.x + round;
	for (int o = 4; o; o >>= 1)
	    v += __shfl_xor_sync(mask, v, o);
	acc += v;
	__syncwarp(mask);
    }
    const int any = __any_sync(mask, g == 3);
    const unsigned bal = __ballot_sync(mask, (g & 1) != 0);
    if (g == 0) {
	out[(blockIdx.x * (blockDim.x / 8) + threadIdx.x / 8) * 3 + 0] = acc;
	out[(blockIdx.x * (blockDim.x / 8) + threadIdx.x / 8) * 3 + 1] = (unsigned)any;
	out[(blockIdx.x * (blockDim.x / 8) + threadIdx.x / 8) * 3 + 2] = bal;
    }
}

/* block-wide staging through dynamic shared memory with __syncthreads */
__global__ void k_block(const float *in, float *out)
{
    FSK_DYN_SMEM(sm4);
    float *sm = reinterpret_cast<float *>(sm4);
    sm[threadIdx.x] = in[blockIdx.x * blockDim.x + threadIdx.x];
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[blockDim.x - 1 - threadIdx.x];
}

/* cp.async: what a lane reads before and after its wait */
__global__ void k_async(const float *src, float *seen_before, float *seen_after, int do_wait)
{
    FSK_DYN_SMEM(sm4);
    float *sm = reinterpret_cast<float *>(sm4);
    for (int i = 0; i < 4; i++)
	sm[threadIdx.x * 4 + i] = -1.0f;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + threadIdx.x * 4);
    ldgsts16(dst, src + threadIdx.x * 4);
    cp_async_commit();
    seen_before[threadIdx.x] = sm[threadIdx.x * 4];
    if (do_wait)
	cp_async_wait<0>();
    seen_after[threadIdx.x] = sm[threadIdx.x * 4];
    if (!do_wait)
	cp_async_wait<0>();
}

__global__ void k_deadlock()
{
    if ((threadIdx.x & 7) != 5)			/* lane 5 of every group never arrives */
	__syncwarp(0xffu << (8 * ((threadIdx.x & 31) >> 3)));
}

__global__ void k_wrong_mask() { __syncwarp(0x1u); }	/* lanes 1..31 are not in the mask they pass */

__global__ void k_oob()
{
    FSK_DYN_SMEM(sm4);
    static const float z[4] = { 0, 0, 0, 0 };
    ldgsts16((unsigned)__cvta_generic_to_shared(sm4) + 4096u, z);	/* past the 256 bytes asked for */
    cp_async_commit();
    cp_async_wait<0>();
}

__global__ void k_exit_in_flight()
{
    FSK_DYN_SMEM(sm4);
    static const float z[4] = { 0, 0, 0, 0 };
    ldgsts16((unsigned)__cvta_generic_to_shared(sm4), z);
    cp_async_commit();				/* and never waits */
}

int main(int argc, char **argv)
{
    const char *c = argc > 1 ? argv[1] : "";
    if (!strcmp(c, "groups")) {
	unsigned out[3 * 8 * 3];
	memset(out, 0, sizeof(out));
	unsigned *po = out;
	FSK_LAUNCH(k_groups, 3, 64, 0, 0, po);
	for (unsigned b = 0; b < 3; b++)
	    for (unsigned grp = 0; grp < 8; grp++) {
		unsigned want = 0;
		const unsigned base = (grp & 3) * 8 + (grp >> 2) * 32;	/* threadIdx of the group's lane 0 */
		for (unsigned round = 0; round <= (grp & 3) + b; round++)
		    for (unsigned l = 0; l < 8; l++)
			want += base + l + round;
		CHECK(out[(b * 8 + grp) * 3 + 0] == want);
		CHECK(out[(b * 8 + grp) * 3 + 1] == 1u);
		CHECK(out[(b * 8 + grp) * 3 + 2] == (0xaau << (8 * (grp & 3))));
	    }
    } else if (!strcmp(c, "block")) {
	float in[4 * 96], out[4 * 96];
	for (int i = 0; i < 4 * 96; i++)
	    in[i] = (float)i;
	const float *pi = in;
	float *po = out;
	FSK_LAUNCH(k_block, 4, 96, 96 * sizeof(float), 0, pi, po);
	for (int b = 0; b < 4; b++)
	    for (int t = 0; t < 96; t++)
		CHECK(out[b * 96 + t] == in[b * 96 + 95 - t]);
    } else if (!strcmp(c, "async")) {
	alignas(16) float src[32 * 4];
	float before[32], after[32];
	for (int i = 0; i < 128; i++)
	    src[i] = 100.0f + i;
	const float *ps = src;
	float *pb = before, *pa = after;
	const char *mode = getenv("FSK_EMU_ASYNC");
	const bool eager = mode && !strcmp(mode, "eager");
	for (int wait = 0; wait < 2; wait++) {
	    FSK_LAUNCH(k_async, 1, 32, 32 * 16, 0, ps, pb, pa, wait);
	    for (int t = 0; t < 32; t++) {
		CHECK(before[t] == (eager ? src[t * 4] : -1.0f));	/* late: not landed before the wait */
		CHECK(after[t] == ((eager || wait) ? src[t * 4] : -1.0f));
	    }
	}
    } else if (!strcmp(c, "deadlock")) {
	FSK_LAUNCH(k_deadlock, 1, 64, 0, 0);
    } else if (!strcmp(c, "wrong_mask")) {
	FSK_LAUNCH(k_wrong_mask, 1, 32, 0, 0);
    } else if (!strcmp(c, "oob")) {
	FSK_LAUNCH(k_oob, 1, 1, 256, 0);
    } else if (!strcmp(c, "exit_in_flight")) {
	FSK_LAUNCH(k_exit_in_flight, 1, 1, 256, 0);
    } else {
	fprintf(stderr, "usage: selftest groups|block|async|deadlock|wrong_mask|oob|exit_in_flight\n");
	return 2;
    }
    if (failures)
	return 1;
    printf("selftest %s ok\n", c);
    return 0;
}
