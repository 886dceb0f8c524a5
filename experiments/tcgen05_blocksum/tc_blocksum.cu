/*
 * tc_blocksum.cu -- EXPERIMENT (not linked into libfsk_b200.so): the dense / tensor-core formulation of the
 * tone correlation, measured on tcgen05.
 *
 * What the CUDA-core kernels compute per bit window is sum_n x[s+n] * exp(-2 pi i k n / F) for two tones.  The
 * formulation that maps onto the 5th-generation tensor cores without knowing the window positions in
 * advance is a BLOCK SUFFIX-SUM GEMM: cut every stream into blocks of K = 32 samples and multiply each
 * block by a [128 x 32] basis whose row (c, p) (c = re/im of mark/space, p = 0..31) is the block-local
 * twiddle b_c[d] for d >= p and 0 below.  D[(c,p)][blk] is then the demodulated sum of the block's
 * samples from p on; any window of any candidate at any alignment is
 *      D[(c, s % 32)][blk(s)] + sum of whole blocks D[(c,0)][.] + (D[(c,0)][blk(e)] - D[(c, e % 32)][blk(e)])
 * (each term rotated by the block's phase): the per-sample work moves to the tensor pipe and the
 * sliding-window search becomes a handful of lookups per window.
 *
 * This file measures that building block on a B200: fp32 samples from HBM -> shared memory (canonical
 * K-major no-swizzle core-matrix layout) -> tcgen05.mma kind::tf32 (M=128, N=128, K=8 x 4) -> TMEM ->
 * tcgen05.ld -> a reduction that stands in for the consumer.  NPASS=1: one TF32 product (samples
 * truncated to 10 mantissa bits by the tensor core); NPASS=3: basis and samples split into hi + lo
 * (A_hi X_hi + A_lo X_hi + A_hi X_lo), the split of the samples done by the loading threads.
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_blocksum tc_blocksum.cu
 *   ./tc_blocksum [GiB of samples]
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define M_ROWS 128	/* 4 components x 32 suffix start points */
#define K_BLK 32	/* samples per block */
#define N_TILE 128	/* blocks per tile (MMA N) */
#define TILE_FLOATS (N_TILE * K_BLK)

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(2); } } while (0)

/* ---- descriptors (cute/arch/mma_sm100_desc.hpp documents the bit fields) ------------------------- */
/* shared-memory matrix descriptor, SWIZZLE_NONE, K-major: 8-row x 16-byte core matrices; LBO = bytes between
 * core matrices that are neighbours in K, SBO = bytes between 8-row groups */
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;			/* descriptor version of sm_100 */
    return d;					/* base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE */
}
/* instruction descriptor: D = F32, A = B = TF32, both K-major, dense */
__host__ __device__ constexpr uint32_t instr_desc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
	    :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; spin++) {
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
		: "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	if (spin > (1u << 26))
	    __trap();				/* a protocol bug must not hang the box */
    }
}

/* CONS: the stand-in consumer reads every CONS-th 32-column chunk of the table (1 = all of it; the real search
 * looks up a few entries per window) */
template <int NPASS, int CONS>
__global__ void __launch_bounds__(128)
k_blocksum(const float *__restrict__ x, size_t ntiles, const float *__restrict__ basis_hi,
	const float *__restrict__ basis_lo, float *__restrict__ out_energy, float *__restrict__ dbg, unsigned dbg_tiles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    /* [A_hi 16 KB][A_lo 16 KB][X 16 KB][X_lo 16 KB][mbarrier 8][tmem base 4] */
    float *sA = reinterpret_cast<float *>(smem);
    float *sAlo = sA + M_ROWS * K_BLK;
    float *sX = sAlo + M_ROWS * K_BLK;
    float *sXlo = sX + TILE_FLOATS;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sXlo + TILE_FLOATS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const unsigned tid = threadIdx.x, warp = tid >> 5;

    /* canonical layout: element (row, k) -> byte 1024*(row/8) + 128*(k/4) + 16*(row%8) + 4*(k%4):
     * LBO (next core matrix in K) = 128, SBO (next 8 rows) = 1024 */
    auto canon = [](unsigned row, unsigned chunk) { return 256u * (row >> 3) + 32u * chunk + 4u * (row & 7u); };	/* in floats */
    for (unsigned c = tid; c < M_ROWS * (K_BLK / 4); c += 128) {
	const unsigned row = c >> 3, j = c & 7u;
	*reinterpret_cast<float4 *>(sA + canon(row, j)) = *reinterpret_cast<const float4 *>(basis_hi + row * K_BLK + 4 * j);
	*reinterpret_cast<float4 *>(sAlo + canon(row, j)) = *reinterpret_cast<const float4 *>(basis_lo + row * K_BLK + 4 * j);
    }
    if (tid == 0) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(bar)) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "n"(N_TILE) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = instr_desc(M_ROWS, N_TILE);
    const uint32_t aA = smem_u32(sA), aAlo = smem_u32(sAlo), aX = smem_u32(sX), aXlo = smem_u32(sXlo);

    float energy = 0.f;
    uint32_t phase = 0;
    /* software pipeline: the next tile's 16 KB are in flight (registers) while this tile is multiplied and consumed */
    float4 nxt[8];
    if ((size_t)blockIdx.x < ntiles) {
#pragma unroll
	for (int i = 0; i < 8; i++)
	    nxt[i] = __ldg(reinterpret_cast<const float4 *>(x + (size_t)blockIdx.x * TILE_FLOATS) + tid + 128u * i);
    }
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
	/* 1024 chunks of 16 bytes: consecutive threads read consecutive chunks of HBM */
#pragma unroll
	for (int i = 0; i < 8; i++) {
	    const unsigned c = tid + 128u * i, row = c >> 3, j = c & 7u;
	    const float4 v = nxt[i];
	    if (NPASS == 3) {
		float4 hi, lo;
		hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
		hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
		hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
		hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
		*reinterpret_cast<float4 *>(sXlo + canon(row, j)) = lo;
	    }
	    *reinterpret_cast<float4 *>(sX + canon(row, j)) = v;	/* the tensor core reads the top 19 bits */
	}
	if (tile + gridDim.x < ntiles) {
#pragma unroll
	    for (int i = 0; i < 8; i++)
		nxt[i] = __ldg(reinterpret_cast<const float4 *>(x + (tile + gridDim.x) * TILE_FLOATS) + tid + 128u * i);
	}
	asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");	/* generic-proxy stores -> async-proxy reads */
	__syncthreads();
	if (tid == 0) {
	    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
	    for (int k = 0; k < K_BLK / 8; k++) {			/* one MMA covers K = 8: two core matrices */
		const uint32_t ko = 256u * k;
		mma_tf32(tmem, smem_desc(aA + ko, 128, 1024), smem_desc(aX + ko, 128, 1024), idesc, k > 0);
		if (NPASS == 3) {
		    mma_tf32(tmem, smem_desc(aAlo + ko, 128, 1024), smem_desc(aX + ko, 128, 1024), idesc, 1);
		    mma_tf32(tmem, smem_desc(aA + ko, 128, 1024), smem_desc(aXlo + ko, 128, 1024), idesc, 1);
		}
	    }
	    mma_commit(smem_u32(bar));
	}
	mbar_wait(smem_u32(bar), phase);
	phase ^= 1u;
	asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
	/* thread (warp w, lane l) owns TMEM lane 32w + l = output row; 128 columns in 4 loads of 32 */
#pragma unroll
	for (int cb = 0; cb < N_TILE; cb += 32 * CONS) {
	    uint32_t r[32];
	    const uint32_t taddr = tmem + ((warp * 32u) << 16) + (uint32_t)cb;
	    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
		    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		    : "r"(taddr) : "memory");
	    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
	    for (int i = 0; i < 32; i++) {
		const float v = __uint_as_float(r[i]);
		energy = fmaf(v, v, energy);			/* stands in for the consumer of the table */
		if (dbg && tile < dbg_tiles)
		    dbg[(tile * M_ROWS + tid) * N_TILE + cb + i] = v;
	    }
	}
	asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
	__syncthreads();				/* X and the accumulator are free again */
    }
    out_energy[(size_t)blockIdx.x * 128 + tid] = energy;
    __syncthreads();
    if (warp == 0)
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(N_TILE) : "memory");
}

/* the same table on the CUDA cores, for the comparison: one thread per (block, component) running sum */
__global__ void k_blocksum_simt(const float *__restrict__ x, size_t nblocks, const float *__restrict__ basis,
	float *__restrict__ out_energy)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float energy = 0.f;
    for (size_t b = i; b < nblocks; b += (size_t)gridDim.x * blockDim.x) {
	const float4 *p = reinterpret_cast<const float4 *>(x + b * K_BLK);
	float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
	for (int j = K_BLK / 4 - 1; j >= 0; j--) {		/* suffix sums: from the block's end backwards */
	    const float4 v = __ldg(p + j);
	    const float xs[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
	    for (int q = 3; q >= 0; q--) {
		const int d = 4 * j + q;
		s0 = fmaf(xs[q], basis[0 * K_BLK + d], s0);
		s1 = fmaf(xs[q], basis[1 * K_BLK + d], s1);
		s2 = fmaf(xs[q], basis[2 * K_BLK + d], s2);
		s3 = fmaf(xs[q], basis[3 * K_BLK + d], s3);
		energy = fmaf(s0, s0, fmaf(s1, s1, fmaf(s2, s2, fmaf(s3, s3, energy))));
	    }
	}
    }
    out_energy[i] = energy;
}

static float tf32_trunc(float v)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    u &= 0xffffe000u;
    memcpy(&v, &u, 4);
    return v;
}

int main(int argc, char **argv)
{
    const double gib = argc > 1 ? atof(argv[1]) : 8.0;
    const size_t ntiles = (size_t)(gib * 1073741824.0 / (TILE_FLOATS * 4));
    const size_t nfloats = ntiles * TILE_FLOATS;
    /* Bell202 at 48 kHz: mark 1200 Hz, space 2200 Hz; block-local twiddles */
    const double wm = 2 * M_PI * 1200.0 / 48000.0, ws = 2 * M_PI * 2200.0 / 48000.0;
    static float bloc[4][K_BLK], basis[M_ROWS][K_BLK], bhi[M_ROWS][K_BLK], blo[M_ROWS][K_BLK];
    for (int d = 0; d < K_BLK; d++) {
	bloc[0][d] = (float)cos(wm * d); bloc[1][d] = (float)-sin(wm * d);
	bloc[2][d] = (float)cos(ws * d); bloc[3][d] = (float)-sin(ws * d);
    }
    for (int c = 0; c < 4; c++)
	for (int p = 0; p < 32; p++)
	    for (int d = 0; d < K_BLK; d++) {
		const float v = d >= p ? bloc[c][d] : 0.f;
		basis[c * 32 + p][d] = v;
		bhi[c * 32 + p][d] = tf32_trunc(v);
		blo[c * 32 + p][d] = tf32_trunc(v - tf32_trunc(v));
	    }
    /* samples: an FSK-like tone with a little noise, amplitude ~1 */
    const size_t nhost = (size_t)64 * TILE_FLOATS;
    float *hx = (float *)malloc(nhost * 4);
    uint64_t seed = 88172645463325252ull;
    for (size_t i = 0; i < nhost; i++) {
	seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
	const double noise = ((double)(seed >> 11) / 9007199254740992.0 - 0.5) * 0.1;
	hx[i] = (float)(sin(((i / 40) % 3 ? wm : ws) * (double)i) + noise);
    }
    float *dx, *dbhi, *dblo, *dbloc, *den, *ddbg;
    CK(cudaMalloc(&dx, nfloats * 4));
    for (size_t off = 0; off < nfloats; off += nhost)	/* the pattern repeated: content does not matter for the rate */
	CK(cudaMemcpy(dx + off, hx, (nfloats - off < nhost ? nfloats - off : nhost) * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dbhi, sizeof(bhi))); CK(cudaMalloc(&dblo, sizeof(blo))); CK(cudaMalloc(&dbloc, sizeof(bloc)));
    CK(cudaMemcpy(dbhi, bhi, sizeof(bhi), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dblo, blo, sizeof(blo), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbloc, bloc, sizeof(bloc), cudaMemcpyHostToDevice));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int ctas_per_sm = 3;			/* 64 KB of shared memory and 128 TMEM columns each */
    const int grid = sms * ctas_per_sm;
    CK(cudaMalloc(&den, (size_t)grid * 128 * 4 + ((size_t)1 << 22)));
    const unsigned dbg_tiles = 8;
    CK(cudaMalloc(&ddbg, (size_t)dbg_tiles * M_ROWS * N_TILE * 4));
    const size_t smem = 4 * 16384 + 64;
    CK(cudaFuncSetAttribute(k_blocksum<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_blocksum<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_blocksum<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_blocksum<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    /* reference for the first tiles, in double */
    double *ref = (double *)malloc((size_t)dbg_tiles * M_ROWS * N_TILE * sizeof(double));
    double *scale = (double *)malloc((size_t)dbg_tiles * M_ROWS * N_TILE * sizeof(double));
    for (unsigned t = 0; t < dbg_tiles; t++)
	for (int r = 0; r < M_ROWS; r++)
	    for (int n = 0; n < N_TILE; n++) {
		double s = 0, a = 0;
		for (int d = 0; d < K_BLK; d++) {
		    const double xv = hx[((size_t)t * N_TILE + n) * K_BLK + d];
		    s += (double)basis[r][d] * xv;
		    a += fabs((double)basis[r][d] * xv);
		}
		ref[((size_t)t * M_ROWS + r) * N_TILE + n] = s;
		scale[((size_t)t * M_ROWS + r) * N_TILE + n] = a;
	    }
    float *hd = (float *)malloc((size_t)dbg_tiles * M_ROWS * N_TILE * 4);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("{\"experiment\": \"tcgen05 block suffix-sum GEMM (kind::tf32, M=128 N=128 K=32)\", \"gib\": %.2f, \"tiles\": %zu, \"grid\": %d", gib, ntiles, grid);
    for (int variant = 0; variant < 4; variant++) {
	const int npass = (variant & 1) ? 3 : 1, cons = (variant & 2) ? 4 : 1;
	auto launch = [&](float *dbgp, unsigned ndbg) {
	    if (npass == 1 && cons == 1) k_blocksum<1, 1><<<grid, 128, smem>>>(dx, ntiles, dbhi, dblo, den, dbgp, ndbg);
	    if (npass == 3 && cons == 1) k_blocksum<3, 1><<<grid, 128, smem>>>(dx, ntiles, dbhi, dblo, den, dbgp, ndbg);
	    if (npass == 1 && cons == 4) k_blocksum<1, 4><<<grid, 128, smem>>>(dx, ntiles, dbhi, dblo, den, dbgp, ndbg);
	    if (npass == 3 && cons == 4) k_blocksum<3, 4><<<grid, 128, smem>>>(dx, ntiles, dbhi, dblo, den, dbgp, ndbg);
	};
	double worst = 0, worst_abs = 0;
	if (cons == 1) {
	    CK(cudaMemset(ddbg, 0, (size_t)dbg_tiles * M_ROWS * N_TILE * 4));
	    launch(ddbg, dbg_tiles);
	    CK(cudaGetLastError());
	    CK(cudaDeviceSynchronize());
	    CK(cudaMemcpy(hd, ddbg, (size_t)dbg_tiles * M_ROWS * N_TILE * 4, cudaMemcpyDeviceToHost));
	    for (size_t i = 0; i < (size_t)dbg_tiles * M_ROWS * N_TILE; i++) {
		const double err = fabs((double)hd[i] - ref[i]);
		if (err / (scale[i] + 1e-30) > worst) worst = err / (scale[i] + 1e-30);
		if (err > worst_abs) worst_abs = err;
	    }
	}
	float best = 1e30f;
	for (int rep = 0; rep < 5; rep++) {
	    CK(cudaEventRecord(e0));
	    launch(NULL, 0);
	    CK(cudaEventRecord(e1));
	    CK(cudaEventSynchronize(e1));
	    float ms;
	    CK(cudaEventElapsedTime(&ms, e0, e1));
	    if (rep > 0 && ms < best) best = ms;
	}
	CK(cudaGetLastError());
	printf(", \"tf32x%d_consume_1_of_%d\": {\"ms\": %.3f, \"GBps\": %.1f, \"Msamples_per_s\": %.0f, \"max_err_rel_to_sum_abs\": %.3e, \"max_abs_err\": %.3e}",
		npass, cons, best, nfloats * 4.0 / best / 1e6, nfloats / best / 1e3, worst, worst_abs);
    }
    {
	float best = 1e30f;
	for (int rep = 0; rep < 4; rep++) {
	    CK(cudaEventRecord(e0));
	    k_blocksum_simt<<<sms * 8, 256>>>(dx, nfloats / K_BLK, dbloc, den);
	    CK(cudaEventRecord(e1));
	    CK(cudaEventSynchronize(e1));
	    float ms;
	    CK(cudaEventElapsedTime(&ms, e0, e1));
	    if (rep > 0 && ms < best) best = ms;
	}
	CK(cudaGetLastError());
	printf(", \"cuda_core_fp32\": {\"ms\": %.3f, \"GBps\": %.1f, \"Msamples_per_s\": %.0f}", best, nfloats * 4.0 / best / 1e6, nfloats / best / 1e3);
    }
    printf("}\n");
    return 0;
}
