// kernels_misc.cu -- the kernels either side of the search: the 565 pre-pass (DITHER_NONE as a
// stand-alone pass, DITHER_SIMPLE as a three-phase carry scan), the random-candidate generator that
// replays glibc's rand() stream in parallel, and the S3TC -> S2TC transcoder.
#include "kernels.cuh"
#include "transcode_core.cuh"

namespace s2tc {

// =====================================================================================================
// DITHER_NONE as its own pass (reference rgb565_image, s2tc_algorithm.cpp:1269-1306).  The encode
// kernels normally fuse these shifts into their loads; this pass backs the exported rgb565_image().
// =====================================================================================================
__global__ void prepass_none_kernel(const uint8_t *__restrict__ src, int srccomps, int alphabits, size_t npixels,
		uint32_t *__restrict__ out)
{
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < npixels; i += stride) {
		if (srccomps == 4) {
			out[i] = reduce_word(__ldg(reinterpret_cast<const uint32_t *>(src) + i), alphabits);
		} else {
			const uint8_t *p = src + i * 3;
			out[i] = reduce_none(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0, alphabits, false);
		}
	}
}

cudaError_t launch_prepass_none(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		cudaStream_t stream)
{
	if (!npixels)
		return cudaSuccess;
	const int threads = 256;
	size_t blocks = (npixels + threads - 1) / threads;
	if (blocks > 148 * 32)
		blocks = 148 * 32;
	prepass_none_kernel<<<(unsigned) blocks, threads, 0, stream>>>((const uint8_t *) d_src, srccomps, alphabits,
			npixels, (uint32_t *) d_reduced);
	return cudaGetLastError();
}

// =====================================================================================================
// DITHER_SIMPLE (reference s2tc_algorithm.cpp:1307-1349): one carry per channel runs through the whole
// image in raster order.  Texels are cut into chunks of kChunk (one thread each), 128 chunks per CTA tile.
//   phase 1  dither_maps_kernel : per chunk, the transfer map carry-in -> carry-out of every channel
//                                 (all start states advanced together); a Hillis-Steele scan inside the
//                                 CTA turns them into "tile start -> chunk start" prefix maps + a tile map
//   phase 2  dither_scan_kernel : one CTA walks the tile maps from the true carry-in and leaves every
//                                 tile's starting carry (or, for sharding, the composed map of the range)
//   phase 3  dither_apply_kernel: every chunk replays the recurrence from its now-known carry
// =====================================================================================================
constexpr int kChunk = 128;       // texels per thread
constexpr int kTileThreads = 128; // chunks per CTA
constexpr int kTilePixels = kChunk * kTileThreads;
constexpr int kTilePitch = kChunk + 1; // words; +1 keeps the per-thread row walks bank-conflict free

struct ChanKinds { int k[4]; };

static ChanKinds chan_kinds(int srccomps, int alphabits)
{
	ChanKinds c;
	c.k[0] = kChanShift3;
	c.k[1] = kChanShift2;
	c.k[2] = kChanShift3;
	c.k[3] = alpha_chan_kind(srccomps, alphabits);
	return c;
}

// stage one tile of source texels in shared memory as 4-byte words (3-byte sources are widened)
__device__ __forceinline__ void load_tile(const uint8_t *__restrict__ src, int srccomps, size_t npixels, size_t tile0,
		uint32_t *tile)
{
	for (int k = threadIdx.x; k < kTilePixels; k += kTileThreads) {
		const size_t p = tile0 + k;
		uint32_t w = 0;
		if (p < npixels) {
			if (srccomps == 4)
				w = __ldg(reinterpret_cast<const uint32_t *>(src) + p);
			else {
				const uint8_t *q = src + p * 3;
				w = (uint32_t) __ldg(q) | ((uint32_t) __ldg(q + 1) << 8) | ((uint32_t) __ldg(q + 2) << 16);
			}
		}
		tile[(k / kChunk) * kTilePitch + (k % kChunk)] = w;
	}
}

__global__ void __launch_bounds__(kTileThreads)
dither_maps_kernel(const uint8_t *__restrict__ src, int srccomps, ChanKinds kinds, size_t npixels,
		CarryMap *__restrict__ prefix /* [chunks][4] */, CarryMap *__restrict__ tilemaps /* [tiles][4] */)
{
	extern __shared__ __align__(16) uint32_t tile[];
	const size_t tile0 = (size_t) blockIdx.x * kTilePixels;
	load_tile(src, srccomps, npixels, tile0, tile);
	__syncthreads();

	const int t = threadIdx.x;
	const size_t first = tile0 + (size_t) t * kChunk;
	const int count = first >= npixels ? 0 : (int) min((size_t) kChunk, npixels - first);
	CarryMap mine[4];
	const uint8_t *row = reinterpret_cast<const uint8_t *>(tile + t * kTilePitch);
#pragma unroll
	for (int ch = 0; ch < 4; ++ch)
		map_of_run(mine[ch], kinds.k[ch], row + ch, 4, count);
	__syncthreads();

	// inclusive scan of the maps across the tile, reusing the texel staging area: buf[2][128][4]
	CarryMap *buf = reinterpret_cast<CarryMap *>(tile);
	int cur = 0;
#pragma unroll
	for (int ch = 0; ch < 4; ++ch)
		buf[t * 4 + ch] = mine[ch];
	__syncthreads();
	for (int d = 1; d < kTileThreads; d <<= 1) {
		CarryMap *in = buf + cur * kTileThreads * 4, *out = buf + (cur ^ 1) * kTileThreads * 4;
#pragma unroll
		for (int ch = 0; ch < 4; ++ch) {
			if (t >= d) {
				CarryMap r;
				map_compose(r, in[(t - d) * 4 + ch], in[t * 4 + ch], kinds.k[ch]);
				out[t * 4 + ch] = r;
			} else {
				out[t * 4 + ch] = in[t * 4 + ch];
			}
		}
		cur ^= 1;
		__syncthreads();
	}
	const CarryMap *incl = buf + cur * kTileThreads * 4;
	const size_t chunk = (size_t) blockIdx.x * kTileThreads + t;
#pragma unroll
	for (int ch = 0; ch < 4; ++ch) {
		CarryMap ex;
		if (t == 0)
			map_identity(ex, kinds.k[ch]);
		else
			ex = incl[(t - 1) * 4 + ch];
		prefix[chunk * 4 + ch] = ex;
		if (t == kTileThreads - 1)
			tilemaps[(size_t) blockIdx.x * 4 + ch] = incl[t * 4 + ch];
	}
}

// One CTA.  summary == nullptr: carry[] (4 ints) is the carry into tile 0; writes tile_carry[tile][4] and
// leaves the carry out of the last tile in carry[].  summary != nullptr: writes the composed maps of all
// tiles instead (the "transfer function" of this texel range, exchanged between GPUs when a carry
// chain is sharded).
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
dither_scan_kernel(const CarryMap *__restrict__ tilemaps, size_t ntiles, ChanKinds kinds, int *carry,
		int *__restrict__ tile_carry, CarryMap *__restrict__ partial /* [kScanThreads][4] */, CarryMap *summary)
{
	__shared__ int start[kScanThreads][4];
	const int t = threadIdx.x;
	const size_t per = (ntiles + kScanThreads - 1) / kScanThreads;
	const size_t lo = min(ntiles, (size_t) t * per), hi = min(ntiles, lo + per);
	for (int ch = 0; ch < 4; ++ch) {
		CarryMap acc;
		map_identity(acc, kinds.k[ch]);
		for (size_t i = lo; i < hi; ++i)
			map_compose(acc, acc, tilemaps[i * 4 + ch], kinds.k[ch]);
		partial[t * 4 + ch] = acc;
	}
	__syncthreads();
	if (summary) {
		if (t < 4) {
			CarryMap acc;
			map_identity(acc, kinds.k[t]);
			for (int i = 0; i < kScanThreads; ++i)
				map_compose(acc, acc, partial[i * 4 + t], kinds.k[t]);
			summary[t] = acc;
		}
		return;
	}
	if (t < 4) {
		int c = carry[t];
		for (int i = 0; i < kScanThreads; ++i) {
			start[i][t] = c;
			c = map_apply(partial[i * 4 + t], kinds.k[t], c);
		}
		carry[t] = c;
	}
	__syncthreads();
	for (int ch = 0; ch < 4; ++ch) {
		int c = start[t][ch];
		for (size_t i = lo; i < hi; ++i) {
			tile_carry[i * 4 + ch] = c;
			c = map_apply(tilemaps[i * 4 + ch], kinds.k[ch], c);
		}
	}
}

__global__ void __launch_bounds__(kTileThreads)
dither_apply_kernel(const uint8_t *__restrict__ src, int srccomps, int alphabits, ChanKinds kinds, size_t npixels,
		const CarryMap *__restrict__ prefix, const int *__restrict__ tile_carry, uint32_t *__restrict__ out)
{
	extern __shared__ __align__(16) uint32_t tile[];
	const size_t tile0 = (size_t) blockIdx.x * kTilePixels;
	load_tile(src, srccomps, npixels, tile0, tile);
	__syncthreads();

	const int t = threadIdx.x;
	const size_t first = tile0 + (size_t) t * kChunk;
	const int count = first >= npixels ? 0 : (int) min((size_t) kChunk, npixels - first);
	const size_t chunk = (size_t) blockIdx.x * kTileThreads + t;
	uint8_t *row = reinterpret_cast<uint8_t *>(tile + t * kTilePitch);
#pragma unroll
	for (int ch = 0; ch < 4; ++ch) {
		const int kind = kinds.k[ch];
		if (kind == kChanCopy) {
			if (srccomps != 4) { // constant alpha (ref :1342-1347); the 8-bit copy needs nothing
				const uint8_t ones = (uint8_t) ((1u << alphabits) - 1u);
				for (int i = 0; i < count; ++i)
					row[i * 4 + 3] = ones;
			}
			continue;
		}
		const int c0 = map_apply(prefix[chunk * 4 + ch], kind, tile_carry[(size_t) blockIdx.x * 4 + ch]);
		replay_run(kind, c0, row + ch, 4, count, row + ch);
	}
	__syncthreads();
	for (int k = threadIdx.x; k < kTilePixels; k += kTileThreads) {
		const size_t p = tile0 + k;
		if (p < npixels)
			out[p] = tile[(k / kChunk) * kTilePitch + (k % kChunk)];
	}
}

static size_t dither_tiles(size_t npixels) { return (npixels + kTilePixels - 1) / kTilePixels; }

// workspace: prefix maps [tiles*128][4] | tile maps [tiles][4] | scan partials [1024][4] | tile carries [tiles][4]
size_t dither_workspace_bytes(size_t npixels)
{
	const size_t tiles = dither_tiles(npixels);
	return (tiles * kTileThreads * 4 + tiles * 4 + (size_t) kScanThreads * 4) * sizeof(CarryMap) + tiles * 4 * sizeof(int) + 64;
}

static cudaError_t run_dither(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		int *d_carry, CarryMap *d_summary, void *d_workspace, cudaStream_t stream)
{
	if (!npixels)
		return cudaSuccess;
	const size_t tiles = dither_tiles(npixels);
	CarryMap *prefix = (CarryMap *) d_workspace;
	CarryMap *tilemaps = prefix + tiles * kTileThreads * 4;
	CarryMap *partial = tilemaps + tiles * 4;
	int *tile_carry = (int *) (partial + (size_t) kScanThreads * 4);
	const ChanKinds kinds = chan_kinds(srccomps, alphabits);
	const size_t smem = (size_t) kTileThreads * kTilePitch * 4;
	static_assert((size_t) kTileThreads * kTilePitch * 4 >= 2 * kTileThreads * 4 * sizeof(CarryMap), "scan buffers must fit the tile");
	cudaError_t e;
	if ((e = cudaFuncSetAttribute(dither_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess)
		return e;
	if ((e = cudaFuncSetAttribute(dither_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess)
		return e;
	dither_maps_kernel<<<(unsigned) tiles, kTileThreads, smem, stream>>>((const uint8_t *) d_src, srccomps, kinds, npixels,
			prefix, tilemaps);
	dither_scan_kernel<<<1, kScanThreads, 0, stream>>>(tilemaps, tiles, kinds, d_carry, tile_carry, partial, d_summary);
	if (!d_summary)
		dither_apply_kernel<<<(unsigned) tiles, kTileThreads, smem, stream>>>((const uint8_t *) d_src, srccomps, alphabits,
				kinds, npixels, prefix, tile_carry, (uint32_t *) d_reduced);
	return cudaGetLastError();
}

cudaError_t launch_prepass_simple(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		int *d_carry, void *d_workspace, cudaStream_t stream)
{
	return run_dither(d_src, srccomps, alphabits, npixels, d_reduced, d_carry, nullptr, d_workspace, stream);
}

cudaError_t launch_dither_summary(const void *d_src, int srccomps, int alphabits, size_t npixels, CarryMap *d_summary,
		void *d_workspace, cudaStream_t stream)
{
	return run_dither(d_src, srccomps, alphabits, npixels, nullptr, nullptr, d_summary, d_workspace, stream);
}

// =====================================================================================================
// Random candidates (reference s2tc_algorithm.cpp:962-993).  The reference pulls 3*nrandom (DXT5:
// 4*nrandom) values per block from one global rand() stream, block after block.  Thread t owns
// `blocks_per_thread` consecutive blocks, seeks its private replica to the first draw of its first block
// with O(log t) polynomial products (glibc_rand.cuh) and then generates sequentially.
// =====================================================================================================
template <int DXT>
__global__ void __launch_bounds__(128)
random_candidates_kernel(ImageView v, int nrandom, const RandPlan *__restrict__ plan, int blocks_per_thread,
		uint16_t *__restrict__ cand_c, uint8_t *__restrict__ cand_a)
{
	const int nblocks = v.blocks_w * v.blocks_h;
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
	const long long b0 = (long long) t * blocks_per_thread;
	if (b0 >= nblocks)
		return;
	GlibcRand rng;
	rand_plan_seek(*plan, t, rng);
	const int b1 = (int) min((long long) nblocks, b0 + blocks_per_thread);
	for (int blk = (int) b0; blk < b1; ++blk) {
		const int by = blk / v.blocks_w, bx = blk - by * v.blocks_w;
		Block b;
		load_block(v, bx, by, b);
		uint32_t c[16];
		uint8_t ca[16];
		const int n = gather_colors<DXT>(b, c, ca);
		const CandBox box = candidate_box(c, ca, n);
		const size_t o = (size_t) blk * nrandom;
		for (int k = 0; k < nrandom; ++k) {
			const uint32_t p = draw_candidate<DXT>(box, rng);
			cand_c[o + k] = (uint16_t) to565(p);
			if (DXT == kDxt5)
				cand_a[o + k] = (uint8_t) (p >> 24);
		}
	}
}

cudaError_t launch_random_candidates(int dxt, int nrandom, const ImageView &v, const RandPlan *d_plan,
		int blocks_per_thread, uint16_t *d_cand_c, uint8_t *d_cand_a, cudaStream_t stream)
{
	const long long nblocks = (long long) v.blocks_w * v.blocks_h;
	if (nblocks == 0 || nrandom <= 0)
		return cudaSuccess;
	const long long threads = (nblocks + blocks_per_thread - 1) / blocks_per_thread;
	const dim3 block(128), grid((unsigned) ((threads + 127) / 128));
	switch (dxt) {
	case kDxt1: random_candidates_kernel<kDxt1><<<grid, block, 0, stream>>>(v, nrandom, d_plan, blocks_per_thread, d_cand_c, d_cand_a); break;
	case kDxt3: random_candidates_kernel<kDxt3><<<grid, block, 0, stream>>>(v, nrandom, d_plan, blocks_per_thread, d_cand_c, d_cand_a); break;
	default: random_candidates_kernel<kDxt5><<<grid, block, 0, stream>>>(v, nrandom, d_plan, blocks_per_thread, d_cand_c, d_cand_a); break;
	}
	return cudaGetLastError();
}

// =====================================================================================================
// S3TC -> S2TC transcode (reference s2tc_from_s3tc.cpp:254-263): 8/16 bytes in, same bytes out, in place.
// Pure streaming: one 64/128-bit load and store per block.
// =====================================================================================================
__global__ void transcode_kernel(int dxt, void *blocks, size_t nblocks)
{
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += stride) {
		if (dxt == kDxt1) {
			uint2 b = reinterpret_cast<uint2 *>(blocks)[i];
			transcode_color_dxt1(b.x, b.y);
			reinterpret_cast<uint2 *>(blocks)[i] = b;
		} else {
			uint4 b = reinterpret_cast<uint4 *>(blocks)[i];
			transcode_color_opaque(b.z, b.w);
			if (dxt == kDxt5) {
				const uint64_t a = transcode_alpha_dxt5((uint64_t) b.x | ((uint64_t) b.y << 32));
				b.x = (uint32_t) a;
				b.y = (uint32_t) (a >> 32);
			}
			reinterpret_cast<uint4 *>(blocks)[i] = b;
		}
	}
}

cudaError_t launch_transcode(int dxt, void *d_blocks, size_t nblocks, cudaStream_t stream)
{
	if (!nblocks)
		return cudaSuccess;
	const int threads = 256;
	size_t blocks = (nblocks + threads - 1) / threads;
	if (blocks > 148 * 16)
		blocks = 148 * 16;
	transcode_kernel<<<(unsigned) blocks, threads, 0, stream>>>(dxt, d_blocks, nblocks);
	return cudaGetLastError();
}

// =====================================================================================================
// Measurement aid: sustained INT32 min+add issue rate, the denominator for the search kernels' roofline
// (search modes are bound by the integer pipes, not by HBM).  8 independent chains per
// thread of exactly the two operations the pair scan is made of (IMNMX + IADD).
// =====================================================================================================
__global__ void __launch_bounds__(256) int32_peak_kernel(int iters, int seed, int *sink)
{
	int a[8], s[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		a[k] = seed + threadIdx.x * (k + 1);
		s[k] = 0;
	}
	int b = seed ^ (blockIdx.x << 8);
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k)
			s[k] += min(a[k], b); // the pair scan's inner operation: accumulate the smaller distance
		b += 3;
	}
	int r = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k)
		r ^= s[k];
	if (r == 0x7FFFFFFF)
		*sink = r;
}

cudaError_t launch_int32_peak(int iters, int ctas, int *d_sink, cudaStream_t stream)
{
	int32_peak_kernel<<<ctas, 256, 0, stream>>>(iters, 12345, d_sink);
	return cudaGetLastError();
}

} // namespace s2tc
