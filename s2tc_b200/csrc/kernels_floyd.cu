// kernels_floyd.cu -- DITHER_FLOYDSTEINBERG pre-pass (reference rgb565_image<..., DITHER_FLOYDSTEINBERG>,
// s2tc_algorithm.cpp:1350-1412 with floyd()/floyd1() :1218-1261), SURVEY.md "next" row N1.
//
// Error diffusion is a 2-D recurrence: texel (x, y) needs the error parts of (x-1, y) and of (x-1..x+1, y-1).
// Rows can therefore run concurrently if each stays two texels behind the row above.  A warp owns a band of 32
// consecutive rows: lane l walks row 32*band + l and at step s handles x = s - 2*l, so inside a warp the
// dependency is satisfied by lock-step execution and the "from above" error travels one lane down with a single
// shuffle per channel.  Between bands it travels through global memory: the last row of a band publishes the
// error it sends below plus a progress counter (release), lane 0 of the next band acquires it.  Bands are
// launched in order (CTA index = band order), so a waiting band always waits on a resident one.
//
// The three colour channels are independent recurrences and share a pass.  Alpha (DXT1: floyd1, DXT3: 4 bits)
// is a second pass because the reference's alpha pass starts from scratch memory the colour pass left behind
// (ref :1380,1397 do not clear the first "this" row): alpha row 0 receives, as incoming error, the RED channel's
// error row of the last image row -- the errors that entered it (odd height) or the ones it sent below (even
// height).  The colour pass exports that row and the alpha pass imports it.
#include "kernels.cuh"

namespace s2tc {

constexpr int kFloydWarps = 4; // bands per CTA

struct FloydArgs {
	const uint8_t *src;
	uint32_t *out;     // reduced texels, 4 B each
	int width, height, srccomps, alphabits;
	int *boundary;     // [bands][width][NCH] error sent below each band's last row
	int *progress;     // [bands] number of boundary entries published
	int *alpha_seed;   // [width] red-channel leftovers for the alpha pass (written by the colour pass)
};

__device__ __forceinline__ int ld_acquire(const int *p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ALPHA == false: r, g, b of every texel (and a copied / constant alpha).  ALPHA == true: the alpha byte only.
template <bool ALPHA, int ASHIFT>
__global__ void __launch_bounds__(kFloydWarps * 32) floyd_kernel(FloydArgs a)
{
	constexpr int NCH = ALPHA ? 1 : 3;
	const int lane = threadIdx.x & 31;
	const int band = blockIdx.x * kFloydWarps + (threadIdx.x >> 5);
	const int nbands = (a.height + 31) >> 5;
	if (band >= nbands)
		return;
	const int row = band * 32 + lane;
	const bool live = row < a.height;
	const int w = a.width;
	const int last_lane = min(31, a.height - 1 - band * 32); // lane of the band's last row
	const bool exports = lane == last_lane;
	const bool image_last = row == a.height - 1;
	const uint8_t *srow = a.src + (size_t) row * w * a.srccomps;
	uint32_t *orow = a.out + (size_t) row * w;
	const int *bin = a.boundary + (size_t) (band - 1) * w * NCH;   // published by the band above
	int *bout = a.boundary + (size_t) band * w * NCH;
	const uint32_t const_alpha = ((1u << a.alphabits) - 1u) << 24;

	int e7[NCH], p5[NCH], a1[NCH], b1[NCH], dout[NCH];
#pragma unroll
	for (int c = 0; c < NCH; ++c)
		e7[c] = p5[c] = a1[c] = b1[c] = dout[c] = 0;
	int known = 0; // boundary entries of the band above known to be published (lane 0 only)
	int above[8][NCH]; // lane 0: errors from the band above for the current group of 8 texels
#pragma unroll
	for (int k = 0; k < 8; ++k)
#pragma unroll
		for (int c = 0; c < NCH; ++c)
			above[k][c] = 0;

	// source texels are fetched kAhead steps before they are needed (a ring of registers, the step loop is unrolled by
	// kAhead so that ring slots are compile-time): with one warp per band nothing else hides the load latency
	constexpr int kAhead = 8;
	auto fetch = [&](int xx) -> uint32_t {
		if (!live || xx < 0 || xx >= w)
			return 0u;
		if (a.srccomps == 4)
			return __ldg(reinterpret_cast<const uint32_t *>(srow) + xx);
		const uint8_t *q = srow + (size_t) xx * 3;
		return (uint32_t) __ldg(q) | ((uint32_t) __ldg(q + 1) << 8) | ((uint32_t) __ldg(q + 2) << 16);
	};
	uint32_t ring[kAhead];
#pragma unroll
	for (int u = 0; u < kAhead; ++u)
		ring[u] = fetch(u - 2 * lane);

	const int steps = w + 1 + 2 * 31;
	for (int s0 = 0; s0 < steps; s0 += kAhead) {
#pragma unroll
	for (int u = 0; u < kAhead; ++u) {
		const int s = s0 + u;
		const int x = s - 2 * lane;
		const uint32_t srcw = ring[u];
		ring[u] = fetch(x + kAhead);
		// error from the row above for texel x: computed by the lane above in the previous step
		int din[NCH];
#pragma unroll
		for (int c = 0; c < NCH; ++c)
			din[c] = __shfl_up_sync(0xFFFFFFFFu, dout[c], 1);
		if (lane == 0) { // x == s here, so x % kAhead == u: once per unrolled group, fetch the next kAhead boundary entries
			if (u == 0 && x < w) {
				const int want = min(x + kAhead, w);
				if (band > 0) {
					while (known < want)
						known = ld_acquire(a.progress + band - 1);
#pragma unroll
					for (int k = 0; k < kAhead; ++k)
#pragma unroll
						for (int c = 0; c < NCH; ++c)
							above[k][c] = x + k < w ? __ldcg(bin + (size_t) (x + k) * NCH + c) : 0;
				} else if (ALPHA) {
#pragma unroll
					for (int k = 0; k < kAhead; ++k)
						above[k][0] = x + k < w ? a.alpha_seed[x + k] : 0; // the colour pass's leftovers seed alpha row 0
				}
			}
#pragma unroll
			for (int c = 0; c < NCH; ++c)
				din[c] = above[u][c];
		}
		if (live && x >= 0 && x <= w) {
			if (x < w) {
				FloydOut o[NCH];
				int incoming[NCH];
#pragma unroll
				for (int c = 0; c < NCH; ++c)
					incoming[c] = din[c] + e7[c];
				if (ALPHA) {
					o[0] = floyd_texel<ASHIFT>((int) (srcw >> 24), incoming[0]);
					reinterpret_cast<uint8_t *>(orow + x)[3] = (uint8_t) o[0].q;
				} else {
					o[0] = floyd_texel<3>((int) (srcw & 0xFFu), incoming[0]);
					o[1 % NCH] = floyd_texel<2>((int) ((srcw >> 8) & 0xFFu), incoming[1 % NCH]);
					o[2 % NCH] = floyd_texel<3>((int) ((srcw >> 16) & 0xFFu), incoming[2 % NCH]);
					const uint32_t alpha = a.srccomps == 4 ? (srcw & 0xFF000000u) : const_alpha; // 8-bit copy or ones; the alpha pass overwrites otherwise
					orow[x] = (uint32_t) o[0].q | ((uint32_t) o[1 % NCH].q << 8) | ((uint32_t) o[2 % NCH].q << 16) | alpha;
					if (image_last && (a.height & 1))
						a.alpha_seed[x] = incoming[0]; // odd height: what entered the red channel of the last row
				}
#pragma unroll
				for (int c = 0; c < NCH; ++c) {
					dout[c] = b1[c] + p5[c] + o[c].e3; // complete error for texel x-1 of the row below
					b1[c] = a1[c];
					a1[c] = o[c].e1;
					p5[c] = o[c].e5;
					e7[c] = o[c].e7;
				}
			} else { // x == w: flush the pipeline, texel w-1 of the row below gets e1(w-2) + e5(w-1)
#pragma unroll
				for (int c = 0; c < NCH; ++c)
					dout[c] = b1[c] + p5[c];
			}
			if (x >= 1) {
				if (exports) {
#pragma unroll
					for (int c = 0; c < NCH; ++c)
						bout[(size_t) (x - 1) * NCH + c] = dout[c];
					if ((x & 7) == 0 || x == w)
						st_release(a.progress + band, x); // entries 0 .. x-1 are visible
				}
				if (!ALPHA && image_last && !(a.height & 1))
					a.alpha_seed[x - 1] = dout[0]; // even height: what the red channel sent below the last row
			}
		}
	}
	}
}

size_t floyd_workspace_bytes(int width, int height)
{
	const size_t bands = (size_t) (height + 31) / 32;
	return bands * (size_t) width * 3 * sizeof(int) + bands * sizeof(int) + (size_t) width * sizeof(int) + 256;
}

cudaError_t launch_prepass_floyd(const void *d_src, int srccomps, int alphabits, int width, int height, void *d_reduced,
		void *d_workspace, cudaStream_t stream)
{
	if (width <= 0 || height <= 0)
		return cudaSuccess;
	const int bands = (height + 31) / 32;
	FloydArgs a;
	a.src = (const uint8_t *) d_src;
	a.out = (uint32_t *) d_reduced;
	a.width = width;
	a.height = height;
	a.srccomps = srccomps;
	a.alphabits = alphabits;
	a.boundary = (int *) d_workspace;
	a.progress = a.boundary + (size_t) bands * width * 3;
	a.alpha_seed = a.progress + bands;
	const dim3 block(kFloydWarps * 32), grid((bands + kFloydWarps - 1) / kFloydWarps);
	cudaError_t e = cudaMemsetAsync(a.progress, 0, (size_t) bands * sizeof(int), stream);
	if (e != cudaSuccess)
		return e;
	floyd_kernel<false, 3><<<grid, block, 0, stream>>>(a);
	if (srccomps == 4 && alphabits != 8) { // ref :1374-1404
		if ((e = cudaMemsetAsync(a.progress, 0, (size_t) bands * sizeof(int), stream)) != cudaSuccess)
			return e;
		if (alphabits == 1)
			floyd_kernel<true, 7><<<grid, block, 0, stream>>>(a);
		else
			floyd_kernel<true, 4><<<grid, block, 0, stream>>>(a);
	}
	return cudaGetLastError();
}

} // namespace s2tc
