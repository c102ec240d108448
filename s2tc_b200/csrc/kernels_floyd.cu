// kernels_floyd.cu -- DITHER_FLOYDSTEINBERG pre-pass (reference rgb565_image<..., DITHER_FLOYDSTEINBERG>,
// s2tc_algorithm.cpp:1350-1412 with floyd()/floyd1() :1218-1261), SURVEY.md "next" row N1.
//
// Error diffusion is a 2-D recurrence: texel (x, y) needs the error parts of (x-1, y) and of (x-1..x+1, y-1).
// Rows can therefore run concurrently if each stays two texels behind the row above: the critical path of an image is
// W + 2 H texel steps, whatever the hardware.  The kernel is built to make one step as short as the dependent chain of
// floyd() allows (~15 integer operations):
//   * a warp owns one CHANNEL of a band of 32 consecutive rows: lane l walks row 32 band + l and at step s handles
//     x = s - 2 l, so inside a warp the dependency is satisfied by lock-step execution and the "from above" error
//     travels one lane down with a single shuffle.  r, g, b are independent recurrences (ref :1366-1368): three warps
//     per band, a third of the work per step each;
//   * source texels arrive 16 steps early as one 128-bit load per lane every four steps; each warp stores its own
//     byte of the reduced texel;
//   * between bands the error row travels through global memory WITHOUT fences: entry x of a band's boundary is one
//     64-bit word {x + 1, error}, written and read with single 64-bit L2 accesses, so a reader either sees the complete
//     entry or the zero the buffer was cleared to.  The next band asks for eight entries at a time, one group of steps
//     before it needs them, and only asks again if one had not arrived;
//   * the groups of 16 steps in which every lane is inside its row (almost all of them) run without a single range test;
//   * bands are handed to CTAs by a ticket counter, so a band only ever waits on a band whose CTA is already running
//     (no assumption about the order in which the hardware starts CTAs).
// 8192^2: 8-10 ms per pass (DXT5: one pass, DXT1 / DXT3: colour pass + alpha pass); round 1 (one warp per band for all
// channels, release/acquire counters every 8 texels) took 69 ms for DXT1.  A step still costs ~270 clocks against a
// dependent chain of ~120: the per-step byte store and the 64-bit source loads touch 32 lines each and share the
// load/store path with the shuffles (next: byte planes staged through shared memory).
//
// Alpha (DXT1: floyd1, DXT3: 4 bits) is a second pass because the reference's alpha pass starts from scratch memory
// the colour pass left behind (ref :1380,1397 do not clear the first "this" row): alpha row 0 receives, as incoming
// error, the RED channel's error row of the last image row -- the errors that entered it (odd height) or the ones it
// sent below (even height).  The red warp of the colour pass exports that row and the alpha pass imports it.  (This
// is also why a Floyd-Steinberg image cannot be split across GPUs with any gain: the colour pass is one dependency
// chain from the first row to the last, and the alpha pass can only start when the colour pass has finished.)
#include "kernels.cuh"

namespace s2tc {

struct FloydArgs {
	const uint8_t *src;
	uint32_t *out;     // reduced texels, 4 B each
	int width, height, srccomps, alphabits;
	uint2 *boundary;   // [bands][channels][width] {x + 1, error sent below the band's last row to texel x}; zeroed per pass
	int *ticket;       // band dispenser, zeroed per pass
	int *alpha_seed;   // [width] red-channel leftovers for the alpha pass (written by the colour pass of the image's last rows)
	// row shards (launch_floyd_rows): `height` rows of an image whose other rows are handled by other calls / GPUs
	const int *seed;   // [channels][width] errors entering row 0 from the rows above (NULL: none); the alpha pass of an image's
	                   // first rows is seeded with the red leftovers of its last row (ref :1380,1397)
	bool more_below;   // the image continues below these rows: the last row publishes its errors like any band boundary
	bool image_last;   // these rows end the image: their last row leaves the alpha seed
	bool image_odd;    // parity of the IMAGE's height
};

// Boundary entries are single 64-bit words read and written with L2-level (.cg) accesses: a reader sees an entry whole or
// not at all.  One load INSTRUCTION per group of eight steps (eight lanes, one entry each): eight separate loads by one
// lane, each holding the warp until it returned, cost 3.7 us per group (measured) and tripled the time per step.
__device__ __forceinline__ uint2 ld_entry(const uint2 *p)
{
	uint2 v;
	asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_entry(uint2 *p, uint32_t x, uint32_t y)
{
	asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y));
}
// One warp, one channel (byte `ch` of the texel), one band.  SHIFT: 3 / 2 (r, b / g), 4 (DXT3 alpha), 7 (floyd1).
// VEC: RGBA source with even width: texels arrive as aligned pairs (x is even at even steps for every lane).
// The state of a lane and one group of 16 steps; STEADY groups are those in which every lane of a full band is inside
// its row (1 <= x, x + 16 < w): no range tests at all.
template <int SHIFT, bool VEC, bool ALPHA>
struct FloydLane {
	static constexpr int kAhead = 16; // source texels in flight per lane
	const FloydArgs &a;
	int band, lane, ch, w;
	bool live, exports, seeds, odd_height, fill_alpha, imports;
	const uint8_t *srow;
	uint8_t *orow;
	const uint2 *bin;
	uint2 *bout;
	const int *seedrow; // band 0: errors entering its first row, or NULL
	uint32_t const_alpha;
	int shift8;
	int e7 = 0, p5 = 0, a1 = 0, b1 = 0, dout = 0;
	uint32_t ring[kAhead]; // slot u: texel s0 + u - 2 lane of the current group
	uint2 cur = make_uint2(0u, 0u), nxt = make_uint2(0u, 0u); // boundary entries: lane k < 8 holds entry s + k / s + 8 + k

	// CHECK: 0 = none (steady groups), 1 = all, 2 = lower bound only (head groups of a full band)
	template <int CHECK> __device__ __forceinline__ uint32_t fetch1(int xx) const
	{
		if ((CHECK == 1 && (!live || xx < 0 || xx >= w)) || (CHECK == 2 && xx < 0))
			return 0u;
		if (a.srccomps == 4)
			return __ldg(reinterpret_cast<const uint32_t *>(srow) + xx);
		const uint8_t *q = srow + (size_t) xx * 3;
		return (uint32_t) __ldg(q) | ((uint32_t) __ldg(q + 1) << 8) | ((uint32_t) __ldg(q + 2) << 16);
	}
	template <int CHECK> __device__ __forceinline__ uint2 fetch2(int xx) const // xx even; w even: the pair is inside or outside the row as a whole
	{
		if ((CHECK == 1 && (!live || xx < 0 || xx >= w)) || (CHECK == 2 && xx < 0))
			return make_uint2(0u, 0u);
		return __ldg(reinterpret_cast<const uint2 *>(srow + (size_t) xx * 4));
	}

	// MODE 1 (steady): every lane inside its row, 1 <= x and x + 16 < w.  MODE 2 (head): the first groups of a full band,
	// x + 16 < w but x may still be negative -- these are the steps the next band waits for.  MODE 0: everything else.
	template <int MODE> __device__ __forceinline__ void group(int s0)
	{
		constexpr bool STEADY = MODE == 1;
		constexpr bool HEAD = MODE == 2;
#pragma unroll
		for (int u = 0; u < kAhead; ++u) {
			const int s = s0 + u;
			const int x = s - 2 * lane;
			const uint32_t srcw = ring[u];
			if (VEC) {
				if (u & 1) { // slots u-1, u are free: texels x - 1 + kAhead (even) and x + kAhead
					const uint2 t = fetch2<STEADY ? 0 : (HEAD ? 2 : 1)>(x - 1 + kAhead);
					ring[u - 1] = t.x;
					ring[u] = t.y;
				}
			} else {
				ring[u] = fetch1<STEADY ? 0 : (HEAD ? 2 : 1)>(x + kAhead);
			}
			if ((u & 7) == 0) { // entries s .. s+7 become cur, s+8 .. s+15 are requested
				if (imports) {
					const bool mine = lane < 8 && (STEADY || HEAD || s + lane < w);
					while (!__all_sync(0xFFFFFFFFu, !mine || nxt.x == (uint32_t) (s + lane + 1))) { // rare: ask again
						if (mine)
							nxt = ld_entry(bin + s + lane);
					}
					cur = nxt;
					if (lane < 8 && (STEADY || HEAD || s + 8 + lane < w))
						nxt = ld_entry(bin + s + 8 + lane);
				} else if (seedrow) { // band 0 of a seeded pass: the rows above (another shard), or the colour pass's leftovers (alpha)
					cur.y = (lane < 8 && s + lane < w) ? (uint32_t) seedrow[s + lane] : 0u;
				}
			}
			// error from the row above for texel x: computed by the lane above in the previous step
			int din = __shfl_up_sync(0xFFFFFFFFu, dout, 1);
			const int from_above = (int) __shfl_sync(0xFFFFFFFFu, cur.y, u & 7);
			if (lane == 0)
				din = from_above;
			if (STEADY || (HEAD && x >= 0) || (live && x >= 0 && x <= w)) {
				if (STEADY || HEAD || x < w) {
					const int incoming = din + e7;
					const FloydOut o = floyd_texel<SHIFT>((int) ((srcw >> shift8) & 0xFFu), incoming);
					orow[(size_t) x * 4 + ch] = (uint8_t) o.q;
					if (fill_alpha)
						orow[(size_t) x * 4 + 3] = (uint8_t) (a.srccomps == 4 ? (srcw >> 24) : const_alpha);
					if (seeds && odd_height)
						a.alpha_seed[x] = incoming; // odd height: what entered the red channel of the last row
					dout = b1 + p5 + o.e3; // complete error for texel x-1 of the row below
					b1 = a1;
					a1 = o.e1;
					p5 = o.e5;
					e7 = o.e7;
				} else { // x == w: flush the pipeline, texel w-1 of the row below gets e1(w-2) + e5(w-1)
					dout = b1 + p5;
				}
				if (STEADY || x >= 1) {
					if (exports)
						st_entry(bout + (x - 1), (uint32_t) x, (uint32_t) dout);
					if (seeds && !odd_height)
						a.alpha_seed[x - 1] = dout; // even height: what the red channel sent below the last row
				}
			}
		}
	}
};

template <int SHIFT, bool VEC, bool ALPHA>
__device__ __forceinline__ void floyd_band(const FloydArgs &a, int band, int lane, int ch, int nch, int chi)
{
	typedef FloydLane<SHIFT, VEC, ALPHA> Lane;
	constexpr int kAhead = Lane::kAhead;
	const int row = band * 32 + lane;
	const int w = a.width;
	const int last_lane = min(31, a.height - 1 - band * 32); // lane of the band's last row
	const bool image_last = row == a.height - 1 && a.image_last;
	Lane L{a, band, lane, ch, w};
	L.live = row < a.height;
	L.exports = lane == last_lane && (row + 1 < a.height || a.more_below);
	L.seeds = !ALPHA && ch == 0 && image_last; // the red warp leaves the alpha pass its seed row
	L.odd_height = a.image_odd;
	L.seedrow = (band == 0 && a.seed) ? a.seed + (size_t) chi * w : nullptr;
	L.fill_alpha = !ALPHA && ch == 0 && (a.srccomps != 4 || a.alphabits == 8); // no alpha pass: copy or all ones
	L.imports = band > 0; // warp-uniform
	L.srow = a.src + (size_t) row * w * a.srccomps;
	L.orow = reinterpret_cast<uint8_t *>(a.out + (size_t) row * w);
	L.bin = a.boundary + ((size_t) (band - 1) * nch + chi) * w; // published by the band above
	L.bout = a.boundary + ((size_t) band * nch + chi) * w;
	L.const_alpha = (1u << a.alphabits) - 1u;
	L.shift8 = 8 * ch;
#pragma unroll
	for (int u = 0; u < kAhead; u += 2) {
		if (VEC) {
			const uint2 t = L.template fetch2<1>(u - 2 * lane);
			L.ring[u] = t.x;
			L.ring[u + 1] = t.y;
		} else {
			L.ring[u] = L.template fetch1<1>(u - 2 * lane);
			L.ring[u + 1] = L.template fetch1<1>(u + 1 - 2 * lane);
		}
	}
	// errors from the band above: lane k < 8 holds entry s + k of the current group of eight steps (cur) and has already
	// asked for entry s + 8 + k of the next (nxt) -- one load instruction per group for the whole warp; lane 0 gets the
	// entry of its texel by shuffle.  A band that follows the band above too closely finds its requests answered before
	// the entries were written and has to ask again, and bands run at the same speed, so it would stay that close for
	// ever: it waits once, at the start, until the band above is kSlack entries ahead of what the first group needs.
	constexpr int slack = 12; // 6 / 12 / 24 measured alike on 8192^2
	if (L.imports) {
		const int far = min(w - 1, 7 + slack);
		while (ld_entry(L.bin + far).x == 0u)
			;
		if (lane < 8 && lane < w)
			L.nxt = ld_entry(L.bin + lane);
	}
	const int steps = w + 1 + 2 * 31;
	const bool full_band = band * 32 + 31 < a.height; // all 32 rows exist
	for (int s0 = 0; s0 < steps; s0 += kAhead) {
		if (full_band && s0 + 2 * kAhead <= w) {
			if (s0 >= 64)
				L.template group<1>(s0);
			else
				L.template group<2>(s0);
		} else {
			L.template group<0>(s0);
		}
	}
}

// ALPHA == false: three warps per band, warp c = channel c of r, g, b (warp 0 also copies / sets alpha when no alpha pass
// follows).  ALPHA == true: one warp per band, the alpha byte only (ASHIFT 7: floyd1, 4: DXT3).
template <bool ALPHA, int ASHIFT>
__global__ void __launch_bounds__(ALPHA ? 32 : 96) floyd_kernel(FloydArgs a)
{
	__shared__ int s_band;
	if (threadIdx.x == 0)
		s_band = atomicAdd(a.ticket, 1);
	__syncthreads();
	const int band = s_band;
	const int lane = threadIdx.x & 31;
	const bool vec = a.srccomps == 4 && (a.width & 1) == 0 && (((size_t) a.src) & 7) == 0;
	if (ALPHA) {
		if (vec)
			floyd_band<ASHIFT, true, true>(a, band, lane, 3, 1, 0);
		else
			floyd_band<ASHIFT, false, true>(a, band, lane, 3, 1, 0);
	} else {
		const int ch = (int) (threadIdx.x >> 5);
		if (ch == 1) {
			if (vec)
				floyd_band<2, true, false>(a, band, lane, 1, 3, 1);
			else
				floyd_band<2, false, false>(a, band, lane, 1, 3, 1);
		} else {
			if (vec)
				floyd_band<3, true, false>(a, band, lane, ch, 3, ch);
			else
				floyd_band<3, false, false>(a, band, lane, ch, 3, ch);
		}
	}
}

size_t floyd_workspace_bytes(int width, int height)
{
	const size_t bands = (size_t) (height + 31) / 32;
	return bands * (size_t) width * 3 * sizeof(uint2) + (size_t) width * sizeof(int) + 512;
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
	a.ticket = (int *) d_workspace;                       // 256 bytes reserved
	a.alpha_seed = a.ticket + 64;
	a.boundary = (uint2 *) (((uintptr_t) (a.alpha_seed + width) + 255) & ~(uintptr_t) 255);
	a.seed = nullptr;
	a.more_below = false;
	a.image_last = true;
	a.image_odd = height & 1;
	const size_t bbytes = (size_t) bands * width * 3 * sizeof(uint2);
	cudaError_t e = cudaMemsetAsync(a.ticket, 0, 256, stream);
	if (e == cudaSuccess)
		e = cudaMemsetAsync(a.boundary, 0, bbytes, stream);
	if (e != cudaSuccess)
		return e;
	floyd_kernel<false, 3><<<bands, 96, 0, stream>>>(a);
	if (srccomps == 4 && alphabits != 8) { // ref :1374-1404
		if ((e = cudaMemsetAsync(a.ticket, 0, 256, stream)) != cudaSuccess)
			return e;
		if ((e = cudaMemsetAsync(a.boundary, 0, (size_t) bands * width * sizeof(uint2), stream)) != cudaSuccess)
			return e;
		a.seed = a.alpha_seed;
		if (alphabits == 1)
			floyd_kernel<true, 7><<<bands, 32, 0, stream>>>(a);
		else
			floyd_kernel<true, 4><<<bands, 32, 0, stream>>>(a);
	}
	return cudaGetLastError();
}

// ---- row shards ---------------------------------------------------------------------------------------------------
// The error row that crosses a shard boundary is the same thing as the row that crosses a band boundary; a shard's last
// band publishes it in the boundary format and this kernel hands the caller plain ints.
__global__ void boundary_errors_kernel(const uint2 *__restrict__ b, int n, int *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = (int) b[i].y;
}

// One pass over `rows` texel rows starting at row y0 of an image of image_height rows.  phase 0: r, g, b (d_err_in /
// d_err_out: [3][width] ints), phase 1: alpha (srccomps 4, alphabits 1 or 4 only; [width] ints).  d_err_in: the errors the
// rows above sent down (NULL for the image's first rows in phase 0; in phase 1 the image's first rows take the seed the
// LAST rows' phase 0 left in their d_err_out).  d_err_out: what these rows send below -- or, for phase 0 of the image's
// last rows, the alpha seed in its first `width` ints.
cudaError_t launch_floyd_rows(const void *d_src_rows, int srccomps, int alphabits, int width, int image_height, int y0, int rows,
		int phase, const int *d_err_in, int *d_err_out, void *d_reduced_rows, void *d_workspace, cudaStream_t stream)
{
	if (width <= 0 || rows <= 0)
		return cudaSuccess;
	const int bands = (rows + 31) / 32;
	const int nch = phase == 0 ? 3 : 1;
	FloydArgs a;
	a.src = (const uint8_t *) d_src_rows;
	a.out = (uint32_t *) d_reduced_rows;
	a.width = width;
	a.height = rows;
	a.srccomps = srccomps;
	a.alphabits = alphabits;
	a.ticket = (int *) d_workspace;
	int *scratch_seed = a.ticket + 64;
	a.boundary = (uint2 *) (((uintptr_t) (scratch_seed + width) + 255) & ~(uintptr_t) 255);
	a.seed = d_err_in;
	a.image_last = y0 + rows >= image_height;
	a.more_below = !a.image_last;
	a.image_odd = image_height & 1;
	a.alpha_seed = (phase == 0 && a.image_last && d_err_out) ? d_err_out : scratch_seed;
	cudaError_t e = cudaMemsetAsync(a.ticket, 0, 256, stream);
	if (e == cudaSuccess)
		e = cudaMemsetAsync(a.boundary, 0, (size_t) bands * width * nch * sizeof(uint2), stream);
	if (e != cudaSuccess)
		return e;
	if (phase == 0)
		floyd_kernel<false, 3><<<bands, 96, 0, stream>>>(a);
	else if (alphabits == 1)
		floyd_kernel<true, 7><<<bands, 32, 0, stream>>>(a);
	else
		floyd_kernel<true, 4><<<bands, 32, 0, stream>>>(a);
	if (a.more_below && d_err_out) {
		const int n = nch * width;
		boundary_errors_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a.boundary + (size_t) (bands - 1) * nch * width, n, d_err_out);
	}
	return cudaGetLastError();
}

} // namespace s2tc
