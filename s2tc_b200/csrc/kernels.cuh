// kernels.cuh -- declarations shared by the CUDA translation units: the image view the encode
// kernels read from, the block loader, and the host-callable launch wrappers.
#pragma once

#include <cuda_runtime.h>

#include "block_core.cuh"
#include "dither_core.cuh"
#include "glibc_rand.cuh"

namespace s2tc {

// A horizontal slab of an image in device memory.  `base` addresses texel row 0 of the slab; rows are
// tightly packed (width * bytes_per_texel), as tx_compress_dxtn receives them
// (ref s2tc_libtxc_dxtn.cpp:146, s2tc_algorithm.cpp:1274).
struct ImageView {
	const uint8_t *base;
	int width;     // texels per row
	int rows;      // texel rows in the slab
	int fmt;       // SrcFormat
	int alphabits; // 1 / 4 / 8, used when fmt is a raw format (fused DITHER_NONE)
	int blocks_w;  // ceil(width / 4)
	int blocks_h;  // ceil(rows / 4)
	// a batch of `images` equally sized images (mip levels of many textures): image k starts image_bytes * k after base
	// and its blocks go out_image_bytes * k after the output base; 1 image for everything else
	int images;
	size_t image_bytes, out_image_bytes;
};

inline ImageView make_view(const void *base, int width, int rows, int fmt, int alphabits)
{
	ImageView v;
	v.base = (const uint8_t *) base;
	v.width = width;
	v.rows = rows;
	v.fmt = fmt;
	v.alphabits = alphabits;
	v.blocks_w = (width + 3) / 4;
	v.blocks_h = (rows + 3) / 4;
	v.images = 1;
	v.image_bytes = v.out_image_bytes = 0;
	return v;
}

inline ImageView make_batch_view(const void *base, int width, int rows, int fmt, int alphabits, int images, size_t image_bytes,
		size_t out_image_bytes)
{
	ImageView v = make_view(base, width, rows, fmt, alphabits);
	v.images = images;
	v.image_bytes = image_bytes;
	v.out_image_bytes = out_image_bytes;
	return v;
}

inline long long view_blocks(const ImageView &v) { return (long long) v.blocks_w * v.blocks_h * v.images; }

#if defined(__CUDACC__)
// Block number t of a (possibly batched) view: moves v to the image that holds it, returns the block's number inside that
// image and, in out_off, where that image's blocks start in the output.
__device__ __forceinline__ int select_image(ImageView &v, int t, size_t &out_off)
{
	out_off = 0;
	if (v.images > 1) {
		const int per = v.blocks_w * v.blocks_h;
		const int img = t / per;
		t -= img * per;
		v.base += (size_t) img * v.image_bytes;
		out_off = (size_t) img * v.out_image_bytes;
	}
	return t;
}

// fused DITHER_NONE on a raw RGBA word (ref s2tc_algorithm.cpp:1274-1297)
__device__ __forceinline__ uint32_t reduce_word(uint32_t w, int alphabits)
{
	uint32_t rb = (w >> 3) & 0x001F001Fu;
	uint32_t g = (w >> 2) & 0x00003F00u;
	uint32_t a = alphabits == 8 ? (w & 0xFF000000u) : (alphabits == 4 ? ((w >> 4) & 0x0F000000u) : ((w >> 7) & 0x01000000u));
	return rb | g | a;
}

// Loads block (bx, by) of the view into registers as reduced texels.  Full-width blocks of 4-byte
// formats whose rows are 16-byte aligned take one 128-bit load per texel row; everything else
// (edge blocks, odd widths, 3-byte texels) goes texel by texel.
__device__ __forceinline__ void load_block(const ImageView &v, int bx, int by, Block &b)
{
	const int x0 = bx * 4, y0 = by * 4;
	const int w = min(4, v.width - x0), h = min(4, v.rows - y0);
	b.valid = valid_mask(w, h);
	if (v.fmt != kSrcRGB8) {
		const size_t pitch = (size_t) v.width * 4;
		const uint8_t *p = v.base + (size_t) y0 * pitch + (size_t) x0 * 4;
		const bool vec = w == 4 && ((pitch | (size_t) v.base) & 15) == 0;
#pragma unroll
		for (int y = 0; y < 4; ++y) {
			uint32_t t[4] = {0, 0, 0, 0};
			if (y < h) {
				const uint8_t *row = p + (size_t) y * pitch;
				if (vec) {
					const uint4 q = __ldg(reinterpret_cast<const uint4 *>(row));
					t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
				} else {
#pragma unroll
					for (int x = 0; x < 4; ++x)
						if (x < w)
							t[x] = __ldg(reinterpret_cast<const uint32_t *>(row) + x);
				}
			}
#pragma unroll
			for (int x = 0; x < 4; ++x)
				b.px[y * 4 + x] = v.fmt == kSrcRGBA8 ? reduce_word(t[x], v.alphabits) : t[x];
		}
	} else {
		const size_t pitch = (size_t) v.width * 3;
		const uint8_t *p = v.base + (size_t) y0 * pitch + (size_t) x0 * 3;
		const uint32_t ones = ((1u << v.alphabits) - 1u) << 24;
#pragma unroll
		for (int y = 0; y < 4; ++y)
#pragma unroll
			for (int x = 0; x < 4; ++x) {
				uint32_t t = 0;
				if (y < h && x < w) {
					const uint8_t *q = p + (size_t) y * pitch + x * 3;
					t = (uint32_t) (__ldg(q) >> 3) | ((uint32_t) (__ldg(q + 1) >> 2) << 8) | ((uint32_t) (__ldg(q + 2) >> 3) << 16) | ones;
				}
				b.px[y * 4 + x] = t;
			}
	}
}
#endif

// per-translation-unit lookup tables (colordist.cuh: S2TC_USE_SRGB_MIXED_LUT); a context fills all of them once
cudaError_t init_luts_fast(cudaStream_t stream);
cudaError_t init_luts_finish(cudaStream_t stream);
cudaError_t init_luts_search(cudaStream_t stream);
cudaError_t init_luts_search16_dxt1(cudaStream_t stream);
cudaError_t init_luts_search16_dxt3(cudaStream_t stream);
cudaError_t init_luts_search16_dxt5(cudaStream_t stream);
inline cudaError_t init_all_luts(cudaStream_t stream)
{
	cudaError_t (*const fns[])(cudaStream_t) = {init_luts_fast, init_luts_finish, init_luts_search, init_luts_search16_dxt1,
			init_luts_search16_dxt3, init_luts_search16_dxt5};
	for (auto fn : fns)
		if (cudaError_t e = fn(stream))
			return e;
	return cudaSuccess;
}

// ---- launch wrappers (defined in the .cu files); all asynchronous on `stream` -------------------

// MODE_FAST: candidates + refinement + packing in one pass, one thread per block.
cudaError_t launch_fast_encode(int dxt, int cd, int refine, const ImageView &v, void *d_out, cudaStream_t stream);

// MODE_NORMAL with random candidates (nrandom > 0): candidate generation + the c0/c1 (and DXT5 a0/a1) pair search, one warp
// per chunk of kSearchChunkBlocks consecutive blocks (kernels_search.cu).
// d_windows: the rand() window of every chunk at its first draw, [31][chunks] words, from launch_rand_windows with a plan
// whose stride is kSearchChunkBlocks * draws_per_block.
// d_ends: [blocks] uint2 {c0_565 | c1_565 << 16, a0 | a1 << 8}
constexpr int kSearchChunkBlocks = 32;
inline size_t rand_windows_bytes(size_t nblocks)
{
	return ((nblocks + kSearchChunkBlocks - 1) / kSearchChunkBlocks) * kLag * sizeof(uint32_t) + 256;
}
// plan: mapped pinned host memory -> device memory, by a kernel (no DMA engine involved)
cudaError_t launch_plan_upload(const RandPlan *mapped_host_plan, RandPlan *d_plan, cudaStream_t stream);
cudaError_t launch_rand_windows(const RandPlan *d_plan, unsigned nsegments, uint32_t *d_windows, cudaStream_t stream);
cudaError_t launch_pair_search(int dxt, int cd, int nrandom, const ImageView &v, const uint32_t *d_windows, uint2 *d_ends,
		cudaStream_t stream);
// largest nrandom the search kernel can hold in shared memory
int pair_search_max_nrandom();

// MODE_NORMAL with at most 16 candidates (nrandom <= 0): gather + colour search (+ DXT5 alpha search), one thread
// per block, the distance matrix in registers (search16.inl; one TU per DXT mode).  Output as launch_pair_search.
cudaError_t launch_search16_dxt1(int cd, const ImageView &v, uint2 *d_ends, cudaStream_t stream);
cudaError_t launch_search16_dxt3(int cd, const ImageView &v, uint2 *d_ends, cudaStream_t stream);
cudaError_t launch_search16_dxt5(int cd, const ImageView &v, uint2 *d_ends, cudaStream_t stream);
inline cudaError_t launch_search16(int dxt, int cd, const ImageView &v, uint2 *d_ends, cudaStream_t stream)
{
	return dxt == kDxt1 ? launch_search16_dxt1(cd, v, d_ends, stream)
			: dxt == kDxt3 ? launch_search16_dxt3(cd, v, d_ends, stream)
			: launch_search16_dxt5(cd, v, d_ends, stream);
}

// MODE_NORMAL step 3: refinement + packing from the searched endpoints, one thread per block.
cudaError_t launch_finish(int dxt, int cd, int refine, const ImageView &v, const uint2 *d_ends, void *d_out,
		cudaStream_t stream);

// 565 pre-pass.
cudaError_t launch_prepass_none(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		cudaStream_t stream);
// DITHER_SIMPLE over `npixels` texels in raster order.  d_carry: 4 ints (r,g,b,a) carried in and
// updated to the carry out.  d_maps: workspace of dither_workspace_bytes(npixels).
size_t dither_workspace_bytes(size_t npixels);
// maps_ready: the workspace already holds this range's chunk/tile maps (left by launch_dither_summary on the same
// texels), so phase 1 is skipped.
cudaError_t launch_prepass_simple(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		int *d_carry, void *d_workspace, bool maps_ready, cudaStream_t stream);
int prepass_simple_launches(size_t npixels, bool maps_ready); // kernel launches the call above makes
// `images` images of npixels texels each, back to back, each from carry 0; d_zero_carry: 4 zeroed ints of scratch;
// workspace: dither_workspace_bytes(npixels * images)
cudaError_t launch_prepass_simple_batch(const void *d_src, int srccomps, int alphabits, size_t npixels, int images, void *d_reduced,
		int *d_zero_carry, void *d_workspace, cudaStream_t stream);
int prepass_simple_batch_launches(size_t npixels, int images);
constexpr int kDitherSummaryLaunches = 3;                     // maps + two scan launches
// transfer maps only (for sharding a carry chain across GPUs): d_summary receives 4 ByteMaps
cudaError_t launch_dither_summary(const void *d_src, int srccomps, int alphabits, size_t npixels,
		ByteMap *d_summary, void *d_workspace, cudaStream_t stream);

// carry entering a range = the first `rank` summaries of d_maps (4 ByteMaps each) applied in order to a zero carry
cudaError_t launch_fold_carry(const ByteMap *d_maps, int rank, int srccomps, int alphabits, int *d_carry, cudaStream_t stream);
cudaError_t launch_identity_maps(ByteMap *d_maps, int count, int srccomps, int alphabits, cudaStream_t stream);
// the same fold over `count` summaries starting from the carry in d_carry_in (device; NULL = zero; may alias d_carry)
cudaError_t launch_fold_carry_from(const ByteMap *d_maps, int count, int srccomps, int alphabits, const int *d_carry_in, int *d_carry,
		cudaStream_t stream);

// DITHER_FLOYDSTEINBERG over a whole width x height image (a 2-D recurrence: it cannot start in the middle)
size_t floyd_workspace_bytes(int width, int height);
cudaError_t launch_prepass_floyd(const void *d_src, int srccomps, int alphabits, int width, int height, void *d_reduced,
		void *d_workspace, cudaStream_t stream);
// one pass (phase 0: r, g, b; phase 1: alpha) over `rows` texel rows starting at row y0 of an image of image_height rows; the error
// rows that cross the cuts travel through d_err_in / d_err_out (kernels_floyd.cu).  Workspace: floyd_workspace_bytes(width, rows).
cudaError_t launch_floyd_rows(const void *d_src_rows, int srccomps, int alphabits, int width, int image_height, int y0, int rows,
		int phase, const int *d_err_in, int *d_err_out, void *d_reduced_rows, void *d_workspace, cudaStream_t stream);

// one mip step of `images` RGBA8 images stored back to back (w x h -> max(w/2,1) x max(h/2,1) each, results back to back);
// d_out must not alias d_in
cudaError_t launch_mip_reduce(const void *d_in, int w, int h, void *d_out, cudaStream_t stream, int images = 1);

// S2TC decode of a whole image to RGBA8 (tightly packed blocks in, width*height*4 bytes out)
cudaError_t launch_decode(int dxt, const void *d_blocks, int width, int height, void *d_rgba, cudaStream_t stream);

// S3TC -> S2TC transcode, in place.
cudaError_t launch_transcode(int dxt, void *d_blocks, size_t nblocks, cudaStream_t stream);

// measurement aid: `ctas` CTAs of 256 threads each run iters * 8 (min, add) pairs per thread; mode: see kernels_misc.cu
cudaError_t launch_int32_peak(int mode, int iters, int ctas, int *d_sink, cudaStream_t stream);

} // namespace s2tc
