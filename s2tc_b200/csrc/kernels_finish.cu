// kernels_finish.cu -- MODE_NORMAL step 3: takes the endpoints the pair search chose for every
// block and produces the finished DXT block: equal-endpoint fix-ups, NEVER/ALWAYS/LOOP refinement,
// DXT3/DXT5 alpha and packing (reference: s2tc_algorithm.cpp:1010-1107).  One thread per block; the
// memory pattern is that of kernels_fast.cu plus one coalesced 8-byte endpoint read per block.
#define S2TC_USE_SRGB_MIXED_LUT
#include "kernels.cuh"

namespace s2tc {

#ifndef S2TC_FINISH_MINBLOCKS
#define S2TC_FINISH_MINBLOCKS 6 // 80 registers: measured 4 / 6 / 8 CTAs per SM -> 0.66 / 0.63 / 0.70 ms on config 2 (8 spills)
#endif
template <int DXT, int CD>
__global__ void __launch_bounds__(128, S2TC_FINISH_MINBLOCKS) finish_kernel(ImageView v, int refine, const uint2 *__restrict__ ends, uint8_t *out)
{
	const int nblocks = v.blocks_w * v.blocks_h * v.images;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nblocks)
		return;
	size_t out_off;
	const int ti = select_image(v, t, out_off);
	const int by = ti / v.blocks_w, bx = ti - by * v.blocks_w;
	Block b;
	load_block(v, bx, by, b);
	const uint2 e = __ldg(ends + t);
	const uint32_t c0 = from565(e.x & 0xFFFFu), c1 = from565(e.x >> 16);
	const int a0 = (int) (e.y & 0xFFu), a1 = (int) ((e.y >> 8) & 0xFFu);
	uint32_t w[4];
	finish_block<DXT, CD>(b, refine, c0, c1, a0, a1, w);
	if (DXT == kDxt1)
		reinterpret_cast<uint2 *>(out + out_off)[ti] = make_uint2(w[0], w[1]);
	else
		reinterpret_cast<uint4 *>(out + out_off)[ti] = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int DXT>
static cudaError_t launch_finish_dxt(int cd, int refine, const ImageView &v, const uint2 *ends, void *d_out, cudaStream_t stream)
{
	const int nblocks = (int) view_blocks(v);
	if (nblocks == 0)
		return cudaSuccess;
	const dim3 block(128), grid((nblocks + 127) / 128);
	uint8_t *out = (uint8_t *) d_out;
	switch (cd) {
	case kRGB: finish_kernel<DXT, kRGB><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kYUV: finish_kernel<DXT, kYUV><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kSRGB: finish_kernel<DXT, kSRGB><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kSRGB_MIXED: finish_kernel<DXT, kSRGB_MIXED><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kAVG: finish_kernel<DXT, kAVG><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kWAVG: finish_kernel<DXT, kWAVG><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kW0AVG: finish_kernel<DXT, kW0AVG><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	case kNORMALMAP: finish_kernel<DXT, kNORMALMAP><<<grid, block, 0, stream>>>(v, refine, ends, out); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

cudaError_t launch_finish(int dxt, int cd, int refine, const ImageView &v, const uint2 *d_ends, void *d_out, cudaStream_t stream)
{
	switch (dxt) {
	case kDxt1: return launch_finish_dxt<kDxt1>(cd, refine, v, d_ends, d_out, stream);
	case kDxt3: return launch_finish_dxt<kDxt3>(cd, refine, v, d_ends, d_out, stream);
	default: return launch_finish_dxt<kDxt5>(cd, refine, v, d_ends, d_out, stream);
	}
}

S2TC_DEFINE_LUT_INIT(init_luts_finish)

} // namespace s2tc
