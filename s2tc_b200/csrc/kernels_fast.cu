// kernels_fast.cu -- MODE_FAST encoder (S2TC_RANDOM_COLORS < 0, every metric but NORMALMAP):
// darkest/brightest endpoint pick, refinement and packing fused in ONE pass over the texels,
// one thread per 4x4 block (reference path: s2tc_algorithm.cpp:879-935 + :1010-1107 per block,
// driven by the loops in s2tc_libtxc_dxtn.cpp:246-294).
//
// Memory behaviour: a warp reads 32 horizontally adjacent blocks, i.e. one contiguous 512-byte
// segment per texel row (4 x LDG.128 per thread) and writes 256 / 512 contiguous bytes
// (STG.64 / STG.128).  Algorithmic traffic is 64 B in + 8|16 B out per block; the rest is integer
// work (~500-600 ops per block), so this kernel sits near the HBM/ALU balance point.
#define S2TC_USE_SRGB_MIXED_LUT
#include "kernels.cuh"

namespace s2tc {

template <int DXT, int CD>
__global__ void __launch_bounds__(128) fast_encode_kernel(ImageView v, int refine, uint8_t *out)
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
	uint32_t c0, c1;
	int a0, a1;
	fast_candidates<DXT, CD>(b, c0, c1, a0, a1);
	uint32_t w[4];
	finish_block<DXT, CD>(b, refine, c0, c1, a0, a1, w);
	if (DXT == kDxt1)
		reinterpret_cast<uint2 *>(out + out_off)[ti] = make_uint2(w[0], w[1]);
	else
		reinterpret_cast<uint4 *>(out + out_off)[ti] = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int DXT>
static cudaError_t launch_fast_dxt(int cd, int refine, const ImageView &v, void *d_out, cudaStream_t stream)
{
	const int nblocks = (int) view_blocks(v);
	if (nblocks == 0)
		return cudaSuccess;
	const dim3 block(128), grid((nblocks + 127) / 128);
	uint8_t *out = (uint8_t *) d_out;
	switch (cd) {
	case kRGB: fast_encode_kernel<DXT, kRGB><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kYUV: fast_encode_kernel<DXT, kYUV><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kSRGB: fast_encode_kernel<DXT, kSRGB><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kSRGB_MIXED: fast_encode_kernel<DXT, kSRGB_MIXED><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kAVG: fast_encode_kernel<DXT, kAVG><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kWAVG: fast_encode_kernel<DXT, kWAVG><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kW0AVG: fast_encode_kernel<DXT, kW0AVG><<<grid, block, 0, stream>>>(v, refine, out); break;
	default: return cudaErrorInvalidValue; // NORMALMAP never takes MODE_FAST (ref :1131-1139)
	}
	return cudaGetLastError();
}

cudaError_t launch_fast_encode(int dxt, int cd, int refine, const ImageView &v, void *d_out, cudaStream_t stream)
{
	switch (dxt) {
	case kDxt1: return launch_fast_dxt<kDxt1>(cd, refine, v, d_out, stream);
	case kDxt3: return launch_fast_dxt<kDxt3>(cd, refine, v, d_out, stream);
	default: return launch_fast_dxt<kDxt5>(cd, refine, v, d_out, stream);
	}
}

S2TC_DEFINE_LUT_INIT(init_luts_fast)

} // namespace s2tc
