// transcode_core.cuh -- S3TC -> S2TC block transcode (reference s2tc_from_s3tc.cpp:77-190), host+device.
//
// S2TC never uses the interpolated palette entries, so an S3TC block is converted by replacing every
// interpolated index with one of the two endpoints according to a fixed checkerboard, and by
// putting the endpoints in S2TC's canonical order (DXT1: c0 <= c1, i.e. always the 3-colour +
// transparent mode; DXT3/DXT5 colour: c0 > c1; DXT5 alpha: a0 <= a1, the 6-value mode with explicit
// 0/255).  All of it is a handful of bitwise operations per block.
#pragma once

#include "s2tc_defs.h"

namespace s2tc {

constexpr uint32_t kChecker2 = 0x22882288u;             // odd texels of a 4x4 checkerboard, as bit 1 of each 2-bit code
constexpr uint64_t kOnes3 = 01111111111111111ull;       // bit 0 of each 3-bit code
constexpr uint64_t kChecker3 = 00101101001011010ull;    // the same checkerboard, as bit 0 of each 3-bit code

// colour words of a DXT1 block: ends = c0 | c1 << 16, idx = 2-bit codes (ref convert_dxt1a, :113-147)
S2TC_HD void transcode_color_dxt1(uint32_t &ends, uint32_t &idx)
{
	const uint32_t c0 = ends & 0xFFFFu, c1 = ends >> 16;
	if (c1 >= c0) {
		// already "3 colours + transparent": 2 -> 0/1 by checkerboard, 3 stays transparent
		idx = (idx & ~((~idx & 0x55555555u) << 1)) | ((idx & kChecker2) >> 1);
	} else {
		// 4-colour mode: 2 and 3 -> 0/1, then swap the endpoints and invert the codes
		idx = (idx & ((~idx & 0xAAAAAAAAu) >> 1)) | ((idx & kChecker2) >> 1);
		ends = c1 | (c0 << 16);
		idx ^= 0x55555555u;
	}
}

// colour words of a DXT3/DXT5 block, where code 3 is never transparent (ref convert_dxt1, :77-111)
S2TC_HD void transcode_color_opaque(uint32_t &ends, uint32_t &idx)
{
	const uint32_t c0 = ends & 0xFFFFu, c1 = ends >> 16;
	idx = (idx & ((~idx & 0xAAAAAAAAu) >> 1)) | ((idx & kChecker2) >> 1);
	if (c1 >= c0) {
		ends = c1 | (c0 << 16);
		idx ^= 0x55555555u;
	}
}

// alpha half of a DXT5 block as a little-endian 64-bit word: a0, a1, 48 bits of codes (ref convert_dxt5, :149-190)
S2TC_HD uint64_t transcode_alpha_dxt5(uint64_t blk)
{
	uint32_t a0 = (uint32_t) (blk & 0xFF), a1 = (uint32_t) ((blk >> 8) & 0xFF);
	uint64_t px = blk >> 16;
	if (a1 >= a0) {
		// 6-value mode: codes 2..5 interpolate -> 0/1 by checkerboard; 6 (=0) and 7 (=255) stay
		const uint64_t sel = (px >> 1) ^ (px >> 2);
		px = (px & ~((sel & kOnes3) * 7)) | (sel & kChecker3);
	} else {
		// 8-value mode: codes 2..7 interpolate; then swap so that a0 <= a1
		const uint64_t sel = (px >> 1) | (px >> 2);
		px = (px & ~((sel & kOnes3) * 7)) | (sel & kChecker3);
		const uint32_t t = a0; a0 = a1; a1 = t;
		px ^= kOnes3;
	}
	return (uint64_t) a0 | ((uint64_t) a1 << 8) | ((px & 0xFFFFFFFFFFFFull) << 16);
}

} // namespace s2tc
