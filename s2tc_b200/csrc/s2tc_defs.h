// s2tc_defs.h -- enums and small helpers shared by every translation unit of the encoder.
//
// The numeric values of the enums are the reference's (s2tc_algorithm.h:31-63) because they cross
// the drop-in boundary unchanged (s2tc_encode_block_func(DxtMode, ColorDistMode, int, RefinementMode)).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define S2TC_HD __host__ __device__ __forceinline__
#define S2TC_D __device__ __forceinline__
#else
#define S2TC_HD inline
#define S2TC_D inline
#endif

namespace s2tc {

enum Dither : int { kDitherNone = 0, kDitherSimple = 1, kDitherFloyd = 2 };
enum Dxt : int { kDxt1 = 0, kDxt3 = 1, kDxt5 = 2 };
enum Refine : int { kRefineNever = 0, kRefineAlways = 1, kRefineLoop = 2 };
enum ColorDist : int { kRGB = 0, kYUV, kSRGB, kSRGB_MIXED, kAVG, kWAVG, kW0AVG, kNORMALMAP, kNumColorDist };

// How the texels handed to an encode kernel are stored.
enum SrcFormat : int {
	kSrcReduced = 0, // 4 B/pixel {r5, g6, b5, a(1|4|8 bit)}: output of the 565 pre-pass (any dither mode)
	kSrcRGBA8 = 1,   // raw 4 B/pixel; DITHER_NONE is fused into the load (shifts only)
	kSrcRGB8 = 2     // raw 3 B/pixel; DITHER_NONE fused, alpha = all ones
};

S2TC_HD int block_bytes(int dxt) { return dxt == kDxt1 ? 8 : 16; }
S2TC_HD int alpha_bits(int dxt) { return dxt == kDxt1 ? 1 : (dxt == kDxt3 ? 4 : 8); }
// rand() draws one block consumes (reference s2tc_algorithm.cpp:984-992): r,g,b and, for DXT5, a
S2TC_HD int draws_per_block(int dxt, int nrandom) { return nrandom > 0 ? nrandom * (dxt == kDxt5 ? 4 : 3) : 0; }

// Normalisation of out-of-range settings exactly as the reference's dispatch does
// (s2tc_algorithm.cpp:1120 refine -> ALWAYS, :1156 dxt -> DXT5, :1183 cd -> WAVG).
S2TC_HD int norm_refine(int r) { return (r == kRefineNever || r == kRefineLoop) ? r : kRefineAlways; }
S2TC_HD int norm_dxt(int d) { return (d == kDxt1 || d == kDxt3) ? d : kDxt5; }
S2TC_HD int norm_cd(int c) { return (c >= 0 && c < kNumColorDist) ? c : kWAVG; }
// MODE_FAST is taken iff nrandom < 0 and the metric is not NORMALMAP (s2tc_algorithm.cpp:1127-1143)
S2TC_HD bool is_fast_mode(int cd, int nrandom) { return nrandom < 0 && cd != kNORMALMAP; }

} // namespace s2tc
