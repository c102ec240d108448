// colordist.cuh -- the eight S2TC colour metrics and the alpha metric, host+device.
//
// Each metric is split into a per-colour "feature" (everything that depends on one colour only:
// squares for SRGB, the sqrt-luma triple for SRGB_MIXED, the unit vector for NORMALMAP) and a
// pairwise distance on features.  The results are bit-identical to the reference's
// color_dist_* functions (s2tc_algorithm.cpp:215-361); the split only removes recomputation
// when one colour meets many others (distance matrix rows, 16 texels against two endpoints).
//
// Exactness notes:
//  * SRGB overflows int32 for saturated pairs and the compiled reference wraps: products are
//    done in uint32_t and shifted arithmetically.  SRGB is also NOT symmetric in its arguments
//    (SHRR rounds -x and x differently), so callers keep the reference's argument order.
//  * SRGB_MIXED needs a correctly rounded sqrtf and a separately rounded +0.5f.
//  * NORMALMAP is a chain of individually rounded fp32 operations: no FMA contraction, IEEE
//    division and square root.  On the device this is spelled with the _rn intrinsics so that no
//    compiler flag can change it.
#pragma once

#include "s2tc_defs.h"

#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace s2tc {

// ---- individually rounded fp32 primitives --------------------------------------------------
#if defined(__CUDA_ARCH__)
S2TC_D float f_add(float a, float b) { return __fadd_rn(a, b); }
S2TC_D float f_sub(float a, float b) { return __fsub_rn(a, b); }
S2TC_D float f_mul(float a, float b) { return __fmul_rn(a, b); }
S2TC_D float f_div(float a, float b) { return __fdiv_rn(a, b); }
S2TC_D float f_sqrt(float a) { return __fsqrt_rn(a); }
S2TC_D int f_trunc(float a) { return __float2int_rz(a); }
#else
// host build (tests only) is compiled with -ffp-contract=off; volatile blocks reassociation
inline float f_add(float a, float b) { volatile float r = a + b; return r; }
inline float f_sub(float a, float b) { volatile float r = a - b; return r; }
inline float f_mul(float a, float b) { volatile float r = a * b; return r; }
inline float f_div(float a, float b) { volatile float r = a / b; return r; }
inline float f_sqrt(float a) { volatile float r = sqrtf(a); return r; }
inline int f_trunc(float a) { return (int) a; }
#endif

// SHRR(a, n) = (a + (1 << (n-1))) >> n with wrapping add and arithmetic shift (ref :215)
S2TC_HD int shrr(int a, int n) { return (int) ((uint32_t) a + (1u << (n - 1))) >> n; }
S2TC_HD int wmul(int a, int b) { return (int) ((uint32_t) a * (uint32_t) b); }
S2TC_HD int wadd(int a, int b) { return (int) ((uint32_t) a + (uint32_t) b); }

S2TC_HD int alpha_dist(int a, int b) { return (a - b) * (a - b); } // ref :358-361

// ---- texel packing ----------------------------------------------------------------------
// A reduced texel is the 4 bytes {r5, g6, b5, a} read as a little-endian word.
S2TC_HD int px_r(uint32_t p) { return (int) (p & 0xFF); }
S2TC_HD int px_g(uint32_t p) { return (int) ((p >> 8) & 0xFF); }
S2TC_HD int px_b(uint32_t p) { return (int) ((p >> 16) & 0xFF); }
S2TC_HD int px_a(uint32_t p) { return (int) (p >> 24); }
S2TC_HD uint32_t px_make(int r, int g, int b, int a = 0)
{
	return (uint32_t) r | ((uint32_t) g << 8) | ((uint32_t) b << 16) | ((uint32_t) a << 24);
}
S2TC_HD uint32_t px_rgb(uint32_t p) { return p & 0x00FFFFFFu; }

// ---- metrics ----------------------------------------------------------------------------
struct FeatRGB { int r, g, b; };
struct FeatF3 { float x, y, z; };

template <int CD> struct Metric;

// linear metrics on the raw 5/6/5 differences
template <int CD> struct LinearMetric {
	typedef FeatRGB Feat;
	static constexpr bool kMayBeNegative = false;
	static S2TC_HD Feat feat(uint32_t p) { return Feat{px_r(p), px_g(p), px_b(p)}; }
};

// AVG / W0AVG / WAVG are sums of squares of integer-weighted channel differences (weights 2,1,2 / 1,1,1 / 2,2,1 on
// dr, dg, db).  The feature is the colour with its channels pre-scaled, one per byte (all < 128); a distance is then one
// borrow-free per-byte subtraction and one 4-way dot product of the difference with itself (IDP.4A): 3 instructions
// instead of ~9, in the fast kernel, the refinement passes and the distance-matrix fills alike.
struct FeatBytes { uint32_t v; };

S2TC_HD int dot4_self(uint32_t d) // sum of squares of the four signed bytes of d
{
#if defined(__CUDA_ARCH__)
	return __dp4a((int) d, (int) d, 0);
#else
	int s = 0;
	for (int i = 0; i < 4; ++i) {
		const int b = (int) (signed char) (d >> (8 * i));
		s += b * b;
	}
	return s;
#endif
}

template <uint32_t DOUBLED /* mask of the channels weighted 2 */> struct SquareMetric {
	typedef FeatBytes Feat;
	static constexpr bool kMayBeNegative = false;
	static S2TC_HD Feat feat(uint32_t p)
	{
		const uint32_t rgb = p & 0x00FFFFFFu;
		return Feat{rgb + (rgb & DOUBLED)}; // r <= 31, g <= 63, b <= 31: doubling never carries into the next byte
	}
	static S2TC_HD int dist(const FeatBytes &a, const FeatBytes &b)
	{
		// bytes of a.v and b.v are < 128, so (a | 0x80) - b cannot borrow across bytes; the xor restores the sign bit
		return dot4_self(((a.v | 0x80808080u) - b.v) ^ 0x80808080u);
	}
};

template <> struct Metric<kAVG> : SquareMetric<0x00FF00FFu> {};   // ref :217-223  4dr^2 +  dg^2 + 4db^2 <= 11657
template <> struct Metric<kW0AVG> : SquareMetric<0x00000000u> {}; // ref :225-232   dr^2 +  dg^2 +  db^2 <= 5891
template <> struct Metric<kWAVG> : SquareMetric<0x0000FFFFu> {};  // ref :234-241  4dr^2 + 4dg^2 +  db^2 <= 20681
template <> struct Metric<kYUV> : LinearMetric<kYUV> { // ref :243-254
	static S2TC_HD int dist(const FeatRGB &a, const FeatRGB &b)
	{
		int dr = a.r - b.r, dg = a.g - b.g, db = a.b - b.b;
		int y = dr * 60 + dg * 59 + db * 22;
		int u = dr * 202 - y;
		int v = db * 202 - y;
		return ((y * y) << 1) + shrr(u * u, 3) + shrr(v * v, 4);
	}
};
template <> struct Metric<kRGB> : LinearMetric<kRGB> { // ref :256-267
	static S2TC_HD int dist(const FeatRGB &a, const FeatRGB &b)
	{
		int dr = a.r - b.r, dg = a.g - b.g, db = a.b - b.b;
		int y = dr * 42 + dg * 72 + db * 14;
		int u = dr * 202 - y;
		int v = db * 202 - y;
		return ((y * y) << 1) + shrr(u * u, 3) + shrr(v * v, 4);
	}
};

template <> struct Metric<kSRGB> { // ref :269-283
	typedef FeatRGB Feat; // squares of the components
	static constexpr bool kMayBeNegative = true;
	static S2TC_HD Feat feat(uint32_t p)
	{
		int r = px_r(p), g = px_g(p), b = px_b(p);
		return Feat{r * r, g * g, b * b};
	}
	static S2TC_HD int dist(const FeatRGB &a, const FeatRGB &b)
	{
		int dr = a.r - b.r, dg = a.g - b.g, db = a.b - b.b;
		int y = dr * 84 + dg * 72 + db * 28;
		int u = dr * 409 - y;
		int v = db * 409 - y;
		int sy = wmul(shrr(y, 3), shrr(y, 4));
		int su = wmul(shrr(u, 3), shrr(u, 4));
		int sv = wmul(shrr(v, 3), shrr(v, 4));
		return wadd(wadd(shrr(sy, 4), shrr(su, 8)), shrr(sv, 9));
	}
};

// Y(c) of SRGB_MIXED (ref :292-294): int(sqrtf(37 * (84 r^2 + 72 g^2 + 28 b^2)) + 0.5f), individually rounded
S2TC_HD int srgb_mixed_y(int r, int g, int b)
{
	int lin = 37 * (r * r * 84 + g * g * 72 + b * b * 28); // < 2^24: exact in fp32
	return f_trunc(f_add(f_sqrt((float) lin), 0.5f));
}

#if defined(__CUDACC__) && defined(S2TC_USE_SRGB_MIXED_LUT)
// Device builds of the per-block kernels read Y from a table of all 65536 RGB565 colours instead of running the
// sqrtf chain per colour (~18 instructions, 16 + 2 per refinement pass of them per block).  One copy per translation
// unit that opts in (no relocatable device code in this build), filled ON THE DEVICE by the formula above, so the
// values are the kernel's own; S2TC_DEFINE_LUT_INIT(fn) defines the host function that fills this unit's copy and
// s2tc_b200_ctx_create calls all of them.  Texels are always in the 5/6/5 domain here.
static __device__ uint16_t g_srgb_mixed_y[65536];
static __global__ void srgb_mixed_lut_kernel()
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x; // r | g << 5 | b << 11
	g_srgb_mixed_y[idx] = (uint16_t) srgb_mixed_y(idx & 31, (idx >> 5) & 63, idx >> 11);
}
#endif

template <> struct Metric<kSRGB_MIXED> { // ref :285-315
	typedef FeatRGB Feat; // {Y, U, V} of one colour
	static constexpr bool kMayBeNegative = false;
	static S2TC_HD Feat feat(uint32_t p)
	{
		int r = px_r(p), b = px_b(p);
#if defined(__CUDA_ARCH__) && defined(S2TC_USE_SRGB_MIXED_LUT)
		int y = __ldg(&g_srgb_mixed_y[(p & 31u) | ((p >> 3) & 0x7E0u) | ((p >> 5) & 0xF800u)]);
#else
		int y = srgb_mixed_y(r, px_g(p), b);
#endif
		return Feat{y, r * 191 - y, b * 191 - y};
	}
	static S2TC_HD int dist(const FeatRGB &a, const FeatRGB &b)
	{
		int y = a.r - b.r, u = a.g - b.g, v = a.b - b.b;
		return ((y * y) << 3) + shrr(u * u, 1) + shrr(v * v, 2);
	}
};

// the unit vector of NORMALMAP (ref :321-336), every operation individually rounded
S2TC_HD FeatF3 normalmap_dir(int r, int g, int b)
{
	float x = f_sub(f_mul(f_div((float) r, 31.0f), 2.0f), 1.0f);
	float y = f_sub(f_mul(f_div((float) g, 63.0f), 2.0f), 1.0f);
	float z = f_sub(f_mul(f_div((float) b, 31.0f), 2.0f), 1.0f);
	float n = f_add(f_add(f_mul(x, x), f_mul(y, y)), f_mul(z, z));
	if (n > 0) {
		n = f_div(1.0f, f_sqrt(n));
		x = f_mul(x, n);
		y = f_mul(y, n);
		z = f_mul(z, n);
	}
	return FeatF3{x, y, z};
}

#if defined(__CUDACC__) && defined(S2TC_USE_SRGB_MIXED_LUT)
// same scheme as g_srgb_mixed_y: the unit vectors of all 65536 colours (three IEEE divisions, a square root and a
// reciprocal per colour otherwise), 1 MB per opted-in translation unit, L2-resident
static __device__ float4 g_normalmap_dir[65536];
static __global__ void normalmap_lut_kernel()
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x; // r | g << 5 | b << 11
	const FeatF3 d = normalmap_dir(idx & 31, (idx >> 5) & 63, idx >> 11);
	g_normalmap_dir[idx] = make_float4(d.x, d.y, d.z, 0.0f);
}
#define S2TC_DEFINE_LUT_INIT(fn)                                      \
	cudaError_t fn(cudaStream_t stream)                               \
	{                                                                 \
		s2tc::srgb_mixed_lut_kernel<<<256, 256, 0, stream>>>();       \
		s2tc::normalmap_lut_kernel<<<256, 256, 0, stream>>>();        \
		return cudaGetLastError();                                    \
	}
#endif

template <> struct Metric<kNORMALMAP> { // ref :317-354
	typedef FeatF3 Feat; // normalised direction
	static constexpr bool kMayBeNegative = false;
	static S2TC_HD Feat feat(uint32_t p)
	{
#if defined(__CUDA_ARCH__) && defined(S2TC_USE_SRGB_MIXED_LUT)
		const float4 d = __ldg(&g_normalmap_dir[(p & 31u) | ((p >> 3) & 0x7E0u) | ((p >> 5) & 0xF800u)]);
		return Feat{d.x, d.y, d.z};
#else
		return normalmap_dir(px_r(p), px_g(p), px_b(p));
#endif
	}
	static S2TC_HD int dist(const FeatF3 &a, const FeatF3 &b)
	{
		float dx = f_sub(b.x, a.x), dy = f_sub(b.y, a.y), dz = f_sub(b.z, a.z);
		float s = f_add(f_add(f_mul(dx, dx), f_mul(dy, dy)), f_mul(dz, dz));
		return f_trunc(f_mul(100000.0f, s));
	}
};

// convenience: distance between two packed texels (used off the hot loops and by tests)
template <int CD> S2TC_HD int color_dist(uint32_t a, uint32_t b)
{
	return Metric<CD>::dist(Metric<CD>::feat(a), Metric<CD>::feat(b));
}

} // namespace s2tc
