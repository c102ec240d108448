// libtxc_dxtn.cpp -- the reference's public surface on top of the s2tc_b200_* C ABI:
//   tx_compress_dxtn, fetch_2d_texel_*            (include/s2tc_b200_txc_dxtn.h; ref txc_dxtn.h:38-49)
//   s2tc_encode_block_func, rgb565_image           (include/s2tc_b200_algorithm.h; ref s2tc_algorithm.h:38,65-66)
// Host code only parses settings, picks the device context and forwards; all encoding runs in the
// CUDA kernels.  The per-texel fetchers are the decode side of the ABI (outside the encode hot path,
// SURVEY.md C5) and are plain host functions, as in the reference: a drop-in .so must export them
// because s2tc_decompress refuses a library without them (ref s2tc_decompress.c:54-63).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>

#include "../../include/s2tc_b200.h"
#include "../../include/s2tc_b200_algorithm.h"
#include "../../include/s2tc_b200_txc_dxtn.h"

namespace {

// Settings from the environment, re-read on every call like the reference does
// (ref s2tc_libtxc_dxtn.cpp:156-216): case-insensitive names, a bad value warns on stderr and keeps
// the default, S2TC_RANDOM_COLORS goes through atoi().
bool match(const char *v, const char *name) { return strcasecmp(v, name) == 0; }

s2tc_b200_settings settings_from_env()
{
	s2tc_b200_settings s;
	s.dxt = S2TC_B200_DXT1;
	s.cd = S2TC_B200_WAVG;
	s.nrandom = -1;
	s.refine = S2TC_B200_REFINE_ALWAYS;
	s.dither = S2TC_B200_DITHER_SIMPLE;
	if (const char *v = getenv("S2TC_DITHER_MODE")) {
		static const char *names[] = {"NONE", "SIMPLE", "FLOYDSTEINBERG"};
		int hit = -1;
		for (int i = 0; i < 3; ++i)
			if (match(v, names[i]))
				hit = i;
		if (hit >= 0)
			s.dither = hit;
		else
			fprintf(stderr, "Invalid dither mode: %s\n", v);
	}
	if (const char *v = getenv("S2TC_COLORDIST_MODE")) {
		static const char *names[] = {"RGB", "YUV", "SRGB", "SRGB_MIXED", "AVG", "WAVG", "W0AVG", "NORMALMAP"};
		int hit = -1;
		for (int i = 0; i < 8; ++i)
			if (match(v, names[i]))
				hit = i;
		if (hit >= 0)
			s.cd = hit;
		else
			fprintf(stderr, "Invalid color dist mode: %s\n", v);
	}
	if (const char *v = getenv("S2TC_RANDOM_COLORS"))
		s.nrandom = atoi(v);
	if (const char *v = getenv("S2TC_REFINE_COLORS")) {
		static const char *names[] = {"NEVER", "ALWAYS", "LOOP"};
		int hit = -1;
		for (int i = 0; i < 3; ++i)
			if (match(v, names[i]))
				hit = i;
		if (hit >= 0)
			s.refine = hit;
		else
			fprintf(stderr, "Invalid refinement mode: %s\n", v);
	}
	return s;
}

s2tc_b200_ctx *ctx_or_complain(const char *who)
{
	s2tc_b200_ctx *c = s2tc_b200_default_ctx();
	if (!c)
		fprintf(stderr, "%s: no usable CUDA device: %s\n", who, s2tc_b200_last_error());
	return c;
}

// ---- s2tc_encode_block_func: one static trampoline per (dxt, cd, fast?, refine) --------------------
// The reference returns a bare function pointer with no closure (ref s2tc_algorithm.cpp:1110-1194), so
// each combination needs its own function.  nrandom is an argument of the returned function; like the
// reference, the compression mode (fast vs normal) is fixed when the pointer is handed out.
template <int DXT, int CD, bool FAST, int REFINE>
void block_trampoline(unsigned char *out, const unsigned char *rgba, int iw, int w, int h, int nrandom)
{
	s2tc_b200_ctx *c = ctx_or_complain("s2tc_encode_block");
	if (!c)
		return;
	s2tc_b200_settings s;
	s.dxt = DXT;
	s.cd = CD;
	// a pointer obtained for normal mode keeps searching even if later called with nrandom < 0
	// (ref :936-1007 only tests nrandom > 0 inside); fast mode ignores nrandom altogether
	s.nrandom = FAST ? -1 : (nrandom < 0 ? 0 : nrandom);
	s.refine = REFINE;
	s.dither = S2TC_B200_DITHER_NONE;
	uint64_t cursor = s2tc_b200_rand_cursor_get();
	if (s2tc_b200_encode_block_host(c, &s, out, rgba, iw, w, h, &cursor) != 0) {
		fprintf(stderr, "s2tc_encode_block: %s\n", s2tc_b200_last_error());
		return;
	}
	s2tc_b200_rand_cursor_set(cursor);
}

template <int DXT, int CD, bool FAST>
s2tc_encode_block_func_t pick_refine(int refine)
{
	switch (refine) {
	case REFINE_NEVER: return block_trampoline<DXT, CD, FAST, REFINE_NEVER>;
	case REFINE_LOOP: return block_trampoline<DXT, CD, FAST, REFINE_LOOP>;
	default: return block_trampoline<DXT, CD, FAST, REFINE_ALWAYS>; // ref :1120
	}
}

template <int DXT, int CD>
s2tc_encode_block_func_t pick_mode(int nrandom, int refine)
{
	if (CD == NORMALMAP || nrandom >= 0) // ref :1139
		return pick_refine<DXT, CD, false>(refine);
	return pick_refine<DXT, CD, true>(refine);
}

template <int CD>
s2tc_encode_block_func_t pick_dxt(int dxt, int nrandom, int refine)
{
	switch (dxt) {
	case DXT1: return pick_mode<DXT1, CD>(nrandom, refine);
	case DXT3: return pick_mode<DXT3, CD>(nrandom, refine);
	default: return pick_mode<DXT5, CD>(nrandom, refine); // ref :1156
	}
}

// decode helpers --------------------------------------------------------------------------------
inline unsigned expand5(unsigned v) { return (v << 3) | (v >> 2); }
inline unsigned expand6(unsigned v) { return (v << 2) | (v >> 4); }

// colour of texel (i,j) from an 8-byte colour block; S2TC shows the would-be interpolated codes as a
// checkerboard of the two endpoints (ref s2tc_libtxc_dxtn.cpp:44-50)
void decode_color(const unsigned char *cb, int i, int j, bool dxt1, unsigned char *t, bool *transparent)
{
	unsigned c0 = cb[0] | (cb[1] << 8), c1 = cb[2] | (cb[3] << 8);
	const unsigned code = (cb[4 + (j & 3)] >> (2 * (i & 3))) & 3u;
	unsigned c = c0;
	*transparent = false;
	if (code == 1) {
		c = c1;
	} else if (code == 3 && dxt1 && c1 >= c0) {
		c = 0;
		*transparent = true;
	} else if (code >= 2) {
		if ((i ^ j) & 1)
			c = c1;
	}
	t[0] = (unsigned char) expand5((c >> 11) & 0x1F);
	t[1] = (unsigned char) expand6((c >> 5) & 0x3F);
	t[2] = (unsigned char) expand5(c & 0x1F);
}

inline const unsigned char *block_at(int row_stride_texels, const unsigned char *pix, int i, int j, int bytes)
{
	return pix + (size_t) (((row_stride_texels + 3) >> 2) * (j >> 2) + (i >> 2)) * bytes;
}

} // namespace

extern "C" {

// the reference's environment parsing as a callable (used by tools that drive the batched API directly)
void s2tc_b200_settings_from_env(int dxt, s2tc_b200_settings *out)
{
	if (!out)
		return;
	*out = settings_from_env();
	out->dxt = dxt;
}

void tx_compress_dxtn(int srccomps, int width, int height, const unsigned char *srcPixData, unsigned int destformat,
		unsigned char *dest, int dstRowStride)
{
	s2tc_b200_settings s = settings_from_env();
	switch (destformat) { // ref :218-236
	case S2TC_B200_GL_RGB_DXT1:
	case S2TC_B200_GL_RGBA_DXT1: s.dxt = S2TC_B200_DXT1; break;
	case S2TC_B200_GL_RGBA_DXT3: s.dxt = S2TC_B200_DXT3; break;
	case S2TC_B200_GL_RGBA_DXT5: s.dxt = S2TC_B200_DXT5; break;
	default:
		fprintf(stderr, "libdxtn: Bad dstFormat %d in tx_compress_dxtn\n", destformat);
		return;
	}
	s2tc_b200_ctx *c = ctx_or_complain("tx_compress_dxtn");
	if (!c)
		return;
	uint64_t cursor = s2tc_b200_rand_cursor_get();
	if (s2tc_b200_compress_host(c, &s, srccomps, width, height, srcPixData, dest, dstRowStride, &cursor) != 0) {
		fprintf(stderr, "tx_compress_dxtn: %s\n", s2tc_b200_last_error());
		return;
	}
	s2tc_b200_rand_cursor_set(cursor);
}

void rgb565_image(unsigned char *out, const unsigned char *rgba, int w, int h, int srccomps, int alphabits,
		enum DitherMode dither)
{
	s2tc_b200_ctx *c = ctx_or_complain("rgb565_image");
	if (!c)
		return;
	if (s2tc_b200_rgb565_host(c, out, rgba, w, h, srccomps, alphabits, (int) dither) != 0)
		fprintf(stderr, "rgb565_image: %s\n", s2tc_b200_last_error());
}

s2tc_encode_block_func_t s2tc_encode_block_func(enum DxtMode dxt, ColorDistMode cd, int nrandom, enum RefinementMode refine)
{
	switch (cd) {
	case RGB: return pick_dxt<RGB>(dxt, nrandom, refine);
	case YUV: return pick_dxt<YUV>(dxt, nrandom, refine);
	case SRGB: return pick_dxt<SRGB>(dxt, nrandom, refine);
	case SRGB_MIXED: return pick_dxt<SRGB_MIXED>(dxt, nrandom, refine);
	case AVG: return pick_dxt<AVG>(dxt, nrandom, refine);
	case W0AVG: return pick_dxt<W0AVG>(dxt, nrandom, refine);
	case NORMALMAP: return pick_dxt<NORMALMAP>(dxt, nrandom, refine);
	default: return pick_dxt<WAVG>(dxt, nrandom, refine); // ref :1183
	}
}

s2tc_encode_block_func_t get_s2tc_encoder(enum DxtMode dxt, ColorDistMode cd, int nrandom, enum RefinementMode refine)
{
	return s2tc_encode_block_func(dxt, cd, nrandom, refine);
}

void fetch_2d_texel_rgb_dxt1(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel)
{
	unsigned char *t = (unsigned char *) texel;
	bool tr;
	decode_color(block_at(srcRowStride, pixdata, i, j, 8), i, j, true, t, &tr);
	t[3] = 255;
}

void fetch_2d_texel_rgba_dxt1(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel)
{
	unsigned char *t = (unsigned char *) texel;
	bool tr;
	decode_color(block_at(srcRowStride, pixdata, i, j, 8), i, j, true, t, &tr);
	t[3] = tr ? 0 : 255;
}

void fetch_2d_texel_rgba_dxt3(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel)
{
	unsigned char *t = (unsigned char *) texel;
	const unsigned char *blk = block_at(srcRowStride, pixdata, i, j, 16);
	bool tr;
	decode_color(blk + 8, i, j, false, t, &tr);
	const unsigned a = (blk[(j & 3) * 2 + ((i & 3) >> 1)] >> (4 * (i & 1))) & 0x0F;
	t[3] = (unsigned char) (a | (a << 4));
}

void fetch_2d_texel_rgba_dxt5(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel)
{
	unsigned char *t = (unsigned char *) texel;
	const unsigned char *blk = block_at(srcRowStride, pixdata, i, j, 16);
	bool tr;
	decode_color(blk + 8, i, j, false, t, &tr);
	unsigned a0 = blk[0], a1 = blk[1];
	uint64_t bits = 0;
	for (int k = 0; k < 6; ++k)
		bits |= (uint64_t) blk[2 + k] << (8 * k);
	const unsigned code = (unsigned) ((bits >> (3 * ((j & 3) * 4 + (i & 3)))) & 7u);
	unsigned a = a0;
	if (code == 1)
		a = a1;
	else if (code == 6 && a1 >= a0)
		a = 0;
	else if (code == 7 && a1 >= a0)
		a = 255;
	else if (code >= 2) {
		if ((i ^ j) & 1)
			a = a1;
	}
	t[3] = (unsigned char) a;
}

} // extern "C"
