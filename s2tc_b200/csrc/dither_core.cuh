// dither_core.cuh -- the 565 pre-pass arithmetic (reference rgb565_image, s2tc_algorithm.cpp:1196-1465)
// in a form that can be evaluated in parallel, host+device.
//
// DITHER_SIMPLE is a serial recurrence: each channel carries its quantisation error to the next
// texel in raster order and the carry is never reset, not even at row ends (ref :1310-1347).  The
// carry, however, lives in a tiny set (shift 3: [-7,7], shift 2: [-3,3], shift 4: [-15,15]), so a
// run of texels is summarised exactly by a transfer map "carry in -> carry out" with at most 31
// entries, and maps compose associatively.  The 1-bit alpha variant (diffuse1, ref :1208-1216)
// has the closed form carry = balanced residue of (carry_in + sum of sources) mod 255.
// Three phases: per-chunk maps, a scan over the maps, a replay of each chunk from its true carry.
//
// Maps are byte tables (entry k = state index carry + radius).  For the 5- and 6-bit channels the map
// of a run is built right to left, A_t = A_{t+1} o M(v_t): the accumulated table A is the byte table
// of a PRMT (byte permute) and the per-texel map M(v) comes from a 256-entry lookup table as PRMT
// selectors, so one texel-channel costs 12 byte permutes for all 15 start states together instead of
// 15 explicit trajectories (profiles/r01a: 1.2 ms -> see DESIGN.md).
#pragma once

#include "s2tc_defs.h"

namespace s2tc {

// one step of diffuse() (ref :1198-1207); SHIFT in {2,3,4}
template <int SHIFT>
S2TC_HD int diffuse_step(int &carry, int src)
{
	constexpr int top = (1 << (8 - SHIFT)) - 1;
	const int s = src + carry;
	int q = s >> SHIFT;
	q = q < 0 ? 0 : (q > top ? top : q);
	const int back = (q << SHIFT) | (q >> (8 - 2 * SHIFT));
	carry = s - back;
	return q;
}

// one step of diffuse1() (ref :1208-1216)
S2TC_HD int diffuse1_step(int &carry, int src)
{
	const int s = src + carry;
	const int q = s >= 128;
	carry = s - (q ? 255 : 0);
	return q;
}

// balanced residue in [-127,127] of v mod 255 (v >= -127)
S2TC_HD int balanced255(int v)
{
	int r = (v + 127) % 255;
	return r - 127;
}

// How one channel of the pre-pass behaves.
enum ChanKind : int {
	kChanShift3 = 0, // r, b
	kChanShift2 = 1, // g
	kChanShift4 = 2, // DXT3 alpha
	kChanBit1 = 3,   // DXT1 alpha (diffuse1)
	kChanCopy = 4    // DXT5 alpha copy, or constant all-ones for 3-component sources: no carry
};

S2TC_HD int chan_radius(int kind) { return kind == kChanShift3 ? 7 : (kind == kChanShift2 ? 3 : (kind == kChanShift4 ? 15 : 0)); }
S2TC_HD int chan_states(int kind) { return kind <= kChanShift4 ? 2 * chan_radius(kind) + 1 : 0; }

S2TC_HD int alpha_chan_kind(int srccomps, int alphabits)
{
	if (srccomps == 3 || alphabits == 8)
		return kChanCopy;
	return alphabits == 1 ? kChanBit1 : kChanShift4;
}

struct ChanKinds { int k[4]; };

S2TC_HD ChanKinds chan_kinds(int srccomps, int alphabits)
{
	ChanKinds c;
	c.k[0] = kChanShift3;
	c.k[1] = kChanShift2;
	c.k[2] = kChanShift3;
	c.k[3] = alpha_chan_kind(srccomps, alphabits);
	return c;
}

// Transfer map of a run of texels for one channel: e[k] = state index out for state index k in.
// For kChanBit1 e[0] holds (sum of sources) mod 255 instead.  32 bytes so that four of them are the
// 128-byte "summary" of a texel range that shards exchange.
struct alignas(16) ByteMap {
	uint8_t e[32];
};

S2TC_HD void bmap_identity(ByteMap &m, int kind)
{
	for (int k = 0; k < 32; ++k)
		m.e[k] = kind <= kChanShift4 ? (uint8_t) k : 0;
}

// carry out for a given carry in
S2TC_HD int bmap_apply(const ByteMap &m, int kind, int carry)
{
	if (kind == kChanCopy)
		return 0;
	if (kind == kChanBit1)
		return balanced255(carry + (int) m.e[0]);
	const int r = chan_radius(kind);
	return (int) m.e[carry + r] - r;
}

// out = second after first (out may alias neither)
S2TC_HD void bmap_compose(ByteMap &out, const ByteMap &first, const ByteMap &second, int kind)
{
	if (kind == kChanBit1) {
		out.e[0] = (uint8_t) (((unsigned) first.e[0] + second.e[0]) % 255u);
	} else if (kind <= kChanShift4) {
		const int ns = chan_states(kind);
		for (int k = 0; k < ns; ++k)
			out.e[k] = second.e[first.e[k]];
	}
}

// ---- byte permute -----------------------------------------------------------------------------------
// out byte i = byte (s >> 4i) & 7 of the 8-byte table {y:x}.  Selectors here never set bit 3 of a nibble, and only
// the low 16 bits of s count, which is exactly PRMT's contract: raw prmt.b32 instead of __byte_perm, which ANDs every
// selector with 0x7777 first (17 LOP3 per texel in the maps kernel, profiles/r01h).
S2TC_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
#if defined(__CUDA_ARCH__)
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(s));
	return d;
#else
	const uint64_t t = ((uint64_t) y << 32) | x;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i)
		r |= (uint32_t) ((t >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
	return r;
#endif
}

// ---- per-texel maps as PRMT selectors ---------------------------------------------------------------
// lut3[v]: map of one texel with source value v for a shift-3 channel (15 states, table of 16 bytes in
//          4 registers).  Word g serves output bytes 4g..4g+3: its low half selects, for each of them,
//          byte (target & 7) out of the lower AND out of the upper 8 table bytes; its high half is a second
//          selector that picks, byte by byte, the lower (0..3) or the upper (4..7) candidate.
// lut2[v]: shift-2 channel (7 states, table of 8 bytes): 8 selector nibbles
struct DitherLut {
	uint32_t lut3[256][4];
	uint32_t lut2[256];
};

inline void build_dither_lut(DitherLut &L)
{
	for (int v = 0; v < 256; ++v) {
		for (int g = 0; g < 4; ++g) {
			uint32_t sel = 0, pick = 0;
			for (int j = 0; j < 4; ++j) {
				const int k = 4 * g + j;
				int target = 0;
				if (k < 15) {
					int c = k - 7;
					(void) diffuse_step<3>(c, v);
					target = c + 7;
				}
				sel |= (uint32_t) (target & 7) << (4 * j);
				pick |= (uint32_t) (j + (target >= 8 ? 4 : 0)) << (4 * j);
			}
			L.lut3[v][g] = sel | (pick << 16);
		}
		uint32_t sel = 0;
		for (int k = 0; k < 7; ++k) {
			int c = k - 3;
			(void) diffuse_step<2>(c, v);
			sel |= (uint32_t) (c + 3) << (4 * k);
		}
		L.lut2[v] = sel;
	}
}

// accumulated tables of the three colour channels of a run (built right to left)
struct RgbTables {
	uint32_t r[4], g[2], b[4];
};

S2TC_HD void rgb_tables_init(RgbTables &t)
{
	t.r[0] = t.b[0] = t.g[0] = 0x03020100u;
	t.r[1] = t.b[1] = t.g[1] = 0x07060504u;
	t.r[2] = t.b[2] = 0x0B0A0908u;
	t.r[3] = t.b[3] = 0x0F0E0D0Cu;
}

// x >> 16 as a multiply-high by k16 = 65536.  The kernels pass k16 as an argument the compiler cannot fold, so the
// nine selector shifts per texel run on the FMA pipe (IMAD.HI) instead of the ALU pipe, which bounds the maps kernel.
S2TC_HD uint32_t upper_half(uint32_t x, uint32_t k16)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(x, k16);
#else
	return (uint32_t) (((uint64_t) x * k16) >> 32);
#endif
}

// A <- A o M(v) for a 16-byte table; L = the four words of lut3[v]
S2TC_HD void table16_step(uint32_t (&T)[4], uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t k16)
{
	const uint32_t L[4] = {l0, l1, l2, l3};
	uint32_t n[4];
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		const uint32_t lo = byte_perm(T[0], T[1], L[g]);
		const uint32_t hi = byte_perm(T[2], T[3], L[g]);
		n[g] = byte_perm(lo, hi, upper_half(L[g], k16));
	}
#pragma unroll
	for (int g = 0; g < 4; ++g)
		T[g] = n[g];
}

// prepend one texel (w = raw r | g << 8 | b << 16 | ...) to the run summarised by t
S2TC_HD void rgb_tables_prepend(RgbTables &t, uint32_t w, const uint32_t (*lut3)[4], const uint32_t *lut2, uint32_t k16 = 65536u)
{
	const uint32_t *lr = lut3[w & 0xFFu], *lb = lut3[(w >> 16) & 0xFFu];
	table16_step(t.r, lr[0], lr[1], lr[2], lr[3], k16);
	table16_step(t.b, lb[0], lb[1], lb[2], lb[3], k16);
	const uint32_t s = lut2[(w >> 8) & 0xFFu];
	const uint32_t g0 = byte_perm(t.g[0], t.g[1], s), g1 = byte_perm(t.g[0], t.g[1], upper_half(s, k16));
	t.g[0] = g0;
	t.g[1] = g1;
}

S2TC_HD void rgb_tables_store(const RgbTables &t, ByteMap &r, ByteMap &g, ByteMap &b)
{
	for (int k = 0; k < 32; ++k)
		r.e[k] = g.e[k] = b.e[k] = 0;
	for (int k = 0; k < 16; ++k) {
		r.e[k] = (uint8_t) (t.r[k >> 2] >> (8 * (k & 3)));
		b.e[k] = (uint8_t) (t.b[k >> 2] >> (8 * (k & 3)));
	}
	for (int k = 0; k < 8; ++k)
		g.e[k] = (uint8_t) (t.g[k >> 2] >> (8 * (k & 3)));
}

// first = tables of the left half of a run, second = tables of the right half: the run's tables.
// total[k] = second[first[k]]: `second` is the byte table, `first` is turned into selectors on the fly.
S2TC_HD void rgb_tables_join(RgbTables &out, const RgbTables &first, const RgbTables &second)
{
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		// bytes of first.x[g] are targets 0..14: low 3 bits select, bit 3 chooses the half
		const uint32_t fr = first.r[g], fb = first.b[g];
		const uint32_t selr = (fr & 7u) | ((fr >> 4) & 0x70u) | ((fr >> 8) & 0x700u) | ((fr >> 12) & 0x7000u);
		const uint32_t selb = (fb & 7u) | ((fb >> 4) & 0x70u) | ((fb >> 8) & 0x700u) | ((fb >> 12) & 0x7000u);
		const uint32_t pickr = 0x3210u + (((fr >> 1) & 4u) | ((fr >> 5) & 0x40u) | ((fr >> 9) & 0x400u) | ((fr >> 13) & 0x4000u));
		const uint32_t pickb = 0x3210u + (((fb >> 1) & 4u) | ((fb >> 5) & 0x40u) | ((fb >> 9) & 0x400u) | ((fb >> 13) & 0x4000u));
		out.r[g] = byte_perm(byte_perm(second.r[0], second.r[1], selr), byte_perm(second.r[2], second.r[3], selr), pickr);
		out.b[g] = byte_perm(byte_perm(second.b[0], second.b[1], selb), byte_perm(second.b[2], second.b[3], selb), pickb);
	}
#pragma unroll
	for (int g = 0; g < 2; ++g) {
		const uint32_t fg = first.g[g];
		const uint32_t sel = (fg & 7u) | ((fg >> 4) & 0x70u) | ((fg >> 8) & 0x700u) | ((fg >> 12) & 0x7000u);
		out.g[g] = byte_perm(second.g[0], second.g[1], sel);
	}
}

// map of the alpha channel of a run (forward, explicit trajectories: only DXT3 pays for the 31 states)
S2TC_HD void alpha_map_of_run(ByteMap &m, int kind, const uint8_t *src, int stride, int count)
{
	for (int k = 0; k < 32; ++k)
		m.e[k] = 0;
	if (kind == kChanShift4) {
		int st[31];
#pragma unroll
		for (int k = 0; k < 31; ++k)
			st[k] = k - 15;
		for (int i = 0; i < count; ++i) {
			const int v = src[(size_t) i * stride];
#pragma unroll
			for (int k = 0; k < 31; ++k)
				(void) diffuse_step<4>(st[k], v);
		}
#pragma unroll
		for (int k = 0; k < 31; ++k)
			m.e[k] = (uint8_t) (st[k] + 15);
	} else if (kind == kChanBit1) {
		uint32_t sum = 0;
		for (int i = 0; i < count; ++i)
			sum += src[(size_t) i * stride];
		m.e[0] = (uint8_t) (sum % 255u);
	}
}

// Replays one texel from the true carries (ref :1314-1347): raw word in, reduced word out
S2TC_HD uint32_t replay_texel(int (&carry)[4], uint32_t w, int alpha_kind, bool has_alpha, int alphabits)
{
	const uint32_t r = (uint32_t) diffuse_step<3>(carry[0], (int) (w & 0xFFu));
	const uint32_t g = (uint32_t) diffuse_step<2>(carry[1], (int) ((w >> 8) & 0xFFu));
	const uint32_t b = (uint32_t) diffuse_step<3>(carry[2], (int) ((w >> 16) & 0xFFu));
	uint32_t a;
	if (alpha_kind == kChanShift4)
		a = (uint32_t) diffuse_step<4>(carry[3], (int) (w >> 24));
	else if (alpha_kind == kChanBit1)
		a = (uint32_t) diffuse1_step(carry[3], (int) (w >> 24));
	else
		a = has_alpha ? (w >> 24) : ((1u << alphabits) - 1u);
	return r | (g << 8) | (b << 16) | (a << 24);
}

// ---- DITHER_FLOYDSTEINBERG (ref :1218-1261, :1350-1412) -----------------------------------------------------
// One texel of floyd() / floyd1(): `incoming` is the error that reached this texel (thisrow[x+1] in the reference),
// the four parts go right (e7), down-left (e3), down (e5) and down-right (e1).  12-bit domain; C division
// truncates toward zero and so does CUDA's.  SHIFT 7 selects the 1-bit variant floyd1().
struct FloydOut {
	int q, e7, e3, e5, e1;
};

template <int SHIFT>
S2TC_HD FloydOut floyd_texel(int src, int incoming)
{
	const int s = ((src << 4) | (src >> 4)) + incoming;
	int q, back;
	if (SHIFT == 7) {
		q = s >= 2048;
		back = q ? 4095 : 0;
	} else {
		constexpr int top = (1 << (8 - (SHIFT == 7 ? 1 : SHIFT))) - 1;
		q = s >> (SHIFT + 4);
		q = q < 0 ? 0 : (q > top ? top : q);
		back = q * 4095 / top;
	}
	int err = s - back;
	FloydOut o;
	o.q = q;
	o.e7 = (err * 7 + 8) / 16;
	err -= o.e7;
	o.e3 = (err * 3 + 4) / 9;
	err -= o.e3;
	o.e5 = (err * 5 + 3) / 6;
	err -= o.e5;
	o.e1 = err;
	return o;
}

// DITHER_NONE on one texel (ref :1269-1306): raw bytes -> reduced texel word
S2TC_HD uint32_t reduce_none(uint32_t r, uint32_t g, uint32_t b, uint32_t a, int alphabits, bool has_alpha)
{
	const uint32_t ra = has_alpha ? (a >> (8 - alphabits)) : ((1u << alphabits) - 1u);
	return (r >> 3) | ((g >> 2) << 8) | ((b >> 3) << 16) | (ra << 24);
}

} // namespace s2tc
