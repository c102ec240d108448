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

S2TC_HD int alpha_chan_kind(int srccomps, int alphabits)
{
	if (srccomps == 3 || alphabits == 8)
		return kChanCopy;
	return alphabits == 1 ? kChanBit1 : kChanShift4;
}

// Transfer map of a run of texels for one channel.  Entry k (carry = k - radius) is stored in 5 bits.
// For kChanBit1 w[0] holds (sum of sources) mod 255 instead.
struct CarryMap {
	uint64_t w[3];
};

S2TC_HD int map_get(const CarryMap &m, int k) { return (int) ((m.w[k / 12] >> (5 * (k % 12))) & 31u); }
S2TC_HD void map_set(CarryMap &m, int k, int v) { m.w[k / 12] |= (uint64_t) v << (5 * (k % 12)); }
S2TC_HD void map_clear(CarryMap &m) { m.w[0] = m.w[1] = m.w[2] = 0; }

S2TC_HD void map_identity(CarryMap &m, int kind)
{
	map_clear(m);
	const int ns = 2 * chan_radius(kind) + 1;
	if (kind <= kChanShift4)
		for (int k = 0; k < ns; ++k)
			map_set(m, k, k);
}

// carry out for a given carry in
S2TC_HD int map_apply(const CarryMap &m, int kind, int carry)
{
	if (kind == kChanCopy)
		return 0;
	if (kind == kChanBit1)
		return balanced255(carry + (int) m.w[0]);
	const int r = chan_radius(kind);
	return map_get(m, carry + r) - r;
}

// second after first
S2TC_HD void map_compose(CarryMap &out, const CarryMap &first, const CarryMap &second, int kind)
{
	CarryMap t;
	map_clear(t);
	if (kind == kChanBit1) {
		t.w[0] = (first.w[0] + second.w[0]) % 255u;
	} else if (kind != kChanCopy) {
		const int ns = 2 * chan_radius(kind) + 1;
		for (int k = 0; k < ns; ++k)
			map_set(t, k, map_get(second, map_get(first, k)));
	}
	out = t;
}

// Transfer map of `count` source bytes src[0], src[stride], ... for one channel.
template <int SHIFT>
S2TC_HD void map_of_run_shift(CarryMap &m, const uint8_t *src, int stride, int count)
{
	constexpr int R = (1 << SHIFT) - 1, NS = 2 * R + 1;
	int st[NS];
#pragma unroll
	for (int k = 0; k < NS; ++k)
		st[k] = k - R;
	for (int i = 0; i < count; ++i) {
		const int v = src[(size_t) i * stride];
#pragma unroll
		for (int k = 0; k < NS; ++k)
			(void) diffuse_step<SHIFT>(st[k], v);
	}
	map_clear(m);
#pragma unroll
	for (int k = 0; k < NS; ++k)
		map_set(m, k, st[k] + R);
}

S2TC_HD void map_of_run(CarryMap &m, int kind, const uint8_t *src, int stride, int count)
{
	switch (kind) {
	case kChanShift3: map_of_run_shift<3>(m, src, stride, count); break;
	case kChanShift2: map_of_run_shift<2>(m, src, stride, count); break;
	case kChanShift4: map_of_run_shift<4>(m, src, stride, count); break;
	case kChanBit1: {
		uint32_t sum = 0;
		for (int i = 0; i < count; ++i)
			sum += src[(size_t) i * stride];
		map_clear(m);
		m.w[0] = sum % 255u;
		break;
	}
	default: map_clear(m); break;
	}
}

// Replays a run from its true carry; writes quantised values to dst[0], dst[4], ...; returns carry out.
S2TC_HD int replay_run(int kind, int carry, const uint8_t *src, int stride, int count, uint8_t *dst)
{
	switch (kind) {
	case kChanShift3:
		for (int i = 0; i < count; ++i)
			dst[(size_t) i * 4] = (uint8_t) diffuse_step<3>(carry, src[(size_t) i * stride]);
		break;
	case kChanShift2:
		for (int i = 0; i < count; ++i)
			dst[(size_t) i * 4] = (uint8_t) diffuse_step<2>(carry, src[(size_t) i * stride]);
		break;
	case kChanShift4:
		for (int i = 0; i < count; ++i)
			dst[(size_t) i * 4] = (uint8_t) diffuse_step<4>(carry, src[(size_t) i * stride]);
		break;
	case kChanBit1:
		for (int i = 0; i < count; ++i)
			dst[(size_t) i * 4] = (uint8_t) diffuse1_step(carry, src[(size_t) i * stride]);
		break;
	default: break;
	}
	return carry;
}

// DITHER_NONE on one texel (ref :1269-1306): raw bytes -> reduced texel word
S2TC_HD uint32_t reduce_none(uint32_t r, uint32_t g, uint32_t b, uint32_t a, int alphabits, bool has_alpha)
{
	const uint32_t ra = has_alpha ? (a >> (8 - alphabits)) : ((1u << alphabits) - 1u);
	return (r >> 3) | ((g >> 2) << 8) | ((b >> 3) << 16) | (ra << 24);
}

} // namespace s2tc
