// block_core.cuh -- per-4x4-block S2TC logic in "one thread owns one block" form, host+device.
//
// Everything here is sequential per block and works on a block held in registers:
//   * fast_candidates   : MODE_FAST endpoint pick            (ref s2tc_algorithm.cpp:879-935)
//   * gather_colors     : MODE_NORMAL colour gathering       (ref :938-959, :997-1001)
//   * candidate_box / draw_candidate : nrandom>0 candidates (ref :962-993)
//   * search_*_scalar   : the c0/c1 pair search, one thread  (ref :367-478)   [generic fallback]
//   * finish_block      : equal-endpoint fix-ups, NEVER/ALWAYS/LOOP refinement of colour and DXT5
//                         alpha, DXT3 alpha nibbles and packing (ref :582-870, :1010-1107)
//
// The GPU kernels call these from device code (kernels_*.cu); tests/hostsim compiles the very
// same header for the CPU to check it against the oracle without a GPU.  The cooperative
// (thread-group) pair search lives in kernels_search.cu.
//
// Index bookkeeping: instead of the reference's bit-array objects we keep one 16-bit mask per
// index value (bit i = texel y*4+x).  The transformations the reference applies to whole index
// arrays (clobber on equal endpoints, flip on endpoint swap) become mask algebra; missing texels
// of partial blocks take part in the flip exactly as in the reference (ref :807-813 loops over all
// 16 positions).
#pragma once

#include "colordist.cuh"

namespace s2tc {

struct Block {
	uint32_t px[16]; // reduced texels {r5,g6,b5,a}, index i = y*4 + x
	uint32_t valid;  // bit i set iff x < w && y < h
};

S2TC_HD uint32_t valid_mask(int w, int h)
{
	uint32_t row = (1u << w) - 1u; // w in 1..4
	uint32_t m = 0;
	for (int y = 0; y < h; ++y)
		m |= row << (4 * y);
	return m;
}

// ---- 565 colour algebra on packed texels (ref :50-138) -----------------------------------------
S2TC_HD uint32_t to565(uint32_t p) { return ((p & 0x1F) << 11) | (((p >> 8) & 0x3F) << 5) | ((p >> 16) & 0x1F); }
S2TC_HD uint32_t from565(uint32_t v) { return px_make((int) (v >> 11) & 31, (int) (v >> 5) & 63, (int) v & 31); }
S2TC_HD bool col_lt(uint32_t a, uint32_t b) { return to565(a) < to565(b); } // lexicographic r,g,b == numeric 565
// operator++ / operator-- are +-1 on the 16-bit 565 value with wrap-around (ref :82-131)
S2TC_HD uint32_t col_bump(uint32_t p)
{
	uint32_t v = to565(p);
	return from565(v == 0xFFFFu ? v - 1u : v + 1u); // "-- if max else ++" (ref :1012-1015)
}
S2TC_HD int alpha_bump(int a) { return a == 255 ? 254 : a + 1; } // ref :1022-1025

// spread the low 16 bits of m so that bit i lands at bit 2i
S2TC_HD uint32_t spread2(uint32_t m)
{
	m = (m | (m << 8)) & 0x00FF00FFu;
	m = (m | (m << 4)) & 0x0F0F0F0Fu;
	m = (m | (m << 2)) & 0x33333333u;
	m = (m | (m << 1)) & 0x55555555u;
	return m;
}
// bit i -> bit 3i (48-bit result)
S2TC_HD uint64_t spread3(uint32_t m)
{
	uint64_t x = m & 0xFFFFu;
	x = (x | (x << 16)) & 0x0000FF0000FFull;  // bytes 8 apart -> 24 apart
	x = (x | (x << 8)) & 0x00F00F00F00Full;   // nibbles 12 apart
	x = (x | (x << 4)) & 0x0C30C30C30C3ull;   // pairs 6 apart
	x = (x | (x << 2)) & 0x249249249249ull;   // singles 3 apart
	return x;
}
S2TC_HD int popc16(uint32_t m)
{
#if defined(__CUDA_ARCH__)
	return __popc(m);
#else
	return __builtin_popcount(m);
#endif
}

// floor(x / (2n)) for n in 1..16 and 0 <= x < 8192, by multiplication with ceil(2^18 / (2n))
// (exact: the excess e = M*2n - 2^18 is < 2n <= 32 and x*e < 2^18; also checked exhaustively in
// tests/test_hostsim.py through the hostsim build)
#if defined(__CUDA_ARCH__)
static __device__ __constant__ uint32_t kHalfRecip18[17] = {0, 131072, 65536, 43691, 32768, 26215, 21846, 18725, 16384, 14564, 13108, 11916, 10923, 10083, 9363, 8739, 8192};
#else
static const uint32_t kHalfRecip18[17] = {0, 131072, 65536, 43691, 32768, 26215, 21846, 18725, 16384, 14564, 13108, 11916, 10923, 10083, 9363, 8739, 8192};
#endif
S2TC_HD uint32_t half_recip18(int n) { return kHalfRecip18[n]; }
S2TC_HD int div_by_2n(int x, uint32_t recip) { return (int) (((uint32_t) x * recip) >> 18); }

// ---- MODE_FAST endpoint pick (ref :879-935) ----------------------------------------------------
template <int DXT, int CD>
S2TC_HD void fast_candidates(const Block &b, uint32_t &c0, uint32_t &c1, int &a0, int &a1)
{
	typedef Metric<CD> M;
	const typename M::Feat black = M::feat(0);
	int dmin = 0x7FFFFFFF, dmax = 0;
	c0 = px_make(31, 63, 31);
	c1 = 0;
	a0 = a1 = px_a(b.px[0]);
#pragma unroll
	for (int x = 0; x < 4; ++x)
#pragma unroll
		for (int y = 0; y < 4; ++y) { // column-major: decides first-wins ties
			const int i = y * 4 + x;
			const uint32_t p = b.px[i];
			bool use = (b.valid >> i) & 1u;
			if (DXT == kDxt1)
				use = use && px_a(p) != 0;
			if (use) {
				int d = M::dist(M::feat(p), black);
				if (d > dmax) {
					dmax = d;
					c1 = px_rgb(p);
				}
				if (d < dmin) {
					dmin = d;
					c0 = px_rgb(p);
				}
				if (DXT == kDxt5) {
					int a = px_a(p);
					if (a != 255) {
						if (a > a1)
							a1 = a;
						if (a < a0)
							a0 = a;
					}
				}
			}
		}
}

// ---- MODE_NORMAL gathering (ref :938-959) -----------------------------------------------------
// c[] receives packed rgb, ca[] alpha; returns n (>= 1).  DXT1 skips texels with alpha 0.
template <int DXT>
S2TC_HD int gather_colors(const Block &b, uint32_t *c, uint8_t *ca)
{
	int n = 0;
#pragma unroll
	for (int x = 0; x < 4; ++x)
#pragma unroll
		for (int y = 0; y < 4; ++y) {
			const int i = y * 4 + x;
			const uint32_t p = b.px[i];
			bool use = (b.valid >> i) & 1u;
			if (DXT == kDxt1)
				use = use && px_a(p) != 0;
			if (use) {
				c[n] = px_rgb(p);
				ca[n] = (uint8_t) px_a(p);
				++n;
			}
		}
	if (n == 0) {
		c[0] = 0;
		ca[0] = 0;
		n = 1;
	}
	return n;
}

// ---- nrandom > 0 candidates (ref :962-993) ------------------------------------------------------
// Bounding box of the gathered colours; random candidates are lo + rand() % len per channel.
struct CandBox {
	int lo[3], len[3];
	int alo, alen;
};

S2TC_HD CandBox candidate_box(const uint32_t *c, const uint8_t *ca, int n)
{
	int lo[3] = {31, 63, 31}, hi[3] = {0, 0, 0};
	int amin = 255, amax = 0;
	for (int i = 0; i < n; ++i) {
		const uint32_t p = c[i];
		const int v[3] = {px_r(p), px_g(p), px_b(p)};
		for (int ch = 0; ch < 3; ++ch) {
			lo[ch] = v[ch] < lo[ch] ? v[ch] : lo[ch];
			hi[ch] = v[ch] > hi[ch] ? v[ch] : hi[ch];
		}
		amin = ca[i] < amin ? ca[i] : amin;
		amax = ca[i] > amax ? ca[i] : amax;
	}
	CandBox b;
	for (int ch = 0; ch < 3; ++ch) {
		b.lo[ch] = lo[ch];
		b.len[ch] = hi[ch] - lo[ch] + 1;
	}
	b.alo = amin;
	b.alen = amax - amin + 1;
	return b;
}

// One candidate = 3 draws (r, g, b) and, for DXT5 only, a 4th for alpha, in that order (ref :986-990).
// RNG must provide int next() returning glibc rand() values in stream order.
// Returns a packed texel (alpha byte 0 unless DXT5).
template <int DXT, class RNG>
S2TC_HD uint32_t draw_candidate(const CandBox &b, RNG &rng)
{
	const int r = b.lo[0] + rng.next() % b.len[0];
	const int g = b.lo[1] + rng.next() % b.len[1];
	const int bl = b.lo[2] + rng.next() % b.len[2];
	int a = 0;
	if (DXT == kDxt5)
		a = b.alo + rng.next() % b.alen;
	return px_make(r, g, bl, a);
}

// ---- pair search, one thread (ref :367-478) ----------------------------------------------------
// d is scratch for m*n (+2n for alpha) ints.
template <int CD>
S2TC_HD void search_colors_scalar(uint32_t *c, int n, int m, int *d)
{
	typedef Metric<CD> M;
	for (int i = 0; i < m; ++i) {
		const typename M::Feat fi = M::feat(c[i]);
		for (int k = 0; k < n; ++k) {
			int v;
			if (i == k)
				v = 0;
			else if (i < n && k < i) // square part: the lower index is the first argument
				v = M::dist(M::feat(c[k]), fi);
			else
				v = M::dist(fi, M::feat(c[k]));
			d[i * n + k] = v;
		}
	}
	int bestsum = -1, bi = 0, bj = 1;
	for (int i = 0; i < m; ++i)
		for (int j = i + 1; j < m; ++j) {
			int sum = 0;
			for (int k = 0; k < n; ++k) {
				int a = d[i * n + k], bq = d[j * n + k];
				sum = wadd(sum, a < bq ? a : bq);
			}
			if (bestsum < 0 || sum < bestsum) {
				bestsum = sum;
				bi = i;
				bj = j;
			}
		}
	uint32_t keep = c[bi];
	c[1] = c[bj];
	c[0] = keep;
}

S2TC_HD void search_alpha_scalar(uint8_t *a, int n, int m, int *d)
{
	for (int i = 0; i < m; ++i)
		for (int k = 0; k < n; ++k)
			d[i * n + k] = alpha_dist(a[i], a[k]);
	for (int k = 0; k < n; ++k) { // the fixed points 0 and 255 folded into one row
		int z = alpha_dist(0, a[k]), f = alpha_dist(255, a[k]);
		d[m * n + k] = z < f ? z : f;
	}
	int bestsum = -1, bi = 0, bj = 1;
	for (int i = 0; i < m; ++i)
		for (int j = i + 1; j < m; ++j) {
			int sum = 0;
			for (int k = 0; k < n; ++k) {
				int v = d[i * n + k], w = d[j * n + k], f = d[m * n + k];
				v = v < w ? v : w;
				sum += v < f ? v : f;
			}
			if (bestsum < 0 || sum < bestsum) {
				bestsum = sum;
				bi = i;
				bj = j;
			}
		}
	if (bi != 0)
		a[0] = a[bi];
	if (bj != 1)
		a[1] = a[bj];
}

// ---- index assignment + refinement, colour (ref :582-643, :767-860) ----------------------------
// One pass: every usable texel picks the nearer of ref0/ref1 (strict <, ref1 wins only if closer).
// m1 = texels that picked 1; s1 = packed channel sums (r<<20 | g<<10 | b) of those texels.
template <int CD>
S2TC_HD uint32_t assign_colors(const typename Metric<CD>::Feat *pf, const uint32_t *pk, uint32_t use,
		uint32_t ref0, uint32_t ref1, uint32_t &m1, uint32_t &s1)
{
	typedef Metric<CD> M;
	const typename M::Feat f0 = M::feat(ref0), f1 = M::feat(ref1);
	uint32_t score = 0, mask = 0, sum = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		if ((use >> i) & 1u) {
			int d0 = M::dist(pf[i], f0);
			int d1 = M::dist(pf[i], f1);
			bool pick1 = d1 < d0;
			score += (uint32_t) (pick1 ? d1 : d0);
			if (pick1) {
				mask |= 1u << i;
				sum += pk[i];
			}
		}
	}
	m1 = mask;
	s1 = sum;
	return score;
}

// rounded per-channel mean of a packed sum over cnt texels (ref :542-551, :201-208)
S2TC_HD uint32_t mean_color(uint32_t packed, int cnt)
{
	const uint32_t rc = half_recip18(cnt);
	int r = div_by_2n((int) (((packed >> 20) & 0x3FF) << 1) + cnt, rc) & 31;
	int g = div_by_2n((int) (((packed >> 10) & 0x3FF) << 1) + cnt, rc) & 63;
	int b = div_by_2n((int) ((packed & 0x3FF) << 1) + cnt, rc) & 31;
	return px_make(r, g, b);
}

// Returns the 2-bit index word; c0/c1 are updated to the final endpoints.
// use  = valid texels that take part (DXT1: alpha != 0);  trans = valid texels coded 3 (DXT1 only)
template <int CD, bool HAVE_TRANS>
S2TC_HD uint32_t refine_colors(const Block &b, int refine, uint32_t use, uint32_t trans, uint32_t &c0, uint32_t &c1)
{
	typedef Metric<CD> M;
	typename M::Feat pf[16];
	uint32_t pk[16];
	uint32_t tot = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		const uint32_t p = b.px[i];
		pf[i] = M::feat(p);
		pk[i] = ((uint32_t) px_r(p) << 20) | ((uint32_t) px_g(p) << 10) | (uint32_t) px_b(p);
		if ((use >> i) & 1u)
			tot += pk[i];
	}
	const int ntot = popc16(use);
	uint32_t m1 = 0, m3 = 0, s1 = 0;

	if (refine == kRefineNever) { // ref :848-860
		if (HAVE_TRANS ? col_lt(c1, c0) : col_lt(c0, c1)) {
			uint32_t t = c0; c0 = c1; c1 = t;
		}
		assign_colors<CD>(pf, pk, use, c0, c1, m1, s1);
		return spread2(m1) | (spread2(trans) * 3u);
	}

	if (refine == kRefineAlways) { // ref :816-846
		assign_colors<CD>(pf, pk, use, c0, c1, m1, s1);
		m3 = trans;
		const int n1 = popc16(m1), n0 = ntot - n1;
		if (n0)
			c0 = mean_color(tot - s1, n0);
		if (n1)
			c1 = mean_color(s1, n1);
	} else { // LOOP, ref :767-814
		uint32_t next0 = c0, next1 = c1;
		uint32_t best = 0x7FFFFFFFu;
		for (;;) {
			uint32_t m1n, s1n;
			uint32_t sc = assign_colors<CD>(pf, pk, use, next0, next1, m1n, s1n);
			if (!(sc < best))
				break;
			best = sc;
			m1 = m1n;
			m3 = trans;
			c0 = next0;
			c1 = next1;
			const int n1 = popc16(m1n), n0 = ntot - n1;
			if (!n0 && !n1)
				break;
			if (n0)
				next0 = mean_color(tot - s1n, n0);
			if (n1)
				next1 = mean_color(s1n, n1);
			// fixed point: the next pass would assign identically, score the same and stop on "not <" without
			// changing anything (ref :781-793), so it is not run
			if (next0 == c0 && next1 == c1)
				break;
		}
	}

	if (px_rgb(c0) == px_rgb(c1)) { // ref :796-805: every index that is not 1 becomes 0, 3s included
		c1 = col_bump(c1);
		m3 = 0;
	}
	if (HAVE_TRANS ? col_lt(c1, c0) : col_lt(c0, c1)) { // ref :807-813: flip bit 0 where bit 1 is clear
		uint32_t t = c0; c0 = c1; c1 = t;
		m1 = ~m1 & ~m3 & 0xFFFFu;
	}
	return spread2(m1) | (spread2(m3) * 3u);
}

// ---- DXT5 alpha (ref :582-643 with have_0_255, :645-765) ------------------------------------
struct AlphaPass {
	uint32_t m1, m6, m7; // texels coded 1 / 6 (=0) / 7 (=255)
	int n0, n1, s0, s1;  // cluster sizes and sums (texels coded 6/7 are not accumulated)
	uint32_t score;
};

S2TC_HD AlphaPass assign_alpha(const Block &b, int r0, int r1)
{
	AlphaPass o;
	o.m1 = o.m6 = o.m7 = 0;
	o.n0 = o.n1 = o.s0 = o.s1 = 0;
	o.score = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		if ((b.valid >> i) & 1u) {
			const int a = px_a(b.px[i]);
			const int d0 = alpha_dist(a, r0), d1 = alpha_dist(a, r1);
			const bool pick1 = d1 < d0;
			const int bd = pick1 ? d1 : d0;
			const int dz = a * a, df = (a - 255) * (a - 255);
			if (dz <= bd) {
				o.m6 |= 1u << i;
				o.score += (uint32_t) dz;
			} else if (df <= bd) {
				o.m7 |= 1u << i;
				o.score += (uint32_t) df;
			} else {
				o.score += (uint32_t) bd;
				if (pick1) {
					o.m1 |= 1u << i;
					o.n1++;
					o.s1 += a;
				} else {
					o.n0++;
					o.s0 += a;
				}
			}
		}
	}
	return o;
}

S2TC_HD int mean_alpha(int sum, int cnt) { return div_by_2n((sum << 1) + cnt, half_recip18(cnt)) & 255; }

// returns the 48-bit index word; a0/a1 updated
S2TC_HD uint64_t refine_alpha(const Block &b, int refine, int &a0, int &a1)
{
	uint32_t m1 = 0, m6 = 0, m7 = 0;
	if (refine == kRefineNever) { // ref :755-765
		if (a1 < a0) {
			int t = a0; a0 = a1; a1 = t;
		}
		AlphaPass p = assign_alpha(b, a0, a1);
		m1 = p.m1; m6 = p.m6; m7 = p.m7;
	} else {
		if (refine == kRefineAlways) { // ref :709-752
			AlphaPass p = assign_alpha(b, a0, a1);
			m1 = p.m1; m6 = p.m6; m7 = p.m7;
			if (p.n0)
				a0 = mean_alpha(p.s0, p.n0);
			if (p.n1)
				a1 = mean_alpha(p.s1, p.n1);
		} else { // ref :646-671
			int next0 = a0, next1 = a1;
			uint32_t best = 0x7FFFFFFFu;
			for (;;) {
				AlphaPass p = assign_alpha(b, next0, next1);
				if (!(p.score < best))
					break;
				best = p.score;
				m1 = p.m1; m6 = p.m6; m7 = p.m7;
				a0 = next0;
				a1 = next1;
				if (!p.n0 && !p.n1)
					break;
				if (p.n0)
					next0 = mean_alpha(p.s0, p.n0);
				if (p.n1)
					next1 = mean_alpha(p.s1, p.n1);
				if (next0 == a0 && next1 == a1) // fixed point, as in refine_colors
					break;
			}
		}
		if (a1 == a0) { // ref :673-685: codes 1 -> 0
			a1 = alpha_bump(a0);
			m1 = 0;
		}
		if (a1 < a0) { // ref :687-705: 0 <-> 1 on all 16 positions, 6/7 stay
			int t = a0; a0 = a1; a1 = t;
			m1 = ~m1 & ~m6 & ~m7 & 0xFFFFu;
		}
	}
	const uint32_t hi = m6 | m7; // 6 = 110b, 7 = 111b
	return spread3(m1 | m7) | (spread3(hi) << 1) | (spread3(hi) << 2);
}

// ---- the tail shared by every mode (ref :1010-1107) --------------------------------------------
// Takes the endpoints chosen by the candidate stage and writes the finished block:
// out[0..1] for DXT1, out[0..3] for DXT3/DXT5 (little-endian words of the byte layout in A.10).
template <int DXT, int CD>
S2TC_HD void finish_block(const Block &b, int refine, uint32_t c0, uint32_t c1, int a0, int a1, uint32_t *out)
{
	// "equal colors are BAD" (ref :1010-1027)
	if (px_rgb(c0) == px_rgb(c1))
		c1 = col_bump(c1);
	if (DXT == kDxt5 && a0 == a1)
		a1 = alpha_bump(a0);

	uint32_t use = b.valid, trans = 0;
	if (DXT == kDxt1) {
#pragma unroll
		for (int i = 0; i < 16; ++i)
			if (px_a(b.px[i]) == 0)
				trans |= 1u << i;
		trans &= b.valid;
		use &= ~trans;
	}
	const uint32_t idx = refine_colors<CD, DXT == kDxt1>(b, refine, use, trans, c0, c1);
	const uint32_t ends = to565(c0) | (to565(c1) << 16);

	if (DXT == kDxt1) {
		out[0] = ends;
		out[1] = idx;
	} else if (DXT == kDxt3) { // ref :862-870: 4-bit alpha of every valid texel
		uint32_t lo = 0, hi = 0;
#pragma unroll
		for (int i = 0; i < 8; ++i)
			if ((b.valid >> i) & 1u)
				lo |= (uint32_t) px_a(b.px[i]) << (4 * i);
#pragma unroll
		for (int i = 8; i < 16; ++i)
			if ((b.valid >> i) & 1u)
				hi |= (uint32_t) px_a(b.px[i]) << (4 * (i - 8));
		out[0] = lo;
		out[1] = hi;
		out[2] = ends;
		out[3] = idx;
	} else {
		const uint64_t aidx = refine_alpha(b, refine, a0, a1);
		out[0] = (uint32_t) a0 | ((uint32_t) a1 << 8) | ((uint32_t) (aidx & 0xFFFFu) << 16);
		out[1] = (uint32_t) (aidx >> 16);
		out[2] = ends;
		out[3] = idx;
	}
}

} // namespace s2tc
