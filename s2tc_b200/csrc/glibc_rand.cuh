// glibc_rand.cuh -- bit-exact, seekable replay of glibc's rand() stream, host+device.
//
// The reference draws its random candidate colours from libc rand() and never seeds it
// (s2tc_algorithm.cpp:986-990), so byte-exact output for S2TC_RANDOM_COLORS > 0 requires the
// exact stream: glibc's TYPE_3 additive-feedback generator with seed 1.  In flat form
//     q[i+31] = q[i] + q[i+28]   (mod 2^32),       rand() number k  =  q[k + 341] >> 1,
// where q[0..30] is the seeded state rotated by three (q[i] = r[i+3], r[31..33] = r[0..2]).
//
// The recurrence is linear over Z/2^32, so the 31-word window at any position is
//     q[e+s] = sum_j a_j q[s+j],  with  sum_j a_j x^j = x^e mod (x^31 - x^28 - 1):
// every block's draws are reachable in O(log e) polynomial products, which is what lets
// thousands of threads generate disjoint parts of one sequential stream.
#pragma once

#include "s2tc_defs.h"

namespace s2tc {

constexpr int kLag = 31;
constexpr int kRandWarmup = 310; // window of draw k starts at q[k + 310]

struct Poly { uint32_t c[kLag]; };

// out = a*b mod (x^31 - x^28 - 1) over Z/2^32; out may alias a or b
S2TC_HD void poly_mulmod(Poly &out, const Poly &a, const Poly &b)
{
	uint32_t t[2 * kLag - 1];
	for (int i = 0; i < 2 * kLag - 1; ++i)
		t[i] = 0;
	for (int i = 0; i < kLag; ++i) {
		const uint32_t ai = a.c[i];
		for (int j = 0; j < kLag; ++j)
			t[i + j] += ai * b.c[j];
	}
	for (int i = 2 * kLag - 2; i >= kLag; --i) { // x^i = x^(i-3) + x^(i-31)
		t[i - 3] += t[i];
		t[i - kLag] += t[i];
	}
	for (int i = 0; i < kLag; ++i)
		out.c[i] = t[i];
}

S2TC_HD void poly_one(Poly &p)
{
	for (int i = 0; i < kLag; ++i)
		p.c[i] = 0;
	p.c[0] = 1;
}

// x^e mod P by square-and-multiply (host-side setup; O(log e) products)
inline Poly poly_xpow(uint64_t e)
{
	Poly acc, sq;
	poly_one(acc);
	for (int i = 0; i < kLag; ++i)
		sq.c[i] = 0;
	sq.c[1] = 1;
	while (e) {
		if (e & 1)
			poly_mulmod(acc, acc, sq);
		poly_mulmod(sq, sq, sq);
		e >>= 1;
	}
	return acc;
}

// q[0 .. 2*31-2], the base values every window is a linear combination of
inline void rand_base(uint32_t *q /* [61] */)
{
	uint32_t r[34];
	r[0] = 1;
	for (int i = 1; i < 31; ++i) { // glibc srandom_r: 16807 * r mod (2^31-1) by Schrage
		int32_t prev = (int32_t) r[i - 1];
		int32_t hi = prev / 127773, lo = prev % 127773;
		int32_t word = 16807 * lo - 2836 * hi;
		if (word < 0)
			word += 2147483647;
		r[i] = (uint32_t) word;
	}
	for (int i = 31; i < 34; ++i)
		r[i] = r[i - 31];
	for (int i = 0; i < 31; ++i)
		q[i] = r[i + 3];
	for (int i = 31; i < 2 * kLag - 1; ++i)
		q[i] = q[i - 31] + q[i - 3];
}

// Sequential generator over a 31-word circular window.
struct GlibcRand {
	uint32_t w[kLag];
	int head;

	// window = q[e .. e+30] given a = x^e mod P and the base values
	S2TC_HD void load(const Poly &a, const uint32_t *base)
	{
		for (int s = 0; s < kLag; ++s) {
			uint32_t v = 0;
			for (int j = 0; j < kLag; ++j)
				v += a.c[j] * base[s + j];
			w[s] = v;
		}
		head = 0;
	}
	S2TC_HD int next()
	{
		int k = head + 28;
		if (k >= kLag)
			k -= kLag;
		const uint32_t v = w[head] + w[k];
		w[head] = v;
		head = head + 1 == kLag ? 0 : head + 1;
		return (int) (v >> 1);
	}
};

// Host-side description of a stream segmentation: thread t of a launch starts at draw
// cursor0 + t * stride.  The device rebuilds its window from `start` (= x^(cursor0+310)) and the
// powers x^(stride * 2^j).
struct RandPlan {
	Poly start;
	Poly step[32];
	uint32_t base[2 * kLag - 1];
};

inline void rand_plan_init(RandPlan &p, uint64_t cursor0, uint64_t stride)
{
	p.start = poly_xpow(cursor0 + kRandWarmup);
	p.step[0] = poly_xpow(stride);
	for (int j = 1; j < 32; ++j)
		poly_mulmod(p.step[j], p.step[j - 1], p.step[j - 1]);
	rand_base(p.base);
}

// window of segment t
S2TC_HD void rand_plan_seek(const RandPlan &p, uint32_t t, GlibcRand &g)
{
	Poly a = p.start;
	for (int j = 0; t; ++j, t >>= 1)
		if (t & 1u)
			poly_mulmod(a, a, p.step[j]);
	g.load(a, p.base);
}

} // namespace s2tc
