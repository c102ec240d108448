/*
 * orc_rand.c -- seekable replica of glibc's rand() stream (TEST INFRASTRUCTURE ONLY,
 * see s2tc_oracle.h).  Split from s2tc_oracle.c so that the thread-local shim used for the
 * multithreaded CPU baseline (tls_rand.c) can link it without the rest of the oracle.
 */
#include "s2tc_oracle.h"

#include <string.h>

/* ========================================================================== */
/* glibc rand(): TYPE_3, x^31 + x^3 + 1 additive feedback, default seed 1      */
/* (glibc stdlib/random_r.c; replayed because ref: s2tc_algorithm.cpp:986-990  */
/* draws its random candidates from libc rand() and never calls srand()).      */
/*                                                                             */
/* Flat model: q[0..30] = state after srandom(1) rotated so that               */
/*   q[i+31] = q[i] + q[i+28]  (mod 2^32) holds for every i >= 0,              */
/* and the k-th rand() result (k = 0,1,..) is q[k+341] >> 1.                   */
/* ========================================================================== */

#define ORC_LAG 31

static void orc_rand_base(uint32_t *q, int n)
{
	uint32_t r[34];
	int i;
	r[0] = 1;
	for (i = 1; i < 31; ++i) {
		/* 16807 * r[i-1] mod (2^31 - 1), Schrage's method as glibc does it */
		int32_t prev = (int32_t) r[i - 1];
		int32_t hi = prev / 127773, lo = prev % 127773;
		int32_t word = 16807 * lo - 2836 * hi;
		if (word < 0)
			word += 2147483647;
		r[i] = (uint32_t) word;
	}
	for (i = 31; i < 34; ++i)
		r[i] = r[i - 31];
	for (i = 0; i < 31 && i < n; ++i)
		q[i] = r[i + 3];
	for (i = 31; i < n; ++i)
		q[i] = q[i - 31] + q[i - 3];
}

/* multiply two residues modulo x^31 - x^28 - 1 over Z/2^32 */
static void orc_poly_mulmod(uint32_t *out, const uint32_t *a, const uint32_t *b)
{
	uint32_t t[2 * ORC_LAG - 1];
	int i, j;
	memset(t, 0, sizeof(t));
	for (i = 0; i < ORC_LAG; ++i) {
		if (!a[i])
			continue;
		for (j = 0; j < ORC_LAG; ++j)
			t[i + j] += a[i] * b[j];
	}
	for (i = 2 * ORC_LAG - 2; i >= ORC_LAG; --i) {
		t[i - 3] += t[i];
		t[i - ORC_LAG] += t[i];
	}
	memcpy(out, t, ORC_LAG * sizeof(uint32_t));
}

void orc_rand_seek(orc_rand_t *g, uint64_t draws)
{
	/* window = q[draws+310 .. draws+340] */
	uint64_t e = draws + 310;
	uint32_t acc[ORC_LAG], sq[ORC_LAG], base[2 * ORC_LAG - 1];
	int i, j;
	memset(acc, 0, sizeof(acc));
	memset(sq, 0, sizeof(sq));
	acc[0] = 1; /* x^0 */
	sq[1] = 1;  /* x^1 */
	while (e) {
		if (e & 1)
			orc_poly_mulmod(acc, acc, sq);
		orc_poly_mulmod(sq, sq, sq);
		e >>= 1;
	}
	orc_rand_base(base, 2 * ORC_LAG - 1);
	for (i = 0; i < ORC_LAG; ++i) {
		uint32_t v = 0;
		for (j = 0; j < ORC_LAG; ++j)
			v += acc[j] * base[i + j];
		g->win[i] = v;
	}
	g->head = 0;
	g->draws = draws;
}

void orc_rand_init(orc_rand_t *g)
{
	uint32_t base[341];
	orc_rand_base(base, 341);
	memcpy(g->win, base + 310, sizeof(g->win));
	g->head = 0;
	g->draws = 0;
}

int orc_rand_next(orc_rand_t *g)
{
	int h = g->head;
	int k = h + 28;
	uint32_t v;
	if (k >= ORC_LAG)
		k -= ORC_LAG;
	v = g->win[h] + g->win[k];
	g->win[h] = v;
	g->head = (h + 1 == ORC_LAG) ? 0 : h + 1;
	g->draws++;
	return (int) (v >> 1);
}

