/*
 * tls_rand.c -- thread-local stand-in for libc rand(), TEST INFRASTRUCTURE ONLY.
 *
 * The upstream encoder draws its random candidate colours from libc rand()
 * (ref: s2tc_algorithm.cpp:986-990), a process-global stream behind a lock.  To time the
 * reference on all host cores without changing a line of it, oracle/Makefile compiles the
 * upstream translation unit a second time with -Drand=s2tc_tls_rand; each worker thread of
 * ref_harness.c then seeks its private replica to the draw index its first block would have
 * seen in the single-threaded run, so the threaded output stays byte-identical to the
 * sequential one.
 */
#include "s2tc_oracle.h"

static __thread orc_rand_t tls_gen;
static __thread int tls_ready;

int s2tc_tls_rand(void)
{
	if (!tls_ready) {
		orc_rand_init(&tls_gen);
		tls_ready = 1;
	}
	return orc_rand_next(&tls_gen);
}

void s2tc_tls_rand_seek(uint64_t draws)
{
	orc_rand_seek(&tls_gen, draws);
	tls_ready = 1;
}
