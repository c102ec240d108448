/*
 * ref_harness.c -- drives a CPU encoder over block rows on several host threads and times it.
 * TEST / BENCH INFRASTRUCTURE ONLY (see s2tc_oracle.h): used by bench.py's cpu_baseline and
 * --impl reference legs and by tests; never by the shipped encoder.
 *
 * Two back ends:
 *   refh_*  : the UNMODIFIED upstream encoder, dlopen()ed from oracle/_ref/ (built by
 *             oracle/Makefile from the upstream sources).  The image loop below mirrors
 *             ref: s2tc_libtxc_dxtn.cpp:246-294 but hands disjoint block-row ranges to
 *             worker threads; upstream itself is single-threaded.
 *   orch_*  : the restatement in s2tc_oracle.c, same threading.
 */
#define _GNU_SOURCE
#include "s2tc_oracle.h"

#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void (*ref_block_fn)(unsigned char *out, const unsigned char *rgba, int iw, int w, int h, int nrandom);
typedef ref_block_fn (*ref_factory_fn)(int dxt, int cd, int nrandom, int refine);
typedef void (*ref_prepass_fn)(unsigned char *out, const unsigned char *rgba, int w, int h, int srccomps, int alphabits, int dither);
typedef void (*ref_seek_fn)(uint64_t draws);
typedef void (*ref_compress_fn)(int srccomps, int width, int height, const unsigned char *src, unsigned destformat, unsigned char *dest, int stride);

typedef struct {
	void *so;
	ref_factory_fn factory;
	ref_prepass_fn prepass;
	ref_seek_fn seek; /* NULL unless the .so was built with -Drand=s2tc_tls_rand */
	ref_compress_fn compress;
} refh_t;

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void *refh_open(const char *path)
{
	refh_t *h = (refh_t *) calloc(1, sizeof(*h));
	h->so = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!h->so) {
		fprintf(stderr, "refh_open: %s\n", dlerror());
		free(h);
		return NULL;
	}
	h->factory = (ref_factory_fn) dlsym(h->so, "s2tc_encode_block_func");
	h->prepass = (ref_prepass_fn) dlsym(h->so, "rgb565_image");
	h->compress = (ref_compress_fn) dlsym(h->so, "tx_compress_dxtn");
	h->seek = (ref_seek_fn) dlsym(h->so, "s2tc_tls_rand_seek");
	if (!h->factory || !h->prepass || !h->compress) {
		fprintf(stderr, "refh_open: %s lacks the s2tc entry points\n", path);
		dlclose(h->so);
		free(h);
		return NULL;
	}
	return h;
}

int refh_has_seek(void *hv) { return ((refh_t *) hv)->seek != NULL; }

void refh_close(void *hv)
{
	refh_t *h = (refh_t *) hv;
	if (h) {
		dlclose(h->so);
		free(h);
	}
}

/* single block through the factory (used by block-level parity tests) */
void refh_encode_block(void *hv, unsigned char *out, const unsigned char *rgba, int iw, int w, int h,
		int dxt, int cd, int nrandom, int refine)
{
	refh_t *r = (refh_t *) hv;
	r->factory(dxt, cd, nrandom, refine)(out, rgba, iw, w, h, nrandom);
}

void refh_prepass(void *hv, unsigned char *out, const unsigned char *src, int w, int h, int srccomps, int alphabits, int dither)
{
	((refh_t *) hv)->prepass(out, src, w, h, srccomps, alphabits, dither);
}

void refh_seek(void *hv, uint64_t draws)
{
	refh_t *r = (refh_t *) hv;
	if (r->seek)
		r->seek(draws);
}

typedef struct {
	refh_t *ref; /* NULL -> oracle restatement */
	const unsigned char *reduced;
	int width, height, dxt, cd, nrandom, refine;
	uint64_t cursor0;
	unsigned char *dest;
	int row_bytes;
	int row0, row1;
} job_t;

static void *worker(void *arg)
{
	job_t *j = (job_t *) arg;
	int bs = j->dxt == ORC_DXT1 ? 8 : 16;
	int bw = (j->width + 3) / 4;
	if (!j->ref) {
		orc_encode_block_rows(j->reduced, j->width, j->height, j->row0, j->row1, j->dxt, j->cd,
				j->nrandom, j->refine, j->cursor0, j->dest, j->row_bytes);
		return NULL;
	}
	{
		ref_block_fn f = j->ref->factory(j->dxt, j->cd, j->nrandom, j->refine);
		int dpb = j->nrandom > 0 ? j->nrandom * (j->dxt == ORC_DXT5 ? 4 : 3) : 0;
		int by, bx;
		if (dpb && j->ref->seek)
			j->ref->seek(j->cursor0 + (uint64_t) j->row0 * bw * dpb);
		for (by = j->row0; by < j->row1; ++by) {
			int y = by * 4;
			int ny = j->height > y + 3 ? 4 : j->height - y;
			unsigned char *blk = j->dest + (size_t) by * j->row_bytes;
			for (bx = 0; bx < bw; ++bx) {
				int x = bx * 4;
				int nx = j->width > x + 3 ? 4 : j->width - x;
				f(blk, j->reduced + ((size_t) y * j->width + x) * 4, j->width, nx, ny, j->nrandom);
				blk += bs;
			}
		}
	}
	return NULL;
}

/* Encodes block rows [row0,row1) of an image with `nthreads` workers.
 * times[0] = 565 pre-pass seconds (single thread, whole image), times[1] = block loop seconds.
 * hv == NULL selects the oracle restatement.  dest is indexed as the full image would be.
 * Returns 0, -1 bad format, -2 threads>1 requested for nrandom>0 without a seekable rand. */
int refh_encode_mt(void *hv, int srccomps, int width, int height, const unsigned char *src,
		unsigned destformat, int dither, int cd, int nrandom, int refine, uint64_t cursor0,
		unsigned char *dest, int dst_row_stride, int row0, int row1, int nthreads, double *times)
{
	refh_t *ref = (refh_t *) hv;
	int dxt, alphabits, bs, tight, row_bytes, t, nrows;
	unsigned char *reduced;
	pthread_t *th;
	job_t *jobs;
	double t0, t1, t2;

	switch (destformat) {
	case 0x83F0:
	case 0x83F1: dxt = ORC_DXT1; alphabits = 1; break;
	case 0x83F2: dxt = ORC_DXT3; alphabits = 4; break;
	case 0x83F3: dxt = ORC_DXT5; alphabits = 8; break;
	default: return -1;
	}
	if (ref && nrandom > 0 && nthreads > 1 && !ref->seek)
		return -2;
	bs = dxt == ORC_DXT1 ? 8 : 16;
	tight = ((width + 3) & ~3) * (bs / 4);
	row_bytes = dst_row_stride >= width * (bs / 4) ? dst_row_stride : tight;
	nrows = row1 - row0;
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > nrows)
		nthreads = nrows > 0 ? nrows : 1;

	reduced = (unsigned char *) malloc((size_t) width * height * 4 + 4);
	t0 = now_s();
	if (ref)
		ref->prepass(reduced, src, width, height, srccomps, alphabits, dither);
	else
		orc_rgb565_image(reduced, src, width, height, srccomps, alphabits, dither);
	t1 = now_s();

	th = (pthread_t *) calloc(nthreads, sizeof(*th));
	jobs = (job_t *) calloc(nthreads, sizeof(*jobs));
	for (t = 0; t < nthreads; ++t) {
		job_t *j = &jobs[t];
		j->ref = ref;
		j->reduced = reduced;
		j->width = width; j->height = height;
		j->dxt = dxt; j->cd = cd; j->nrandom = nrandom; j->refine = refine;
		j->cursor0 = cursor0;
		j->dest = dest;
		j->row_bytes = row_bytes;
		j->row0 = row0 + (int) ((long long) nrows * t / nthreads);
		j->row1 = row0 + (int) ((long long) nrows * (t + 1) / nthreads);
		if (nthreads == 1)
			worker(j);
		else
			pthread_create(&th[t], NULL, worker, j);
	}
	if (nthreads > 1)
		for (t = 0; t < nthreads; ++t)
			pthread_join(th[t], NULL);
	t2 = now_s();
	if (times) {
		times[0] = t1 - t0;
		times[1] = t2 - t1;
	}
	free(th);
	free(jobs);
	free(reduced);
	return 0;
}

/* the reference's own entry point, untouched (reads the S2TC_* environment itself) */
void refh_tx_compress(void *hv, int srccomps, int width, int height, const unsigned char *src,
		unsigned destformat, unsigned char *dest, int stride)
{
	((refh_t *) hv)->compress(srccomps, width, height, src, destformat, dest, stride);
}
