/*
 * s2tc_oracle.h -- CPU restatement of the S2TC encode hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may link or call anything in oracle/.  The shipped encoder
 * (s2tc_b200/) never does; it fails loudly when its CUDA library is missing.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_vs_ref.py) against
 * the UNMODIFIED upstream sources compiled into oracle/_ref/ by oracle/Makefile,
 * against the committed vectors in tests/golden/ that were generated from that
 * build (tests/golden/make_golden.py), and against the block-level known
 * answers recorded in SURVEY.md App. B.4.  Upstream itself ships no golden
 * vectors.
 *
 * Every function names the reference lines it restates ("ref:" = path relative
 * to the upstream checkout, divVerent/s2tc).
 */
#ifndef S2TC_ORACLE_H
#define S2TC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numeric values follow ref: s2tc_algorithm.h:31-63 */
enum { ORC_DITHER_NONE = 0, ORC_DITHER_SIMPLE = 1, ORC_DITHER_FLOYDSTEINBERG = 2 };
enum { ORC_DXT1 = 0, ORC_DXT3 = 1, ORC_DXT5 = 2 };
enum { ORC_REFINE_NEVER = 0, ORC_REFINE_ALWAYS = 1, ORC_REFINE_LOOP = 2 };
enum { ORC_RGB = 0, ORC_YUV, ORC_SRGB, ORC_SRGB_MIXED, ORC_AVG, ORC_WAVG, ORC_W0AVG, ORC_NORMALMAP };

/* ---- glibc rand() replica (TYPE_3 additive feedback, seed 1) ------------- */
typedef struct {
	uint32_t win[31]; /* r[pos .. pos+30], r[] in the flat indexing of orc_rand_seek */
	int head;         /* index in win[] of r[pos] */
	uint64_t draws;   /* number of rand() results handed out so far */
} orc_rand_t;

void orc_rand_init(orc_rand_t *g);                  /* == fresh process, never srand()ed */
void orc_rand_seek(orc_rand_t *g, uint64_t draws);  /* position so the next result is draw #draws */
int orc_rand_next(orc_rand_t *g);

/* ---- colour metrics (ref: s2tc_algorithm.cpp:215-361) --------------------- */
/* a and b are {r5,g6,b5}; argument order matters for ORC_SRGB */
int orc_color_dist(int cd, const signed char a[3], const signed char b[3]);
int orc_alpha_dist(int a, int b);

/* ---- 565 pre-pass (ref: s2tc_algorithm.cpp:1196-1465) -------------------- */
void orc_rgb565_image(unsigned char *out, const unsigned char *src, int w, int h,
		int srccomps, int alphabits, int dither);

/* ---- one 4x4 block (ref: s2tc_algorithm.cpp:872-1194) --------------------- */
/* rgba: block's top-left texel inside a pre-reduced 4-byte/pixel image of row stride iw pixels.
 * rng may be NULL when nrandom <= 0. */
void orc_encode_block(unsigned char *out, const unsigned char *rgba, int iw, int w, int h,
		int dxt, int cd, int nrandom, int refine, orc_rand_t *rng);

/* ---- whole image (ref: s2tc_libtxc_dxtn.cpp:142-299), settings passed explicitly */
/* destformat is the GL enum 0x83F0..0x83F3; returns 0, or -1 for a bad destformat (dest untouched) */
int orc_compress_image(int srccomps, int width, int height, const unsigned char *src,
		unsigned int destformat, unsigned char *dest, int dst_row_stride,
		int dither, int cd, int nrandom, int refine, orc_rand_t *rng);

/* same, but only block rows [row0, row1) of the image are encoded, into dest + the offset the
 * full-image call would use; `reduced` must be the complete output of orc_rgb565_image and rng is
 * sought to the cursor of row0 (cursor0 + row0 * blocks_per_row * draws_per_block). */
void orc_encode_block_rows(const unsigned char *reduced, int width, int height, int row0, int row1,
		int dxt, int cd, int nrandom, int refine, uint64_t cursor0,
		unsigned char *dest, int dst_row_stride);

/* ---- S3TC -> S2TC transcode (ref: s2tc_from_s3tc.cpp:77-190, 254-263) ---- */
void orc_transcode_blocks(unsigned char *blocks, size_t nblocks, int dxt);

/* ---- decode (ref: s2tc_libtxc_dxtn.cpp:35-140) --------------------------- */
void orc_fetch_texel(int dxt, int rgb_only, int src_row_stride, const unsigned char *pixdata,
		int i, int j, unsigned char texel[4]);

/* ---- mip reduce (ref: s2tc_compress.c:427-493) --------------------------- */
void orc_mip_reduce(const unsigned char *in, unsigned char *out, int *width, int *height,
		int destwidth, int destheight);

#ifdef __cplusplus
}
#endif
#endif
