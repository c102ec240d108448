/*
 * s2tc_b200_algorithm.h -- the reference's "algorithm API" as exported by this library
 * (reference declarations: s2tc_algorithm.h:31-66; definitions: s2tc_algorithm.cpp:1164-1194 and
 * :1453-1465).  Same names, same enumerator values, same call shapes; every call runs on the GPU.
 *
 * The reference's factory hands out a per-block function pointer, which a host loop then calls
 * once per 4x4 block (s2tc_libtxc_dxtn.cpp:246-294).  That shape is kept for source compatibility --
 * each call encodes one block on the device -- but it is a latency-bound way to drive a GPU.
 * Whole images should go through tx_compress_dxtn or s2tc_b200_compress_host /
 * s2tc_b200_encode_rows_device (include/s2tc_b200.h), which is what tx_compress_dxtn itself uses.
 *
 * BASELINE.json calls the factory get_s2tc_encoder(); upstream's symbol is s2tc_encode_block_func.
 * Both names are exported; they are the same function.
 */
#ifndef S2TC_B200_ALGORITHM_H
#define S2TC_B200_ALGORITHM_H

#ifdef __cplusplus
extern "C" {
#endif

enum DitherMode { DITHER_NONE, DITHER_SIMPLE, DITHER_FLOYDSTEINBERG };
enum DxtMode { DXT1, DXT3, DXT5 };
enum RefinementMode { REFINE_NEVER, REFINE_ALWAYS, REFINE_LOOP };
typedef enum { RGB, YUV, SRGB, SRGB_MIXED, AVG, WAVG, W0AVG, NORMALMAP } ColorDistMode;

/* RGB(A)8 -> 4 bytes per texel {r5, g6, b5, a reduced to alphabits}; srccomps 3 or 4, alphabits 1/4/8 */
void rgb565_image(unsigned char *out, const unsigned char *rgba, int w, int h, int srccomps, int alphabits,
		enum DitherMode dither);

/* out: 8 (DXT1) or 16 bytes; rgba: first texel of the block inside a reduced image of row stride iw
 * texels; w, h in 1..4 valid texels; nrandom as given to the factory */
typedef void (*s2tc_encode_block_func_t)(unsigned char *out, const unsigned char *rgba, int iw, int w, int h, int nrandom);

s2tc_encode_block_func_t s2tc_encode_block_func(enum DxtMode dxt, ColorDistMode cd, int nrandom, enum RefinementMode refine);
s2tc_encode_block_func_t get_s2tc_encoder(enum DxtMode dxt, ColorDistMode cd, int nrandom, enum RefinementMode refine);

#ifdef __cplusplus
}
#endif
#endif
