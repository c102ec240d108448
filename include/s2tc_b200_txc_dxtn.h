/*
 * s2tc_b200_txc_dxtn.h -- the libtxc_dxtn ABI as exported by this library.
 *
 * These five symbols are what Mesa's texcompress_s3tc and the reference CLI tools dlsym() from
 * "libtxc_dxtn.so" (reference declarations: txc_dxtn.h:38-49; definitions:
 * s2tc_libtxc_dxtn.cpp:35-140 for the texel fetchers, :142-299 for the compressor; users:
 * s2tc_compress.c:47-53, s2tc_decompress.c:54-63).  Signatures are identical, so the shared object
 * built from s2tc_b200/csrc can be passed to `s2tc_compress -l` unchanged.
 *
 * GL scalar types are spelled out (GLint = int, GLenum = unsigned int, GLubyte = unsigned char,
 * GLvoid = void) so that no GL header is needed.
 */
#ifndef S2TC_B200_TXC_DXTN_H
#define S2TC_B200_TXC_DXTN_H

#ifdef __cplusplus
extern "C" {
#endif

/* destformat values (GL_COMPRESSED_*_S3TC_DXT*_EXT) */
#define S2TC_B200_GL_RGB_DXT1 0x83F0
#define S2TC_B200_GL_RGBA_DXT1 0x83F1
#define S2TC_B200_GL_RGBA_DXT3 0x83F2
#define S2TC_B200_GL_RGBA_DXT5 0x83F3

/* Compress width x height texels (srccomps = 3: RGB, else RGBA; tightly packed) into destformat.
 * Encoder settings come from the environment on every call: S2TC_DITHER_MODE, S2TC_COLORDIST_MODE,
 * S2TC_RANDOM_COLORS, S2TC_REFINE_COLORS (ref s2tc_libtxc_dxtn.cpp:156-216).  Runs on the GPU
 * selected by S2TC_B200_DEVICE (default 0).  On any failure a message goes to stderr and dest is left
 * untouched, like the reference's bad-format path (ref :232-235). */
void tx_compress_dxtn(int srccomps, int width, int height, const unsigned char *srcPixData,
		unsigned int destformat, unsigned char *dest, int dstRowStride);

/* Decode one texel (i, j) of an S2TC-encoded image; srcRowStride is the image width in texels. */
void fetch_2d_texel_rgb_dxt1(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel);
void fetch_2d_texel_rgba_dxt1(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel);
void fetch_2d_texel_rgba_dxt3(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel);
void fetch_2d_texel_rgba_dxt5(int srcRowStride, const unsigned char *pixdata, int i, int j, void *texel);

#ifdef __cplusplus
}
#endif
#endif
