/* Minimal stand-in for <GL/gl.h>, used ONLY to compile the upstream reference
 * sources (which include it from txc_dxtn.h) into oracle/_ref/ and to compile
 * our own libtxc_dxtn-compatible host shim.  The GL headers are not installed
 * in this image; the four typedefs and four enum values below are the whole
 * of what the libtxc_dxtn ABI needs (values from the EXT_texture_compression_s3tc
 * registry entry). */
#ifndef S2TC_B200_GL_STUB_H
#define S2TC_B200_GL_STUB_H
typedef unsigned int GLenum;
typedef int GLint;
typedef unsigned char GLubyte;
typedef void GLvoid;
#define GL_COMPRESSED_RGB_S3TC_DXT1_EXT 0x83F0
#define GL_COMPRESSED_RGBA_S3TC_DXT1_EXT 0x83F1
#define GL_COMPRESSED_RGBA_S3TC_DXT3_EXT 0x83F2
#define GL_COMPRESSED_RGBA_S3TC_DXT5_EXT 0x83F3
#endif
