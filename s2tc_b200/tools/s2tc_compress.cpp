// s2tc_compress -- TGA -> DDS (DXT1/DXT3/DXT5, S2TC-encoded, full mip chain) on the B200 encoder.
//
// Command-line surface of the reference tool (s2tc_compress.c:544-739): -i infile.tga (default stdin),
// -o outfile.dds (default stdout), -t DXT1|DXT3|DXT5 (default DXT1), -l path_to_libtxc_dxtn.so (default: the
// libtxc_dxtn.so built from s2tc_b200/csrc).  Encoder settings come from the S2TC_* environment, read
// by tx_compress_dxtn itself.  Output is byte-identical to the reference tool for the same input
// (tests/test_gpu_cli.py compares the two).  This file is host plumbing only: every texel is encoded by
// the library's CUDA kernels (the whole mip chain on the device; with -l, level by level through tx_compress_dxtn).
#include <dlfcn.h>
#include <getopt.h>
#include <strings.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/s2tc_b200.h"
#include "../../include/s2tc_b200_txc_dxtn.h"

typedef void (*compress_fn)(int, int, int, const unsigned char *, unsigned int, unsigned char *, int);

static int usage(const char *me)
{
	fprintf(stderr, "usage:\n%s \n    [-i infile.tga]\n    [-o outfile.dds]\n    [-t {DXT1|DXT3|DXT5}]\n    [-l path_to_libtxc_dxtn.so]\n", me);
	return 1;
}

static bool read_all(FILE *f, std::vector<unsigned char> &buf)
{
	unsigned char chunk[65536];
	size_t n;
	while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0)
		buf.insert(buf.end(), chunk, chunk + n);
	return !ferror(f);
}

// Truevision TGA, the subset the reference accepts (s2tc_compress.c:82-424): types 1/9 (colour-mapped, 8-bit
// indices, 24- or 32-bit palette), 2/10 (24/32-bit BGR(A)), 3/11 (8-bit grey), origin top-left or bottom-left,
// 0 or 8 attribute bits (a 32-bit image whose descriptor claims 0 alpha bits is opaque).  Result: RGBA, top row first.
static bool load_tga(const std::vector<unsigned char> &d, int &w, int &h, std::vector<unsigned char> &rgba)
{
	if (d.size() < 19)
		return false;
	const int idlen = d[0], cmaptype = d[1], type = d[2];
	const int cmapindex = d[3] | d[4] << 8, cmaplen = d[5] | d[6] << 8, cmapbits = d[7];
	w = d[12] | d[13] << 8;
	h = d[14] | d[15] << 8;
	const int bpp = d[16], attr = d[17];
	if (w <= 0 || h <= 0 || w > 32768 || h > 32768) {
		fprintf(stderr, "LoadTGA: invalid size\n");
		return false;
	}
	size_t pos = 18 + idlen;
	uint32_t palette[256];
	for (int i = 0; i < 256; ++i)
		palette[i] = 0xFF000000u | (uint32_t) i * 0x010101u; // grey ramp, b g r a in memory order
	if (cmaptype) {
		if (cmaplen > 256 || cmapindex || (cmapbits != 24 && cmapbits != 32)) {
			fprintf(stderr, "LoadTGA: unsupported colormap\n");
			return false;
		}
		const int eb = cmapbits / 8;
		if (pos + (size_t) cmaplen * eb > d.size())
			return false;
		for (int i = 0; i < cmaplen; ++i, pos += eb)
			palette[i] = d[pos] | d[pos + 1] << 8 | d[pos + 2] << 16 | (uint32_t) (eb == 4 ? d[pos + 3] : 255) << 24;
	}
	const int base = type & ~8;
	const bool rle = type & 8;
	if (!((base == 2 && (bpp == 24 || bpp == 32)) || ((base == 1 || base == 3) && bpp == 8))) {
		fprintf(stderr, "LoadTGA: unsupported image type %d / pixel size %d\n", type, bpp);
		return false;
	}
	if (attr & 0x10) {
		fprintf(stderr, "LoadTGA: origin must be in top left or bottom left\n");
		return false;
	}
	const int alphabits = attr & 0x0F;
	if (alphabits != 0 && alphabits != 8) {
		fprintf(stderr, "LoadTGA: only 0 or 8 attribute (alpha) bits supported\n");
		return false;
	}
	const int eb = bpp / 8;
	const bool use_alpha = base == 2 && bpp == 32 && alphabits;
	rgba.assign((size_t) w * h * 4, 0);
	auto texel = [&](size_t p, unsigned char *out) { // file bytes at p -> RGBA
		uint32_t bgra;
		if (base == 2)
			bgra = d[p] | d[p + 1] << 8 | d[p + 2] << 16 | (uint32_t) (use_alpha ? d[p + 3] : 255) << 24;
		else
			bgra = palette[d[p]];
		out[0] = (bgra >> 16) & 0xFF;
		out[1] = (bgra >> 8) & 0xFF;
		out[2] = bgra & 0xFF;
		out[3] = bgra >> 24;
	};
	const size_t n = (size_t) w * h;
	size_t i = 0;
	auto dest = [&](size_t k) { // k-th texel of the file in scan order -> position in the top-down image
		const size_t y = k / w, x = k % w;
		return ((attr & 0x20) ? y : (size_t) h - 1 - y) * w + x;
	};
	while (i < n) {
		if (!rle) {
			if (pos + eb > d.size())
				return false;
			texel(pos, &rgba[dest(i) * 4]);
			pos += eb;
			++i;
			continue;
		}
		if (pos >= d.size())
			return false;
		const int c = d[pos++];
		size_t run = (size_t) (c & 127) + 1;
		if (run > n - i)
			run = n - i;
		if (c & 128) {
			if (pos + eb > d.size())
				return false;
			unsigned char px[4];
			texel(pos, px);
			pos += eb;
			for (size_t k = 0; k < run; ++k)
				memcpy(&rgba[dest(i + k) * 4], px, 4);
		} else {
			if (pos + run * eb > d.size())
				return false;
			for (size_t k = 0; k < run; ++k, pos += eb)
				texel(pos, &rgba[dest(i + k) * 4]);
		}
		i += run;
	}
	return true;
}

// 2x2 box filter of the reference (s2tc_compress.c:427-493): halves every axis that is still > 1,
// odd sizes drop the last row/column, (a + b + c + d) >> 2 (or (a + b) >> 1 on a single axis).
static void mip_reduce(std::vector<unsigned char> &pic, int &w, int &h)
{
	const int nw = w > 1 ? w >> 1 : w, nh = h > 1 ? h >> 1 : h;
	const int sx = w > 1 ? 2 : 1, sy = h > 1 ? 2 : 1;
	const size_t row = (size_t) w * 4;
	std::vector<unsigned char> out((size_t) nw * nh * 4);
	for (int y = 0; y < nh; ++y)
		for (int x = 0; x < nw; ++x)
			for (int ch = 0; ch < 4; ++ch) {
				const unsigned char *p = &pic[(size_t) y * sy * row + (size_t) x * sx * 4 + ch];
				unsigned v;
				if (sx == 2 && sy == 2)
					v = (p[0] + p[4] + p[row] + p[row + 4]) >> 2;
				else if (sx == 2)
					v = (p[0] + p[4]) >> 1;
				else
					v = (p[0] + p[row]) >> 1;
				out[((size_t) y * nw + x) * 4 + ch] = (unsigned char) v;
			}
	pic.swap(out);
	w = nw;
	h = nh;
}

static void put32(unsigned char *p, uint32_t v)
{
	p[0] = v & 0xFF; p[1] = (v >> 8) & 0xFF; p[2] = (v >> 16) & 0xFF; p[3] = v >> 24;
}

int main(int argc, char **argv)
{
	const char *infile = nullptr, *outfile = nullptr, *library = nullptr;
	unsigned int format = S2TC_B200_GL_RGBA_DXT1;
	int opt;
	while ((opt = getopt(argc, argv, "i:o:t:l:")) != -1) {
		switch (opt) {
		case 'i': infile = optarg; break;
		case 'o': outfile = optarg; break;
		case 'l': library = optarg; break;
		case 't':
			if (!strcasecmp(optarg, "DXT1")) format = S2TC_B200_GL_RGBA_DXT1;
			else if (!strcasecmp(optarg, "DXT3")) format = S2TC_B200_GL_RGBA_DXT3;
			else if (!strcasecmp(optarg, "DXT5")) format = S2TC_B200_GL_RGBA_DXT5;
			else return usage(argv[0]);
			break;
		default: return usage(argv[0]);
		}
	}
	compress_fn compress = tx_compress_dxtn;
	if (library) {
		void *l = dlopen(library, RTLD_NOW);
		if (!l) {
			fprintf(stderr, "Cannot load library: %s\n", dlerror());
			return 1;
		}
		compress = (compress_fn) dlsym(l, "tx_compress_dxtn");
		if (!compress) {
			fprintf(stderr, "The selected libtxc_dxtn.so does not contain all required symbols.");
			return 1;
		}
	}
	FILE *out = outfile ? fopen(outfile, "wb") : stdout;
	if (!out) {
		printf("opening output failed\n");
		return 2;
	}
	FILE *in = infile ? fopen(infile, "rb") : stdin;
	std::vector<unsigned char> file;
	if (!in || !read_all(in, file)) {
		printf("FS_LoadFile failed\n");
		return 2;
	}
	int w, h;
	std::vector<unsigned char> pic;
	if (!load_tga(file, w, h, pic)) {
		printf("LoadTGA failed\n");
		return 2;
	}
	int mips = 0;
	while (w >= (1 << mips) || h >= (1 << mips))
		++mips;
	const int bs = format == S2TC_B200_GL_RGBA_DXT1 ? 8 : 16;
	bool alpha = false;
	for (size_t i = 3; i < pic.size(); i += 4)
		alpha |= pic[i] != 255;

	unsigned char hdr[128] = {0}; // DDS header as the reference lays it out (s2tc_compress.c:662-720)
	memcpy(hdr, "DDS ", 4);
	put32(hdr + 4, 124);
	put32(hdr + 8, 0x000A1007);
	put32(hdr + 12, (uint32_t) h);
	put32(hdr + 16, (uint32_t) w);
	put32(hdr + 20, (uint32_t) (((w + 3) / 4) * ((h + 3) / 4) * bs));
	put32(hdr + 28, (uint32_t) mips);
	put32(hdr + 76, 32);
	put32(hdr + 80, alpha ? 5 : 4);
	memcpy(hdr + 84, format == S2TC_B200_GL_RGBA_DXT1 ? "DXT1" : (format == S2TC_B200_GL_RGBA_DXT3 ? "DXT3" : "DXT5"), 4);
	put32(hdr + 108, 0x00401008);
	fwrite(hdr, 1, sizeof(hdr), out);

	if (!library) {
		// our own library: the whole chain stays on the GPU (one upload, every level encoded and halved there,
		// one download) -- same bytes as the per-level loop below
		s2tc_b200_settings st;
		s2tc_b200_settings_from_env(format == S2TC_B200_GL_RGBA_DXT1 ? S2TC_B200_DXT1 : (format == S2TC_B200_GL_RGBA_DXT3 ? S2TC_B200_DXT3 : S2TC_B200_DXT5), &st);
		s2tc_b200_ctx *ctx = s2tc_b200_default_ctx();
		std::vector<unsigned char> obuf(s2tc_b200_mipchain_bytes(st.dxt, w, h));
		uint64_t cursor = s2tc_b200_rand_cursor_get();
		if (!ctx || s2tc_b200_compress_mipchain_host(ctx, &st, w, h, pic.data(), obuf.data(), &cursor) != 0) {
			fprintf(stderr, "s2tc_compress: %s\n", s2tc_b200_last_error());
			return 3;
		}
		s2tc_b200_rand_cursor_set(cursor);
		fwrite(obuf.data(), 1, obuf.size(), out);
		if (outfile)
			fclose(out);
		return 0;
	}
	for (;;) {
		const int bw = (w + 3) / 4, bh = (h + 3) / 4;
		std::vector<unsigned char> obuf((size_t) bs * bw * bh);
		compress(4, w, h, pic.data(), format, obuf.data(), bw * bs);
		fwrite(obuf.data(), 1, obuf.size(), out);
		if (w == 1 && h == 1)
			break;
		mip_reduce(pic, w, h);
	}
	if (outfile)
		fclose(out);
	return 0;
}
