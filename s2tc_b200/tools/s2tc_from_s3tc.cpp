// s2tc_from_s3tc -- rewrites an S3TC (DXT1/DXT3/DXT5) DDS file into its S2TC subset on the GPU.
//
// Command-line surface of the reference tool (s2tc_from_s3tc.cpp:192-271): -i infile.dds (default stdin),
// -o outfile.dds (default stdout).  The 128-byte header is copied verbatim, every complete block after
// it is transcoded by the CUDA kernel behind s2tc_b200_transcode_host; a trailing partial block is
// dropped, as the reference's fread loop does.
#include <getopt.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/s2tc_b200.h"

int main(int argc, char **argv)
{
	const char *infile = nullptr, *outfile = nullptr;
	int opt;
	while ((opt = getopt(argc, argv, "i:o:")) != -1) {
		switch (opt) {
		case 'i': infile = optarg; break;
		case 'o': outfile = optarg; break;
		default:
			fprintf(stderr, "usage:\n%s \n    [-i infile.dds]\n    [-o outfile.dds]\n", argv[0]);
			return 1;
		}
	}
	FILE *in = infile ? fopen(infile, "rb") : stdin;
	if (!in) {
		printf("opening input failed\n");
		return 2;
	}
	FILE *out = outfile ? fopen(outfile, "wb") : stdout;
	if (!out) {
		printf("opening output failed\n");
		return 2;
	}
	std::vector<unsigned char> data;
	unsigned char chunk[65536];
	size_t n;
	while ((n = fread(chunk, 1, sizeof(chunk), in)) > 0)
		data.insert(data.end(), chunk, chunk + n);
	if (data.size() < 128) {
		fprintf(stderr, "Only DXT1, DXT3, DXT5 are supported!\n");
		return 1;
	}
	const uint32_t fourcc = data[84] | data[85] << 8 | data[86] << 16 | (uint32_t) data[87] << 24;
	int dxt, bs;
	switch (fourcc) {
	case 0x31545844: dxt = S2TC_B200_DXT1; bs = 8; break;
	case 0x33545844: dxt = S2TC_B200_DXT3; bs = 16; break;
	case 0x35545844: dxt = S2TC_B200_DXT5; bs = 16; break;
	default:
		fprintf(stderr, "Only DXT1, DXT3, DXT5 are supported!\n");
		return 1;
	}
	const size_t nblocks = (data.size() - 128) / bs;
	if (nblocks) {
		s2tc_b200_ctx *ctx = s2tc_b200_default_ctx();
		if (!ctx || s2tc_b200_transcode_host(ctx, dxt, data.data() + 128, nblocks) != 0) {
			fprintf(stderr, "s2tc_from_s3tc: %s\n", s2tc_b200_last_error());
			return 3;
		}
	}
	fwrite(data.data(), 1, 128 + nblocks * bs, out);
	if (outfile)
		fclose(out);
	return 0;
}
