// hostsim.cpp -- TEST-ONLY CPU build of the encoder's host+device headers (s2tc_b200/csrc/*.cuh).
//
// The per-block logic, the metric arithmetic, the rand() jump-ahead, the dither transfer maps and the
// transcoder are written as __host__ __device__ code; this file instantiates them for the CPU and
// walks an image the way the kernels do (same chunking, same per-thread rand segments), so that the
// exact source the GPU runs can be checked against the oracle in the CPU-only test tier.
// It is never linked into, loaded by, or reachable from the shipped library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../s2tc_b200/csrc/block_core.cuh"
#include "../../s2tc_b200/csrc/dither_core.cuh"
#include "../../s2tc_b200/csrc/glibc_rand.cuh"
#include "../../s2tc_b200/csrc/transcode_core.cuh"

using namespace s2tc;

namespace {

constexpr int kChunk = 128, kTileChunks = 128;      // mirrors kernels_misc.cu
constexpr int kBlocksPerRandThread = 8;               // mirrors api.cu

int fs_width = 0; // image width for the Floyd-Steinberg walk (the other modes only need the texel count)

void prepass(const uint8_t *src, int comps, int abits, int dither, size_t npix, uint32_t *out,
		int *carry_io = nullptr, ByteMap *summary = nullptr)
{
	if (dither == kDitherNone) {
		for (size_t i = 0; i < npix; ++i) {
			const uint8_t *p = src + i * comps;
			out[i] = reduce_none(p[0], p[1], p[2], comps == 4 ? p[3] : 0, abits, comps == 4);
		}
		return;
	}
	if (dither == kDitherFloyd) { // sequential walk with the kernel's texel arithmetic and its alpha-seed rule
		const int w = fs_width, h = (int) (npix / fs_width);
		std::vector<int> cur(3 * (w + 2)), nxt(3 * (w + 2)), seed(w + 2, 0);
		for (int y = 0; y < h; ++y) {
			std::fill(nxt.begin(), nxt.end(), 0);
			for (int x = 0; x < w; ++x) {
				const uint8_t *p = src + ((size_t) y * w + x) * comps;
				uint32_t word = comps == 4 ? ((uint32_t) p[3] << 24) : (((1u << abits) - 1u) << 24);
				for (int c = 0; c < 3; ++c) {
					int *tr = cur.data() + c * (w + 2), *dr = nxt.data() + c * (w + 2);
					if (c == 0 && y == h - 1 && (h & 1))
						seed[x + 1] = tr[x + 1];
					const FloydOut o = c == 1 ? floyd_texel<2>(p[c], tr[x + 1]) : floyd_texel<3>(p[c], tr[x + 1]);
					tr[x + 2] += o.e7; dr[x] += o.e3; dr[x + 1] += o.e5; dr[x + 2] += o.e1;
					word |= (uint32_t) o.q << (8 * c);
				}
				out[(size_t) y * w + x] = word;
			}
			if (y == h - 1 && !(h & 1))
				for (int x = 0; x < w + 2; ++x)
					seed[x] = nxt[x];
			cur.swap(nxt);
		}
		if (comps == 4 && abits != 8) {
			std::vector<int> ca(seed), na(w + 2);
			for (int y = 0; y < h; ++y) {
				std::fill(na.begin(), na.end(), 0);
				for (int x = 0; x < w; ++x) {
					const int a = src[((size_t) y * w + x) * 4 + 3];
					const FloydOut o = abits == 1 ? floyd_texel<7>(a, ca[x + 1]) : floyd_texel<4>(a, ca[x + 1]);
					ca[x + 2] += o.e7; na[x] += o.e3; na[x + 1] += o.e5; na[x + 2] += o.e1;
					out[(size_t) y * w + x] = (out[(size_t) y * w + x] & 0x00FFFFFFu) | ((uint32_t) o.q << 24);
				}
				ca.swap(na);
			}
		}
		return;
	}
	// DITHER_SIMPLE through the three phases of kernels_misc.cu (same chunking, same map arithmetic)
	const ChanKinds kinds = chan_kinds(comps, abits);
	static DitherLut lut;
	static bool lut_ready = false;
	if (!lut_ready) {
		build_dither_lut(lut);
		lut_ready = true;
	}
	std::vector<uint32_t> wide(npix);
	for (size_t i = 0; i < npix; ++i) {
		const uint8_t *p = src + i * comps;
		wide[i] = (uint32_t) p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) (comps == 4 ? p[3] : 0) << 24);
	}
	const size_t nchunks = (npix + kChunk - 1) / kChunk;
	const size_t ntiles = (nchunks + kTileChunks - 1) / kTileChunks;
	// chunkmap[chunk][0..2]: EXCLUSIVE prefix maps of the colour channels inside the tile (composition of the tile's
	// earlier chunks), [3]: alpha -- prefix sum mod 255 (1-bit kind) or the chunk's own 31-state map (DXT3)
	std::vector<ByteMap> chunkmap(ntiles * kTileChunks * 4), tilemap(ntiles * 4);
	std::vector<RgbTables> tabs(ntiles * kTileChunks);
	std::vector<uint32_t> asums(ntiles * kTileChunks);
	std::vector<ByteMap> amaps(ntiles * kTileChunks);
	for (size_t chunk = 0; chunk < ntiles * kTileChunks; ++chunk) { // phase 1a: right-to-left PRMT composition per chunk
		const size_t first = chunk * kChunk;
		const int count = first >= npix ? 0 : (int) std::min<size_t>(kChunk, npix - first);
		RgbTables tab;
		rgb_tables_init(tab);
		uint32_t asum = 0;
		for (int i = 0; i < count; ++i)
			asum += wide[first + i] >> 24;
		if (count == kChunk) { // full chunks: two halves built independently and joined, as the kernel does
			RgbTables left;
			rgb_tables_init(left);
			for (int i = kChunk / 2; i > 0; --i) {
				rgb_tables_prepend(left, wide[first + i - 1], lut.lut3, lut.lut2);
				rgb_tables_prepend(tab, wide[first + kChunk / 2 + i - 1], lut.lut3, lut.lut2);
			}
			RgbTables whole;
			rgb_tables_join(whole, left, tab);
			tab = whole;
		} else {
			for (int i = count; i > 0; --i)
				rgb_tables_prepend(tab, wide[first + i - 1], lut.lut3, lut.lut2);
		}
		tabs[chunk] = tab;
		asums[chunk] = asum;
		memset(amaps[chunk].e, 0, sizeof(amaps[chunk].e));
		if (kinds.k[3] == kChanShift4)
			alpha_map_of_run(amaps[chunk], kChanShift4, (const uint8_t *) (wide.data() + (count ? first : 0)) + 3, 4, count);
	}
	for (size_t t = 0; t < ntiles; ++t) { // phase 1b: the kernel's scan over the tile's chunk tables
		RgbTables incl[kTileChunks];
		uint32_t aincl[kTileChunks];
		for (int k = 0; k < kTileChunks; ++k) {
			incl[k] = tabs[t * kTileChunks + k];
			aincl[k] = asums[t * kTileChunks + k];
		}
		for (int w = 0; w < kTileChunks / 32; ++w) // Kogge-Stone inside every group of 32 chunks (a warp)
			for (int delta = 1; delta < 32; delta <<= 1) {
				RgbTables next[32];
				uint32_t anext[32];
				for (int l = 0; l < 32; ++l) {
					next[l] = incl[w * 32 + l];
					anext[l] = aincl[w * 32 + l];
					if (l >= delta) {
						rgb_tables_join(next[l], incl[w * 32 + l - delta], incl[w * 32 + l]);
						anext[l] += aincl[w * 32 + l - delta];
					}
				}
				for (int l = 0; l < 32; ++l) {
					incl[w * 32 + l] = next[l];
					aincl[w * 32 + l] = anext[l];
				}
			}
		for (int k = 0; k < kTileChunks; ++k) {
			const int w = k / 32, l = k % 32;
			RgbTables excl, left;
			rgb_tables_init(left);
			uint32_t aexcl = aincl[k] - asums[t * kTileChunks + k];
			for (int v = 0; v < w; ++v) { // the warps to the left
				RgbTables both;
				rgb_tables_join(both, left, incl[v * 32 + 31]);
				left = both;
				aexcl += aincl[v * 32 + 31];
			}
			if (l == 0)
				rgb_tables_init(excl);
			else
				excl = incl[k - 1];
			if (w > 0) {
				RgbTables both;
				rgb_tables_join(both, left, excl);
				excl = both;
			}
			ByteMap *m = &chunkmap[(t * kTileChunks + k) * 4];
			rgb_tables_store(excl, m[0], m[1], m[2]);
			if (kinds.k[3] == kChanShift4)
				m[3] = amaps[t * kTileChunks + k];
			else {
				memset(m[3].e, 0, sizeof(m[3].e));
				m[3].e[0] = (uint8_t) (aexcl % 255u);
			}
			if (k == kTileChunks - 1) { // the last chunk's inclusive prefix is the tile's map
				RgbTables whole;
				rgb_tables_join(whole, left, incl[k]);
				ByteMap *tm = &tilemap[t * 4];
				rgb_tables_store(whole, tm[0], tm[1], tm[2]);
				memset(tm[3].e, 0, sizeof(tm[3].e));
				tm[3].e[0] = (uint8_t) ((aexcl + asums[t * kTileChunks + k]) % 255u);
			}
		}
		if (kinds.k[3] == kChanShift4) { // 31-state alpha maps: composed one by one
			ByteMap acc;
			bmap_identity(acc, kChanShift4);
			for (int k = 0; k < kTileChunks; ++k) {
				ByteMap r = acc;
				bmap_compose(r, acc, amaps[t * kTileChunks + k], kChanShift4);
				acc = r;
			}
			tilemap[t * 4 + 3] = acc;
		}
	}
	// phase 2: partial maps of 32 tiles each, walked from the carry; then the tiles of every part
	const size_t nparts = (ntiles + 31) / 32;
	std::vector<ByteMap> parts(nparts * 4);
	for (size_t p = 0; p < nparts; ++p)
		for (int ch = 0; ch < 4; ++ch) {
			ByteMap acc;
			bmap_identity(acc, kinds.k[ch]);
			for (size_t t = p * 32; t < std::min(ntiles, p * 32 + 32); ++t) {
				ByteMap r = acc;
				bmap_compose(r, acc, tilemap[t * 4 + ch], kinds.k[ch]);
				acc = r;
			}
			parts[p * 4 + ch] = acc;
		}
	if (summary) { // transfer function of the whole range = composition of the partial maps
		for (int ch = 0; ch < 4; ++ch) {
			bmap_identity(summary[ch], kinds.k[ch]);
			for (size_t p = 0; p < nparts; ++p) {
				ByteMap r = summary[ch];
				bmap_compose(r, summary[ch], parts[p * 4 + ch], kinds.k[ch]);
				summary[ch] = r;
			}
		}
		return;
	}
	int carry[4] = {0, 0, 0, 0};
	if (carry_io)
		memcpy(carry, carry_io, sizeof(carry));
	std::vector<int> tile_carry(ntiles * 4);
	for (size_t p = 0; p < nparts; ++p)
		for (int ch = 0; ch < 4; ++ch) {
			int c = carry[ch]; // carry entering the part
			for (size_t t = p * 32; t < std::min(ntiles, p * 32 + 32); ++t) {
				tile_carry[t * 4 + ch] = c;
				c = bmap_apply(tilemap[t * 4 + ch], kinds.k[ch], c);
			}
			carry[ch] = bmap_apply(parts[p * 4 + ch], kinds.k[ch], carry[ch]);
		}
	if (carry_io)
		memcpy(carry_io, carry, sizeof(carry));
	for (size_t t = 0; t < ntiles; ++t) { // phase 3: carry entering a chunk = its prefix map applied to the tile's carry
		int ac = tile_carry[t * 4 + 3]; // DXT3 alpha only: walked through the per-chunk maps
		for (int k = 0; k < kTileChunks; ++k) {
			const size_t chunk = t * kTileChunks + k, first = chunk * kChunk;
			const int count = first >= npix ? 0 : (int) std::min<size_t>(kChunk, npix - first);
			const ByteMap *m = &chunkmap[chunk * 4];
			int cc[4];
			for (int ch = 0; ch < 3; ++ch)
				cc[ch] = bmap_apply(m[ch], kinds.k[ch], tile_carry[t * 4 + ch]);
			if (kinds.k[3] == kChanShift4) {
				cc[3] = ac;
				ac = bmap_apply(m[3], kChanShift4, ac);
			} else
				cc[3] = bmap_apply(m[3], kinds.k[3], tile_carry[t * 4 + 3]);
			for (int i = 0; i < count; ++i)
				wide[first + i] = replay_texel(cc, wide[first + i], kinds.k[3], comps == 4, abits);
		}
	}
	memcpy(out, wide.data(), npix * 4);
}

void load(const uint32_t *img, int width, int height, int bx, int by, Block &b)
{
	const int w = std::min(4, width - bx * 4), h = std::min(4, height - by * 4);
	b.valid = valid_mask(w, h);
	for (int y = 0; y < 4; ++y)
		for (int x = 0; x < 4; ++x)
			b.px[y * 4 + x] = (y < h && x < w) ? img[(size_t) (by * 4 + y) * width + bx * 4 + x] : 0;
}

template <int DXT, int CD>
void encode_image(const uint32_t *img, int width, int height, int nrandom, int refine, uint64_t cursor, uint8_t *dest)
{
	const int bw = (width + 3) / 4, bh = (height + 3) / 4;
	const bool fast = is_fast_mode(CD, nrandom);
	const int nr = nrandom > 0 ? nrandom : 0;
	RandPlan plan;
	GlibcRand rng;
	if (nr)
		rand_plan_init(plan, cursor, (uint64_t) kBlocksPerRandThread * draws_per_block(DXT, nr));
	std::vector<uint32_t> c(16 + nr + 2);
	std::vector<uint8_t> ca(16 + nr + 2);
	std::vector<int> d((size_t) (16 + nr + 3) * 16);
	for (int blk = 0; blk < bw * bh; ++blk) {
		Block b;
		load(img, width, height, blk % bw, blk / bw, b);
		uint32_t c0, c1;
		int a0 = 0, a1 = 0;
		if (fast) {
			fast_candidates<DXT, CD>(b, c0, c1, a0, a1);
		} else {
			int n = gather_colors<DXT>(b, c.data(), ca.data());
			int m = n;
			if (nr) {
				if (blk % kBlocksPerRandThread == 0) // a new "thread" seeks its own window
					rand_plan_seek(plan, (uint32_t) (blk / kBlocksPerRandThread), rng);
				const CandBox box = candidate_box(c.data(), ca.data(), n);
				for (int k = 0; k < nr; ++k) {
					const uint32_t p = draw_candidate<DXT>(box, rng);
					c[n + k] = px_rgb(p);
					ca[n + k] = (uint8_t) (p >> 24);
				}
				m = n + nr;
			} else if (n == 1) {
				c[1] = c[0];
				ca[1] = ca[0];
				m = n = 2;
			}
			search_colors_scalar<CD>(c.data(), n, m, d.data());
			if (DXT == kDxt5)
				search_alpha_scalar(ca.data(), n, m, d.data());
			c0 = c[0];
			c1 = c[1];
			a0 = ca[0];
			a1 = ca[1];
		}
		uint32_t w[4];
		finish_block<DXT, CD>(b, refine, c0, c1, a0, a1, w);
		memcpy(dest + (size_t) blk * block_bytes(DXT), w, block_bytes(DXT));
	}
}

template <int DXT>
void encode_cd(int cd, const uint32_t *img, int width, int height, int nrandom, int refine, uint64_t cursor, uint8_t *dest)
{
	switch (cd) {
	case kRGB: encode_image<DXT, kRGB>(img, width, height, nrandom, refine, cursor, dest); break;
	case kYUV: encode_image<DXT, kYUV>(img, width, height, nrandom, refine, cursor, dest); break;
	case kSRGB: encode_image<DXT, kSRGB>(img, width, height, nrandom, refine, cursor, dest); break;
	case kSRGB_MIXED: encode_image<DXT, kSRGB_MIXED>(img, width, height, nrandom, refine, cursor, dest); break;
	case kAVG: encode_image<DXT, kAVG>(img, width, height, nrandom, refine, cursor, dest); break;
	case kW0AVG: encode_image<DXT, kW0AVG>(img, width, height, nrandom, refine, cursor, dest); break;
	case kNORMALMAP: encode_image<DXT, kNORMALMAP>(img, width, height, nrandom, refine, cursor, dest); break;
	default: encode_image<DXT, kWAVG>(img, width, height, nrandom, refine, cursor, dest); break;
	}
}

} // namespace

extern "C" {

// tight output; returns 0
int hostsim_compress(int srccomps, int width, int height, const uint8_t *src, int dxt, int cd, int nrandom, int refine,
		int dither, uint64_t cursor, uint8_t *dest)
{
	fs_width = width;
	const int comps = srccomps == 3 ? 3 : 4;
	dxt = norm_dxt(dxt);
	cd = norm_cd(cd);
	refine = norm_refine(refine);
	std::vector<uint32_t> img((size_t) width * height);
	prepass(src, comps, alpha_bits(dxt), dither, img.size(), img.data());
	switch (dxt) {
	case kDxt1: encode_cd<kDxt1>(cd, img.data(), width, height, nrandom, refine, cursor, dest); break;
	case kDxt3: encode_cd<kDxt3>(cd, img.data(), width, height, nrandom, refine, cursor, dest); break;
	default: encode_cd<kDxt5>(cd, img.data(), width, height, nrandom, refine, cursor, dest); break;
	}
	return 0;
}

int hostsim_prepass(int srccomps, int abits, int dither, size_t npix, const uint8_t *src, uint8_t *out)
{
	if (dither == kDitherFloyd)
		return -1; // needs the image width: go through hostsim_compress
	prepass(src, srccomps == 3 ? 3 : 4, abits, dither, npix, (uint32_t *) out);
	return 0;
}

// DITHER_SIMPLE over a texel range with an explicit carry in/out (what one shard of a sharded image runs)
void hostsim_prepass_range(int srccomps, int abits, size_t npix, const uint8_t *src, int *carry_io, uint8_t *out)
{
	prepass(src, srccomps == 3 ? 3 : 4, abits, kDitherSimple, npix, (uint32_t *) out, carry_io);
}

// the range's transfer function: 4 channels x 32 bytes, the layout s2tc_b200_dither_summary_device returns
void hostsim_dither_summary(int srccomps, int abits, size_t npix, const uint8_t *src, uint64_t *maps)
{
	ByteMap m[4];
	prepass(src, srccomps == 3 ? 3 : 4, abits, kDitherSimple, npix, nullptr, nullptr, m);
	for (int ch = 0; ch < 4; ++ch)
		memcpy(maps + 4 * ch, m[ch].e, sizeof(m[ch].e));
}

void hostsim_transcode(int dxt, uint8_t *blocks, size_t nblocks)
{
	for (size_t i = 0; i < nblocks; ++i) {
		if (dxt == kDxt1) {
			uint32_t w[2];
			memcpy(w, blocks + i * 8, 8);
			transcode_color_dxt1(w[0], w[1]);
			memcpy(blocks + i * 8, w, 8);
		} else {
			uint32_t w[4];
			memcpy(w, blocks + i * 16, 16);
			transcode_color_opaque(w[2], w[3]);
			if (dxt == kDxt5) {
				uint64_t a = transcode_alpha_dxt5((uint64_t) w[0] | ((uint64_t) w[1] << 32));
				w[0] = (uint32_t) a;
				w[1] = (uint32_t) (a >> 32);
			}
			memcpy(blocks + i * 16, w, 16);
		}
	}
}

// exhaustive check of the reciprocal-multiply division used for the cluster means
int hostsim_check_division(void)
{
	for (int n = 1; n <= 16; ++n)
		for (int x = 0; x < 8192; ++x)
			if (div_by_2n(x, half_recip18(n)) != x / (2 * n))
				return 1000 * n + 1;
	return 0;
}

int hostsim_rand(uint64_t cursor, uint64_t stride, uint32_t t, int n, int *out)
{
	RandPlan plan;
	GlibcRand g;
	rand_plan_init(plan, cursor, stride);
	rand_plan_seek(plan, t, g);
	for (int i = 0; i < n; ++i)
		out[i] = g.next();
	return 0;
}

int hostsim_color_dist(int cd, uint32_t a, uint32_t b)
{
	switch (cd) {
	case kRGB: return color_dist<kRGB>(a, b);
	case kYUV: return color_dist<kYUV>(a, b);
	case kSRGB: return color_dist<kSRGB>(a, b);
	case kSRGB_MIXED: return color_dist<kSRGB_MIXED>(a, b);
	case kAVG: return color_dist<kAVG>(a, b);
	case kW0AVG: return color_dist<kW0AVG>(a, b);
	case kNORMALMAP: return color_dist<kNORMALMAP>(a, b);
	default: return color_dist<kWAVG>(a, b);
	}
}

} // extern "C"
