// kernels_misc.cu -- the kernels either side of the search: the 565 pre-pass (DITHER_NONE as a
// stand-alone pass, DITHER_SIMPLE as a three-phase carry scan), the random-candidate generator that
// replays glibc's rand() stream in parallel, and the S3TC -> S2TC transcoder.
#include "kernels.cuh"
#include "transcode_core.cuh"

namespace s2tc {

// =====================================================================================================
// DITHER_NONE as its own pass (reference rgb565_image, s2tc_algorithm.cpp:1269-1306).  The encode
// kernels normally fuse these shifts into their loads; this pass backs the exported rgb565_image().
// =====================================================================================================
__global__ void prepass_none_kernel(const uint8_t *__restrict__ src, int srccomps, int alphabits, size_t npixels,
		uint32_t *__restrict__ out)
{
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < npixels; i += stride) {
		if (srccomps == 4) {
			out[i] = reduce_word(__ldg(reinterpret_cast<const uint32_t *>(src) + i), alphabits);
		} else {
			const uint8_t *p = src + i * 3;
			out[i] = reduce_none(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0, alphabits, false);
		}
	}
}

cudaError_t launch_prepass_none(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		cudaStream_t stream)
{
	if (!npixels)
		return cudaSuccess;
	const int threads = 256;
	size_t blocks = (npixels + threads - 1) / threads;
	if (blocks > 148 * 32)
		blocks = 148 * 32;
	prepass_none_kernel<<<(unsigned) blocks, threads, 0, stream>>>((const uint8_t *) d_src, srccomps, alphabits,
			npixels, (uint32_t *) d_reduced);
	return cudaGetLastError();
}

// =====================================================================================================
// DITHER_SIMPLE (reference s2tc_algorithm.cpp:1307-1349): one carry per channel runs through the whole
// image in raster order.  Texels are cut into chunks of kChunk (one thread each), 128 chunks per CTA tile.
//   phase 1  dither_maps_kernel : per chunk, the transfer map carry-in -> carry-out of every channel, built right to
//                                 left with byte permutes (dither_core.cuh); then a scan over the tile's 128 chunk
//                                 maps: every chunk's EXCLUSIVE prefix map is stored, and the tile's map
//   phase 2  scan_*_kernel      : compose / walk the tile maps from the true carry-in (three small launches) and leave
//                                 every tile's starting carry (or, for sharding, the map of the whole range)
//   phase 3  dither_apply_kernel: carry entering a chunk = its prefix map applied to the tile's carry; the texels go
//                                 through shared memory per warp (coalesced global access) and are replayed in place
// =====================================================================================================
constexpr int kChunk = 128;       // texels per thread
constexpr int kTileThreads = 128; // chunks per CTA
constexpr int kTilePixels = kChunk * kTileThreads;

static_assert(kTileThreads == 128, "one warp per channel walks the chunk carries");

// 5 KB of selectors, global -> shared: 128-bit loads, all of a thread's loads in flight together (as a word loop this
// prologue was a quarter of the maps kernel's stall samples: ten dependent round trips per CTA)
__device__ __forceinline__ void load_dither_lut(const DitherLut *__restrict__ lut, uint32_t (*s_lut3)[4], uint32_t *s_lut2, int t)
{
	static_assert(sizeof(DitherLut) == (256 * 4 + 256) * 4, "lut3 then lut2, contiguous");
	const uint4 *g = reinterpret_cast<const uint4 *>(lut);
	uint4 v[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const int i = t + k * kTileThreads;
		v[k] = i < 320 ? __ldg(g + i) : make_uint4(0, 0, 0, 0);
	}
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const int i = t + k * kTileThreads;
		if (i < 256)
			reinterpret_cast<uint4 *>(s_lut3)[i] = v[k];
		else if (i < 320)
			reinterpret_cast<uint4 *>(s_lut2)[i - 256] = v[k];
	}
}

// Walks one channel through the tile's 128 chunk maps: s_carry[i * 4 + ch] = carry entering chunk i; returns the carry
// after the last chunk.  Called by ONE lane per channel, each in a different warp: the kinds are different code
// paths, and a warp that ran them in four of its lanes serialised them (the walk was a third of the apply kernel's
// time, profiles/r01h).  Shift kinds stay in table-index form, the 1-bit kind in biased-residue form, so a step is
// one dependent shared-memory load.
__device__ __forceinline__ int walk_chunk_carries(const ByteMap *s_maps, int *s_carry, int ch, int kind, int c)
{
	if (kind <= kChanShift4) {
		const int r = chan_radius(kind);
		uint32_t idx = (uint32_t) (c + r);
#pragma unroll 4
		for (int i = 0; i < kTileThreads; ++i) {
			s_carry[i * 4 + ch] = (int) idx - r;
			idx = s_maps[i * 4 + ch].e[idx];
		}
		return (int) idx - r;
	}
	if (kind == kChanBit1) { // balanced255(c + sum): keep u = c + 127 in [0, 254]
		uint32_t u = (uint32_t) (c + 127);
#pragma unroll 4
		for (int i = 0; i < kTileThreads; ++i) {
			s_carry[i * 4 + ch] = (int) u - 127;
			u += s_maps[i * 4 + ch].e[0];
			u = u >= 255u ? u - 255u : u;
		}
		return (int) u - 127;
	}
	for (int i = 0; i < kTileThreads; ++i)
		s_carry[i * 4 + ch] = 0;
	return 0;
}

__device__ __forceinline__ RgbTables shfl_up_tables(const RgbTables &t, int delta)
{
	RgbTables o;
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		o.r[g] = __shfl_up_sync(0xFFFFFFFFu, t.r[g], delta);
		o.b[g] = __shfl_up_sync(0xFFFFFFFFu, t.b[g], delta);
	}
	o.g[0] = __shfl_up_sync(0xFFFFFFFFu, t.g[0], delta);
	o.g[1] = __shfl_up_sync(0xFFFFFFFFu, t.g[1], delta);
	return o;
}

// the tables of a run as the three ByteMaps the scan and apply kernels read (little-endian words = table bytes)
__device__ __forceinline__ void store_rgb_maps(const RgbTables &t, ByteMap *dst /* [>= 3] */)
{
	uint4 *d = reinterpret_cast<uint4 *>(dst);
	const uint4 z = make_uint4(0, 0, 0, 0);
	d[0] = make_uint4(t.r[0], t.r[1], t.r[2], t.r[3]);
	d[1] = z;
	d[2] = make_uint4(t.g[0], t.g[1], 0, 0);
	d[3] = z;
	d[4] = make_uint4(t.b[0], t.b[1], t.b[2], t.b[3]);
	d[5] = z;
}

#ifndef S2TC_MAPS_MINBLOCKS
#define S2TC_MAPS_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(kTileThreads, S2TC_MAPS_MINBLOCKS)
dither_maps_kernel(const uint8_t *__restrict__ src, int srccomps, ChanKinds kinds, size_t npixels,
		const DitherLut *__restrict__ lut, uint32_t k16 /* 65536, opaque: see upper_half */,
		ByteMap *__restrict__ chunkmaps /* [chunks][4] */, ByteMap *__restrict__ tilemaps /* [tiles][4] */)
{
	__shared__ __align__(16) uint32_t s_lut3[256][4];
	__shared__ __align__(16) uint32_t s_lut2[256];
	__shared__ ByteMap s_amap[kTileThreads]; // DXT3 alpha maps only
	const int t = threadIdx.x;
	load_dither_lut(lut, s_lut3, s_lut2, t);
	__syncthreads();

	const size_t first = ((size_t) blockIdx.x * kTileThreads + t) * kChunk;
	const int count = first >= npixels ? 0 : (int) min((size_t) kChunk, npixels - first);
	RgbTables tab;
	rgb_tables_init(tab);
	uint32_t asum = 0;
	if (srccomps == 4 && count == kChunk && ((size_t) src & 15) == 0) {
		// two half-chunks advance together (two independent dependency chains per channel), joined at the end
		const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(src) + first);
		RgbTables left;
		rgb_tables_init(left);
		uint4 nl = __ldg(p + kChunk / 8 - 1), nr = __ldg(p + kChunk / 4 - 1);
#pragma unroll 2
		for (int i = kChunk / 8 - 1; i >= 0; --i) { // right to left, four texels per 128-bit load, the next pair in flight
			const uint4 ql = nl, qr = nr;
			if (i > 0) {
				nl = __ldg(p + i - 1);
				nr = __ldg(p + kChunk / 8 + i - 1);
			}
			rgb_tables_prepend(left, ql.w, s_lut3, s_lut2, k16);
			rgb_tables_prepend(tab, qr.w, s_lut3, s_lut2, k16);
			rgb_tables_prepend(left, ql.z, s_lut3, s_lut2, k16);
			rgb_tables_prepend(tab, qr.z, s_lut3, s_lut2, k16);
			rgb_tables_prepend(left, ql.y, s_lut3, s_lut2, k16);
			rgb_tables_prepend(tab, qr.y, s_lut3, s_lut2, k16);
			rgb_tables_prepend(left, ql.x, s_lut3, s_lut2, k16);
			rgb_tables_prepend(tab, qr.x, s_lut3, s_lut2, k16);
			asum += (ql.x >> 24) + (ql.y >> 24) + (ql.z >> 24) + (ql.w >> 24) + (qr.x >> 24) + (qr.y >> 24) + (qr.z >> 24) + (qr.w >> 24);
		}
		RgbTables whole;
		rgb_tables_join(whole, left, tab);
		tab = whole;
	} else if (srccomps == 4) {
		const uint32_t *p = reinterpret_cast<const uint32_t *>(src) + first;
		for (int i = count; i > 0; --i) {
			const uint32_t w = __ldg(p + i - 1);
			rgb_tables_prepend(tab, w, s_lut3, s_lut2, k16);
			asum += w >> 24;
		}
	} else {
		const uint8_t *p = src + first * 3;
		for (int i = count; i > 0; --i) {
			const uint8_t *q = p + (size_t) (i - 1) * 3;
			rgb_tables_prepend(tab, (uint32_t) __ldg(q) | ((uint32_t) __ldg(q + 1) << 8) | ((uint32_t) __ldg(q + 2) << 16),
					s_lut3, s_lut2, k16);
		}
	}
	// ---- prefix maps out; tile map = composition of the 128 chunk maps ---------------------------------------------
	// What the replay needs for chunk t is the carry entering it: the composition of chunks 0 .. t-1 of the tile applied
	// to the tile's carry.  So the tile is scanned here: an inclusive scan over the warp's 32 chunk tables through
	// shuffles (5 levels; a join is ~60 byte permutes, tables stay in registers), the four warp totals meet in shared
	// memory, and every thread stores the EXCLUSIVE prefix of its chunk.  (First version: per-chunk maps were stored
	// and the replay kernel walked all 128 of them serially per channel before it could start.)
	const size_t chunk = (size_t) blockIdx.x * kTileThreads + t;
	const int lane = t & 31, warp = t >> 5;
	const uint32_t asum_own = asum;
#pragma unroll
	for (int delta = 1; delta < 32; delta <<= 1) {
		const RgbTables before = shfl_up_tables(tab, delta);
		const uint32_t asum_before = __shfl_up_sync(0xFFFFFFFFu, asum, delta);
		RgbTables both;
		rgb_tables_join(both, before, tab); // the run to my left first, then mine
		if (lane >= delta) {
			tab = both;
			asum += asum_before;
		}
	}
	__shared__ RgbTables s_warp[kTileThreads / 32];
	__shared__ uint32_t s_asum[kTileThreads / 32];
	if (lane == 31) {
		s_warp[warp] = tab;
		s_asum[warp] = asum;
	}
	ByteMap ma; // alpha of THIS chunk: explicit 31-state map (DXT3) -- those the replay kernel still walks
	if (kinds.k[3] == kChanShift4) {
		alpha_map_of_run(ma, kChanShift4, src + first * 4 + 3, 4, count); // DXT3 only: 31 explicit trajectories
		s_amap[t] = ma;
	}
	__syncthreads();
	// exclusive prefix inside the warp, then the warps to the left in front of it
	RgbTables excl = shfl_up_tables(tab, 1);
	if (lane == 0)
		rgb_tables_init(excl);
	uint32_t asum_excl = asum - asum_own;
	RgbTables left; // chunks of warps 0 .. warp-1
	rgb_tables_init(left);
	for (int w = 0; w < warp; ++w) {
		RgbTables both;
		rgb_tables_join(both, left, s_warp[w]);
		left = both;
		asum_excl += s_asum[w];
	}
	if (warp > 0) {
		RgbTables both;
		rgb_tables_join(both, left, excl);
		excl = both;
	}
	store_rgb_maps(excl, chunkmaps + chunk * 4);
	if (kinds.k[3] == kChanShift4)
		chunkmaps[chunk * 4 + 3] = ma;
	else {
		ByteMap m3;
#pragma unroll
		for (int k = 0; k < 32; ++k)
			m3.e[k] = 0;
		m3.e[0] = (uint8_t) (asum_excl % 255u); // diffuse1: sum of the sources in front of this chunk, mod 255
		chunkmaps[chunk * 4 + 3] = m3;
	}
	if (kinds.k[3] == kChanShift4) { // byte-wise tree for the 31-state alpha maps only
		for (int stride = 1; stride < kTileThreads; stride <<= 1) {
			if ((t & (2 * stride - 1)) == 0) {
				ByteMap r;
				bmap_compose(r, s_amap[t], s_amap[t + stride], kChanShift4);
				s_amap[t] = r;
			}
			__syncthreads();
		}
	}
	if (t == kTileThreads - 1) { // this thread's inclusive prefix is the whole tile
		RgbTables whole;
		rgb_tables_join(whole, left, tab);
		ByteMap *tm = tilemaps + (size_t) blockIdx.x * 4;
		store_rgb_maps(whole, tm);
		if (kinds.k[3] == kChanShift4)
			tm[3] = s_amap[0];
		else {
			ByteMap m3;
#pragma unroll
			for (int k = 0; k < 32; ++k)
				m3.e[k] = 0;
			m3.e[0] = (uint8_t) ((asum_excl + asum_own) % 255u);
			tm[3] = m3;
		}
	}
}

// ---- phase 2: scan over the tile maps, three small launches ----------------------------------------------------
// A warp holds a map with entry k in lane k, so composing two maps is one table lookup per lane
// (out[k] = second[first[k]]) and following a carry is the same lookup in one lane.
//   scan_partial_kernel : every warp composes the maps of its 32 tiles (grid: 256 tiles per CTA)
//   scan_carry_kernel   : one CTA, one warp per channel (the kinds are different code paths), walks the partial
//                         maps from the true carry-in and leaves the carry entering every warp's tiles -- or, for
//                         sharding (summary != nullptr), folds them into the map of the whole range
//   scan_walk_kernel    : every warp walks its 32 tiles from its carry and writes the carry entering each tile
// (The first version did all three in ONE CTA of 32 warps: 0.076 ms for 4096 tiles, issue-bound on a single SM,
// 17 % of the whole pre-pass by the time the other two phases had been tuned.)
constexpr int kScanWarpTiles = 32;
constexpr int kScanCtaWarps = 8;
constexpr int kScanCtaTiles = kScanWarpTiles * kScanCtaWarps;
constexpr int kCarryStage = 64; // partial maps the carry kernel stages in shared memory at a time

// this warp's tile maps -> its 4 KB of shared memory (coalesced 128-bit loads, zeros past the end)
__device__ __forceinline__ void scan_load_warp_tiles(const ByteMap *__restrict__ tilemaps, size_t tile0, size_t ntiles, int lane, uint4 *s_warp)
{
	const uint4 *g = reinterpret_cast<const uint4 *>(tilemaps + tile0 * 4); // 8 uint4 per tile (4 channels x 32 B)
#pragma unroll
	for (int q = 0; q < kScanWarpTiles * 8 / 32; ++q) {
		const int idx = q * 32 + lane;
		s_warp[idx] = tile0 + (size_t) (idx >> 3) < ntiles ? __ldg(g + idx) : make_uint4(0, 0, 0, 0);
	}
	__syncwarp();
}

// tiles_per_image > 0: the texel stream is a batch of images of that many tiles each and the carry restarts at every
// image (each mip level of each texture is its own tx_compress_dxtn call, ref s2tc_compress.c:722-733).  A restart is
// the constant map "everything -> carry 0", so it composes like any other map; only the 1-bit alpha channel, whose map
// is a sum, needs a flag (byte 1 of its partial map: the part contains a restart and byte 0 counts from there).
__global__ void __launch_bounds__(kScanCtaWarps * 32)
scan_partial_kernel(const ByteMap *__restrict__ tilemaps, size_t ntiles, ChanKinds kinds, uint8_t *__restrict__ parts /* [warps][4][32] */,
		unsigned tiles_per_image)
{
	__shared__ __align__(16) uint8_t s_tiles[kScanCtaWarps][kScanWarpTiles][4][32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const size_t gw = (size_t) blockIdx.x * kScanCtaWarps + warp;
	const size_t tile0 = gw * kScanWarpTiles;
	if (tile0 >= ntiles)
		return;
	scan_load_warp_tiles(tilemaps, tile0, ntiles, lane, reinterpret_cast<uint4 *>(&s_tiles[warp][0][0][0]));
	const int cnt = (int) min((size_t) kScanWarpTiles, ntiles - tile0);
	uint32_t acc[4];
#pragma unroll
	for (int ch = 0; ch < 4; ++ch)
		acc[ch] = kinds.k[ch] == kChanBit1 ? 0u : (uint32_t) lane;
	bool restarted = false;
	for (int i = 0; i < cnt; ++i) {
		const bool restart = tiles_per_image && (tile0 + i) % tiles_per_image == 0;
		restarted |= restart;
#pragma unroll
		for (int ch = 0; ch < 4; ++ch) {
			if (kinds.k[ch] <= kChanShift4) {
				if (restart)
					acc[ch] = (uint32_t) chan_radius(kinds.k[ch]);
				acc[ch] = s_tiles[warp][i][ch][acc[ch]]; // out[k] = tile[acc[k]]
			} else if (kinds.k[ch] == kChanBit1) {
				if (restart)
					acc[ch] = 0;
				acc[ch] += s_tiles[warp][i][ch][0];
				acc[ch] = acc[ch] >= 255u ? acc[ch] - 255u : acc[ch];
			}
		}
	}
#pragma unroll
	for (int ch = 0; ch < 4; ++ch)
		parts[(gw * 4 + ch) * 32 + lane] = kinds.k[ch] == kChanBit1 && lane == 1 ? (uint8_t) restarted : (uint8_t) acc[ch];
}

__global__ void __launch_bounds__(128)
scan_carry_kernel(const uint8_t *__restrict__ parts, size_t nparts, ChanKinds kinds, int *carry, int *__restrict__ starts /* [nparts][4] */,
		ByteMap *summary)
{
	__shared__ __align__(16) uint8_t s_parts[kCarryStage][4][32];
	const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5; // one warp per channel
	const int kind = kinds.k[ch];
	const int r = chan_radius(kind);
	// shift kinds: table index followed by this lane (summary: lane k follows state k; carry: every lane the carry);
	// 1-bit kind: residue of the running sum, biased by 127 in carry mode so that it stays in [0, 254]
	uint32_t v = kind <= kChanShift4 ? (summary ? (uint32_t) lane : (uint32_t) (carry[ch] + r))
			: (kind == kChanBit1 ? (summary ? 0u : (uint32_t) (carry[ch] + 127)) : 0u);
	for (size_t base = 0; base < nparts; base += kCarryStage) {
		const int cnt = (int) min((size_t) kCarryStage, nparts - base);
		__syncthreads();
		{
			const uint4 *g = reinterpret_cast<const uint4 *>(parts + base * 128);
			uint4 *sp = reinterpret_cast<uint4 *>(&s_parts[0][0][0]);
			for (int i = threadIdx.x; i < cnt * 8; i += 128)
				sp[i] = __ldg(g + i);
		}
		__syncthreads();
		if (kind <= kChanShift4) {
			for (int i = 0; i < cnt; ++i) {
				if (!summary && lane == 0)
					starts[(base + i) * 4 + ch] = (int) v - r;
				v = s_parts[i][ch][v & 31u];
			}
		} else if (kind == kChanBit1) {
			for (int i = 0; i < cnt; ++i) {
				if (!summary && lane == 0)
					starts[(base + i) * 4 + ch] = (int) v - 127;
				if (s_parts[i][ch][1]) // the part contains a restart: its sum counts from carry 0
					v = summary ? 0u : 127u;
				v += s_parts[i][ch][0];
				v = v >= 255u ? v - 255u : v;
			}
		} else if (!summary && lane == 0) {
			for (int i = 0; i < cnt; ++i)
				starts[(base + i) * 4 + ch] = 0;
		}
	}
	if (summary) {
		const int ns = chan_states(kind);
		summary[ch].e[lane] = (uint8_t) ((kind == kChanBit1 ? lane == 0 : lane < ns) ? v : 0u);
	} else if (lane == 0)
		carry[ch] = kind <= kChanShift4 ? (int) v - r : (kind == kChanBit1 ? (int) v - 127 : 0);
}

__global__ void __launch_bounds__(kScanCtaWarps * 32)
scan_walk_kernel(const ByteMap *__restrict__ tilemaps, size_t ntiles, ChanKinds kinds, const int *__restrict__ starts,
		int *__restrict__ tile_carry, unsigned tiles_per_image)
{
	__shared__ __align__(16) uint8_t s_tiles[kScanCtaWarps][kScanWarpTiles][4][32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const size_t gw = (size_t) blockIdx.x * kScanCtaWarps + warp;
	const size_t tile0 = gw * kScanWarpTiles;
	if (tile0 >= ntiles)
		return;
	scan_load_warp_tiles(tilemaps, tile0, ntiles, lane, reinterpret_cast<uint4 *>(&s_tiles[warp][0][0][0]));
	const int cnt = (int) min((size_t) kScanWarpTiles, ntiles - tile0);
	// lane ch (< 4) follows channel ch; the others idle (32 steps of one dependent lookup each)
	if (lane < 4) {
		const int ch = lane, kind = kinds.k[ch];
		const int r = chan_radius(kind);
		const int c0 = starts[gw * 4 + ch];
		uint32_t v = kind <= kChanShift4 ? (uint32_t) (c0 + r) : (uint32_t) (c0 + 127);
		for (int i = 0; i < cnt; ++i) {
			int c = 0;
			if (tiles_per_image && (tile0 + i) % tiles_per_image == 0)
				v = kind <= kChanShift4 ? (uint32_t) r : 127u; // a new image: carry 0
			if (kind <= kChanShift4) {
				c = (int) v - r;
				v = s_tiles[warp][i][ch][v & 31u];
			} else if (kind == kChanBit1) {
				c = (int) v - 127;
				v += s_tiles[warp][i][ch][0];
				v = v >= 255u ? v - 255u : v;
			}
			tile_carry[(tile0 + i) * 4 + ch] = c;
		}
	}
}

constexpr int kStagePitch = 36; // words per staging row: 32 texels + 4, so that rows 4 banks apart serve 128-bit accesses conflict-free
#ifndef S2TC_APPLY_MINBLOCKS
#define S2TC_APPLY_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(kTileThreads, S2TC_APPLY_MINBLOCKS)
dither_apply_kernel(const uint8_t *__restrict__ src, int srccomps, int alphabits, ChanKinds kinds, size_t npixels,
		const ByteMap *__restrict__ chunkmaps, const int *__restrict__ tile_carry, uint32_t *__restrict__ out)
{
	const int t = threadIdx.x;
	const size_t chunk = (size_t) blockIdx.x * kTileThreads + t;
	const size_t first = chunk * kChunk;
	const int count = first >= npixels ? 0 : (int) min((size_t) kChunk, npixels - first);
	const bool vec = srccomps == 4 && count == kChunk && (((size_t) src | (size_t) out) & 15) == 0;
	// Full warps move their texels through shared memory so that global accesses are coalesced: a thread's chunk is 512
	// contiguous bytes, so thread-private 128-bit accesses touch 32 different lines per warp instruction.  Instead the
	// warp copies one 128-byte quarter of each of its 32 chunks per step (8 instructions, 4 whole lines each) into a
	// staging row per thread (36-word pitch: conflict-free both ways), replays in place and copies back.
	const int lane = t & 31, warp = t >> 5;
	const bool staged = __all_sync(0xFFFFFFFFu, vec);
	__shared__ __align__(16) uint32_t s_stage[kTileThreads * kStagePitch];
	uint32_t *ws = s_stage + warp * 32 * kStagePitch;
	const size_t warp_chunk0 = (size_t) blockIdx.x * kTileThreads + warp * 32;
	const uint4 *gin = reinterpret_cast<const uint4 *>(src) + warp_chunk0 * (kChunk / 4);
	const int cl = lane >> 3, piece = lane & 7; // this lane moves piece `piece` of chunks cl, cl + 4, ..., cl + 28
	const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(src) + first);
	uint4 q[8];
	if (staged) { // the first quarter is on its way while the carry is looked up
#pragma unroll
		for (int k = 0; k < 8; ++k)
			q[k] = __ldg(gin + (size_t) (4 * k + cl) * (kChunk / 4) + piece);
	} else if (vec) {
#pragma unroll
		for (int j = 0; j < 8; ++j)
			q[j] = __ldg(p + j);
	}
	// carry entering this chunk = the chunk's prefix map (composition of the tile's earlier chunks, left by the maps
	// kernel) applied to the carry entering the tile (left by the scan kernel)
	const int *tc = tile_carry + (size_t) blockIdx.x * 4;
	const ByteMap *cm = chunkmaps + chunk * 4;
	const int ak = kinds.k[3];
	int carry[4];
	carry[0] = (int) cm[0].e[__ldg(tc + 0) + 7] - 7;
	carry[1] = (int) cm[1].e[__ldg(tc + 1) + 3] - 3;
	carry[2] = (int) cm[2].e[__ldg(tc + 2) + 7] - 7;
	carry[3] = ak == kChanBit1 ? balanced255(__ldg(tc + 3) + (int) cm[3].e[0]) : 0;
	if (ak == kChanShift4) { // DXT3 alpha: 31-state per-chunk maps, walked by one lane (uniform branch)
		__shared__ ByteMap s_amap[kTileThreads];
		__shared__ int s_acarry[kTileThreads];
		s_amap[t] = cm[3];
		__syncthreads();
		if (t == 0) {
			uint32_t idx = (uint32_t) (__ldg(tc + 3) + 15);
			for (int i = 0; i < kTileThreads; ++i) {
				s_acarry[i] = (int) idx - 15;
				idx = s_amap[i].e[idx];
			}
		}
		__syncthreads();
		carry[3] = s_acarry[t];
	}
	const bool has_alpha = srccomps == 4;
	if (staged) {
		uint4 *gout = reinterpret_cast<uint4 *>(out) + warp_chunk0 * (kChunk / 4);
		for (int quarter = 0; quarter < 4; ++quarter) {
#pragma unroll
			for (int k = 0; k < 8; ++k)
				*reinterpret_cast<uint4 *>(ws + (4 * k + cl) * kStagePitch + piece * 4) = q[k];
			__syncwarp();
			if (quarter < 3) {
#pragma unroll
				for (int k = 0; k < 8; ++k)
					q[k] = __ldg(gin + (size_t) (4 * k + cl) * (kChunk / 4) + (quarter + 1) * 8 + piece);
			}
			uint4 *mine = reinterpret_cast<uint4 *>(ws + lane * kStagePitch);
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				uint4 v = mine[j];
				v.x = replay_texel(carry, v.x, ak, true, alphabits);
				v.y = replay_texel(carry, v.y, ak, true, alphabits);
				v.z = replay_texel(carry, v.z, ak, true, alphabits);
				v.w = replay_texel(carry, v.w, ak, true, alphabits);
				mine[j] = v;
			}
			__syncwarp();
#pragma unroll
			for (int k = 0; k < 8; ++k)
				gout[(size_t) (4 * k + cl) * (kChunk / 4) + quarter * 8 + piece] =
						*reinterpret_cast<const uint4 *>(ws + (4 * k + cl) * kStagePitch + piece * 4);
			__syncwarp();
		}
	} else if (vec) {
		// each thread streams its own 512 contiguous bytes: 8 x 128-bit loads in flight, replay, 128-bit stores
		uint4 *o = reinterpret_cast<uint4 *>(out + first);
		for (int i = 0; i < kChunk / 4; i += 8) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				uint4 v = q[j];
				if (i + 8 < kChunk / 4)
					q[j] = __ldg(p + i + 8 + j); // next batch
				v.x = replay_texel(carry, v.x, ak, true, alphabits);
				v.y = replay_texel(carry, v.y, ak, true, alphabits);
				v.z = replay_texel(carry, v.z, ak, true, alphabits);
				v.w = replay_texel(carry, v.w, ak, true, alphabits);
				o[i + j] = v;
			}
		}
	} else {
		for (int i = 0; i < count; ++i) {
			uint32_t w;
			if (srccomps == 4)
				w = __ldg(reinterpret_cast<const uint32_t *>(src) + first + i);
			else {
				const uint8_t *qq = src + (first + i) * 3;
				w = (uint32_t) __ldg(qq) | ((uint32_t) __ldg(qq + 1) << 8) | ((uint32_t) __ldg(qq + 2) << 16);
			}
			out[first + i] = replay_texel(carry, w, ak, has_alpha, alphabits);
		}
	}
}

// Images of at most one tile (16384 texels: every mip level from 128x128 down) do all three phases in ONE CTA:
// chunk maps -> walk from the given carry -> replay.  The three-kernel path costs ~70 us of fixed latency per image,
// which dominated mip chains (12 levels per texture).
__global__ void __launch_bounds__(kTileThreads)
dither_small_kernel(const uint8_t *__restrict__ src, int srccomps, int alphabits, ChanKinds kinds, int npixels,
		const DitherLut *__restrict__ lut, int *carry /* 4 ints in/out; NULL: zero in, nothing out */, uint32_t *__restrict__ out)
{
	// one CTA per image of a batch (images back to back, npixels each)
	src += (size_t) blockIdx.x * npixels * srccomps;
	out += (size_t) blockIdx.x * npixels;
	__shared__ __align__(16) uint32_t s_lut3[256][4];
	__shared__ __align__(16) uint32_t s_lut2[256];
	__shared__ ByteMap s_maps[kTileThreads * 4];
	__shared__ int s_carry[kTileThreads * 4];
	const int t = threadIdx.x;
	load_dither_lut(lut, s_lut3, s_lut2, t);
	__syncthreads();
	const int first = t * kChunk;
	const int count = first >= npixels ? 0 : min(kChunk, npixels - first);
	auto texel = [&](int i) -> uint32_t {
		if (srccomps == 4)
			return __ldg(reinterpret_cast<const uint32_t *>(src) + first + i);
		const uint8_t *q = src + (size_t) (first + i) * 3;
		return (uint32_t) __ldg(q) | ((uint32_t) __ldg(q + 1) << 8) | ((uint32_t) __ldg(q + 2) << 16);
	};
	{
		RgbTables tab;
		rgb_tables_init(tab);
		uint32_t asum = 0;
		for (int i = count; i > 0; --i) {
			const uint32_t w = texel(i - 1);
			rgb_tables_prepend(tab, w, s_lut3, s_lut2);
			asum += w >> 24;
		}
		ByteMap m[4];
		rgb_tables_store(tab, m[0], m[1], m[2]);
		if (kinds.k[3] == kChanShift4)
			alpha_map_of_run(m[3], kChanShift4, src + (size_t) first * 4 + 3, 4, count);
		else {
#pragma unroll
			for (int k = 0; k < 32; ++k)
				m[3].e[k] = 0;
			m[3].e[0] = (uint8_t) (asum % 255u);
		}
#pragma unroll
		for (int ch = 0; ch < 4; ++ch)
			s_maps[t * 4 + ch] = m[ch];
	}
	__syncthreads();
	if ((t & 31) == 0) {
		const int ch = t >> 5;
		const int cout = walk_chunk_carries(s_maps, s_carry, ch, kinds.k[ch], carry ? carry[ch] : 0);
		if (carry)
			carry[ch] = cout;
	}
	__syncthreads();
	int cc[4] = {s_carry[t * 4 + 0], s_carry[t * 4 + 1], s_carry[t * 4 + 2], s_carry[t * 4 + 3]};
	for (int i = 0; i < count; ++i)
		out[first + i] = replay_texel(cc, texel(i), kinds.k[3], srccomps == 4, alphabits);
}

static size_t dither_tiles(size_t npixels) { return (npixels + kTilePixels - 1) / kTilePixels; }

// kernel launches of launch_prepass_simple (for the launch counters): one fused CTA for images of at most one tile,
// else maps (unless ready) + three scan launches + replay
int prepass_simple_launches(size_t npixels, bool maps_ready)
{
	if (!npixels)
		return 0;
	return !maps_ready && npixels <= (size_t) kTilePixels ? 1 : (maps_ready ? 4 : 5);
}

static size_t scan_parts_count(size_t tiles) { return (tiles + 31) / 32; } // = kScanWarpTiles tiles per partial map

// workspace: chunk prefix maps [tiles*128][4] | tile maps [tiles][4] | tile carries [tiles][4] | scan: partial maps
// [parts][4][32] bytes | scan: carries entering each part [parts][4]
size_t dither_workspace_bytes(size_t npixels)
{
	const size_t tiles = dither_tiles(npixels), parts = scan_parts_count(tiles);
	return (tiles * kTileThreads * 4 + tiles * 4) * sizeof(ByteMap) + tiles * 4 * sizeof(int) + parts * (128 + 4 * sizeof(int)) + 256;
}

static const DitherLut *device_dither_lut(cudaError_t *err)
{
	// one copy per device, built on first use (host computes 9 KB of selectors from diffuse_step itself)
	static const DitherLut *per_device[64] = {nullptr};
	int dev = 0;
	*err = cudaGetDevice(&dev);
	if (*err != cudaSuccess || dev < 0 || dev >= 64)
		return nullptr;
	if (!per_device[dev]) {
		static DitherLut host;
		build_dither_lut(host);
		DitherLut *d = nullptr;
		if ((*err = cudaMalloc((void **) &d, sizeof(DitherLut))) != cudaSuccess)
			return nullptr;
		if ((*err = cudaMemcpy(d, &host, sizeof(DitherLut), cudaMemcpyHostToDevice)) != cudaSuccess)
			return nullptr;
		per_device[dev] = d;
	}
	return per_device[dev];
}

// phases: 1 = chunk/tile maps, 2 = scan (+ summary), 4 = apply; any combination, in order
static cudaError_t run_dither(int phases, const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		int *d_carry, ByteMap *d_summary, void *d_workspace, cudaStream_t stream, unsigned tiles_per_image = 0)
{
	if (!npixels)
		return cudaSuccess;
	const size_t tiles = dither_tiles(npixels);
	ByteMap *chunkmaps = (ByteMap *) d_workspace;
	ByteMap *tilemaps = chunkmaps + tiles * kTileThreads * 4;
	int *tile_carry = (int *) (tilemaps + tiles * 4);
	uint8_t *scan_parts = (uint8_t *) (((uintptr_t) (tile_carry + tiles * 4) + 15) & ~(uintptr_t) 15);
	int *scan_starts = (int *) (scan_parts + scan_parts_count(tiles) * 128);
	const ChanKinds kinds = chan_kinds(srccomps, alphabits);
	cudaError_t e;
	const DitherLut *lut = device_dither_lut(&e);
	if (!lut)
		return e;
	if (phases == 7 && npixels <= (size_t) kTilePixels) { // one tile, full pass: the fused single-CTA kernel
		dither_small_kernel<<<1, kTileThreads, 0, stream>>>((const uint8_t *) d_src, srccomps, alphabits, kinds, (int) npixels, lut,
				d_carry, (uint32_t *) d_reduced);
		return cudaGetLastError();
	}
	if (phases & 1)
		dither_maps_kernel<<<(unsigned) tiles, kTileThreads, 0, stream>>>((const uint8_t *) d_src, srccomps, kinds, npixels, lut,
				65536u, chunkmaps, tilemaps);
	if (phases & 2) {
		static_assert(kScanWarpTiles == 32, "scan_parts_count");
		const size_t nparts = scan_parts_count(tiles);
		const unsigned ctas = (unsigned) ((tiles + kScanCtaTiles - 1) / kScanCtaTiles);
		scan_partial_kernel<<<ctas, kScanCtaWarps * 32, 0, stream>>>(tilemaps, tiles, kinds, scan_parts, tiles_per_image);
		scan_carry_kernel<<<1, 128, 0, stream>>>(scan_parts, nparts, kinds, d_carry, scan_starts, d_summary);
		if (!d_summary)
			scan_walk_kernel<<<ctas, kScanCtaWarps * 32, 0, stream>>>(tilemaps, tiles, kinds, scan_starts, tile_carry, tiles_per_image);
	}
	if (phases & 4)
		dither_apply_kernel<<<(unsigned) tiles, kTileThreads, 0, stream>>>((const uint8_t *) d_src, srccomps, alphabits,
				kinds, npixels, chunkmaps, tile_carry, (uint32_t *) d_reduced);
	return cudaGetLastError();
}

// carry entering shard `rank` = summaries of shards 0 .. rank-1 applied in order to a zero carry (one thread per channel)
// carry_in: the carry the chain starts from (NULL = zero; may be the same address as carry)
__global__ void fold_carry_kernel(const ByteMap *__restrict__ maps, int rank, ChanKinds kinds, const int *carry_in, int *carry)
{
	const int ch = threadIdx.x;
	if (ch >= 4)
		return;
	int c = carry_in ? carry_in[ch] : 0;
	for (int r = 0; r < rank; ++r)
		c = bmap_apply(maps[r * 4 + ch], kinds.k[ch], c);
	carry[ch] = c;
}

cudaError_t launch_fold_carry(const ByteMap *d_maps, int rank, int srccomps, int alphabits, int *d_carry, cudaStream_t stream)
{
	fold_carry_kernel<<<1, 32, 0, stream>>>(d_maps, rank, chan_kinds(srccomps, alphabits), nullptr, d_carry);
	return cudaGetLastError();
}

cudaError_t launch_fold_carry_from(const ByteMap *d_maps, int count, int srccomps, int alphabits, const int *d_carry_in, int *d_carry,
		cudaStream_t stream)
{
	fold_carry_kernel<<<1, 32, 0, stream>>>(d_maps, count, chan_kinds(srccomps, alphabits), d_carry_in, d_carry);
	return cudaGetLastError();
}

// `count` identity summaries (4 maps each): what an empty texel range contributes to a carry chain
__global__ void identity_maps_kernel(ByteMap *maps, int count, ChanKinds kinds)
{
	for (int i = threadIdx.x; i < count * 4; i += blockDim.x)
		bmap_identity(maps[i], kinds.k[i & 3]);
}

cudaError_t launch_identity_maps(ByteMap *d_maps, int count, int srccomps, int alphabits, cudaStream_t stream)
{
	if (count <= 0)
		return cudaSuccess;
	identity_maps_kernel<<<1, 128, 0, stream>>>(d_maps, count, chan_kinds(srccomps, alphabits));
	return cudaGetLastError();
}

cudaError_t launch_prepass_simple(const void *d_src, int srccomps, int alphabits, size_t npixels, void *d_reduced,
		int *d_carry, void *d_workspace, bool maps_ready, cudaStream_t stream)
{
	return run_dither(maps_ready ? 6 : 7, d_src, srccomps, alphabits, npixels, d_reduced, d_carry, nullptr, d_workspace, stream);
}

// DITHER_SIMPLE over `images` images of npixels texels each, stored back to back, every image starting from carry 0 (a
// batch of tx_compress_dxtn calls).  Small images: one fused CTA each, one launch; images of whole tiles: the ordinary
// three phases over the concatenated texels with the carry restarting at image boundaries; any other size: image by image.
// d_zero_carry: 4 ints that are zero (and stay zero only in the first two cases; the third rewrites them per image).
int prepass_simple_batch_launches(size_t npixels, int images)
{
	if (!npixels || images <= 0)
		return 0;
	if (npixels <= (size_t) kTilePixels)
		return 1;
	return npixels % kTilePixels == 0 ? 5 : 6 * images;
}

cudaError_t launch_prepass_simple_batch(const void *d_src, int srccomps, int alphabits, size_t npixels, int images, void *d_reduced,
		int *d_zero_carry, void *d_workspace, cudaStream_t stream)
{
	if (!npixels || images <= 0)
		return cudaSuccess;
	const ChanKinds kinds = chan_kinds(srccomps, alphabits);
	if (npixels <= (size_t) kTilePixels) {
		cudaError_t e;
		const DitherLut *lut = device_dither_lut(&e);
		if (!lut)
			return e;
		dither_small_kernel<<<images, kTileThreads, 0, stream>>>((const uint8_t *) d_src, srccomps, alphabits, kinds, (int) npixels, lut,
				nullptr, (uint32_t *) d_reduced);
		return cudaGetLastError();
	}
	if (npixels % kTilePixels == 0)
		return run_dither(7, d_src, srccomps, alphabits, npixels * images, d_reduced, d_zero_carry, nullptr, d_workspace, stream,
				(unsigned) (npixels / kTilePixels));
	for (int i = 0; i < images; ++i) {
		cudaError_t e = cudaMemsetAsync(d_zero_carry, 0, 4 * sizeof(int), stream);
		if (e == cudaSuccess)
			e = run_dither(7, (const uint8_t *) d_src + (size_t) i * npixels * srccomps, srccomps, alphabits, npixels,
					(uint32_t *) d_reduced + (size_t) i * npixels, d_zero_carry, nullptr, d_workspace, stream);
		if (e != cudaSuccess)
			return e;
	}
	return cudaSuccess;
}

cudaError_t launch_dither_summary(const void *d_src, int srccomps, int alphabits, size_t npixels, ByteMap *d_summary,
		void *d_workspace, cudaStream_t stream)
{
	return run_dither(3, d_src, srccomps, alphabits, npixels, nullptr, nullptr, d_summary, d_workspace, stream);
}

// =====================================================================================================
// Random candidates (reference s2tc_algorithm.cpp:962-993).  The reference pulls 3*nrandom (DXT5:
// 4*nrandom) values per block from one global rand() stream, block after block.  Thread t owns
// `blocks_per_thread` consecutive blocks, seeks its private replica to the first draw of its first block
// with O(log t) polynomial products (glibc_rand.cuh) and then generates sequentially.
// =====================================================================================================
// a <- a * b mod (x^31 - x^28 - 1), everything in registers (all indices compile-time)
__device__ __forceinline__ void poly_mulmod_regs(uint32_t (&a)[kLag], const uint32_t *__restrict__ bsrc)
{
	uint32_t b[kLag], t[2 * kLag - 1];
#pragma unroll
	for (int j = 0; j < kLag; ++j)
		b[j] = __ldg(bsrc + j);
#pragma unroll
	for (int i = 0; i < 2 * kLag - 1; ++i)
		t[i] = 0;
#pragma unroll
	for (int i = 0; i < kLag; ++i)
#pragma unroll
		for (int j = 0; j < kLag; ++j)
			t[i + j] += a[i] * b[j];
#pragma unroll
	for (int i = 2 * kLag - 2; i >= kLag; --i) {
		t[i - 3] += t[i];
		t[i - kLag] += t[i];
	}
#pragma unroll
	for (int i = 0; i < kLag; ++i)
		a[i] = t[i];
}

// Step 1 of candidate generation: every generator thread's window of the rand() stream,
// x^(cursor of its first block) = start * prod step[j]^(bit j of t), written as windows[slot][thread] (coalesced).
// Kept apart from the generator loop because the polynomial products need ~120 registers and the loop ~60:
// fused, the loop ran at 16 warps per SM and 22 % issue utilisation.
__global__ void __launch_bounds__(128)
rand_windows_kernel(const RandPlan *__restrict__ plan, unsigned nthreads, uint32_t *__restrict__ windows)
{
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nthreads)
		return;
	uint32_t a[kLag], w[kLag];
#pragma unroll
	for (int i = 0; i < kLag; ++i)
		a[i] = __ldg(&plan->start.c[i]);
	unsigned bits = t;
	for (int j = 0; bits; ++j, bits >>= 1)
		if (bits & 1u)
			poly_mulmod_regs(a, plan->step[j].c);
#pragma unroll
	for (int s = 0; s < kLag; ++s)
		w[s] = 0;
#pragma unroll
	for (int q = 0; q < 2 * kLag - 1; ++q) { // w[s] = sum_j a[j] * base[s + j]
		const uint32_t bq = __ldg(&plan->base[q]);
#pragma unroll
		for (int j = 0; j < kLag; ++j)
			if (q - j >= 0 && q - j < kLag)
				w[q - j] += a[j] * bq;
	}
#pragma unroll
	for (int s = 0; s < kLag; ++s)
		windows[(size_t) s * nthreads + t] = w[s];
}

// Brings one jump-ahead plan (4.3 KB) from mapped pinned host memory into device memory.  A kernel, not a
// cudaMemcpyAsync: on the compute stream a small host-to-device copy queues behind the slab uploads of the copy
// stream in the DMA engine, and the first slab's kernels then start only when the whole image has been uploaded
// (measured: config 3 end to end 85.5 ms = upload + kernels in sequence; S2TC_B200_TRACE shows the timeline).
__global__ void plan_upload_kernel(const uint32_t *__restrict__ host_plan, uint32_t *__restrict__ dev_plan, int words)
{
	for (int i = threadIdx.x; i < words; i += blockDim.x)
		dev_plan[i] = host_plan[i];
}

cudaError_t launch_plan_upload(const RandPlan *mapped_host_plan, RandPlan *d_plan, cudaStream_t stream)
{
	static_assert(sizeof(RandPlan) % 4 == 0, "plan is copied word by word");
	plan_upload_kernel<<<1, 256, 0, stream>>>((const uint32_t *) mapped_host_plan, (uint32_t *) d_plan, (int) (sizeof(RandPlan) / 4));
	return cudaGetLastError();
}

cudaError_t launch_rand_windows(const RandPlan *d_plan, unsigned nsegments, uint32_t *d_windows, cudaStream_t stream)
{
	if (nsegments == 0)
		return cudaSuccess;
	rand_windows_kernel<<<(nsegments + 127) / 128, 128, 0, stream>>>(d_plan, nsegments, d_windows);
	return cudaGetLastError();
}

// =====================================================================================================
// S3TC -> S2TC transcode (reference s2tc_from_s3tc.cpp:254-263): 8/16 bytes in, same bytes out, in place.
// Pure streaming: one 64/128-bit load and store per block.
// =====================================================================================================
__global__ void transcode_kernel(int dxt, void *blocks, size_t nblocks)
{
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += stride) {
		if (dxt == kDxt1) {
			uint2 b = reinterpret_cast<uint2 *>(blocks)[i];
			transcode_color_dxt1(b.x, b.y);
			reinterpret_cast<uint2 *>(blocks)[i] = b;
		} else {
			uint4 b = reinterpret_cast<uint4 *>(blocks)[i];
			transcode_color_opaque(b.z, b.w);
			if (dxt == kDxt5) {
				const uint64_t a = transcode_alpha_dxt5((uint64_t) b.x | ((uint64_t) b.y << 32));
				b.x = (uint32_t) a;
				b.y = (uint32_t) (a >> 32);
			}
			reinterpret_cast<uint4 *>(blocks)[i] = b;
		}
	}
}

cudaError_t launch_transcode(int dxt, void *d_blocks, size_t nblocks, cudaStream_t stream)
{
	if (!nblocks)
		return cudaSuccess;
	const int threads = 256;
	size_t blocks = (nblocks + threads - 1) / threads;
	if (blocks > 148 * 16)
		blocks = 148 * 16;
	transcode_kernel<<<(unsigned) blocks, threads, 0, stream>>>(dxt, d_blocks, nblocks);
	return cudaGetLastError();
}

// =====================================================================================================
// Mip reduction (reference Image_MipReduce32, s2tc_compress.c:427-493, as its mip loop :722-733 uses it):
// every axis still larger than 1 is halved, odd sizes drop the last row/column, texels are the truncated
// mean of the 2x2 (or 2x1 / 1x2) source texels per channel.  One thread per output texel, 32-bit loads.
// =====================================================================================================
__global__ void mip_reduce_kernel(const uint32_t *__restrict__ in, int w, int h, uint32_t *__restrict__ out, int nw, int nh, int images)
{
	const int sx = w > 1 ? 2 : 1, sy = h > 1 ? 2 : 1;
	const size_t per = (size_t) nw * nh, n = per * images, stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const size_t img = i / per, r = i - img * per;
		const int y = (int) (r / nw), x = (int) (r - (size_t) y * nw);
		const uint32_t *p = in + img * ((size_t) w * h) + (size_t) y * sy * w + (size_t) x * sx;
		const uint32_t a = __ldg(p), b = sx == 2 ? __ldg(p + 1) : 0u;
		const uint32_t c = sy == 2 ? __ldg(p + w) : 0u, d = (sx == 2 && sy == 2) ? __ldg(p + w + 1) : 0u;
		const int sh = (sx == 2) + (sy == 2);
		// two channels per word at a time: bytes 0,2 and bytes 1,3 (sums of four bytes need 10 bits)
		const uint32_t lo = (a & 0x00FF00FFu) + (b & 0x00FF00FFu) + (c & 0x00FF00FFu) + (d & 0x00FF00FFu);
		const uint32_t hi = ((a >> 8) & 0x00FF00FFu) + ((b >> 8) & 0x00FF00FFu) + ((c >> 8) & 0x00FF00FFu) + ((d >> 8) & 0x00FF00FFu);
		out[i] = ((lo >> sh) & 0x00FF00FFu) | (((hi >> sh) & 0x00FF00FFu) << 8);
	}
}

cudaError_t launch_mip_reduce(const void *d_in, int w, int h, void *d_out, cudaStream_t stream, int images)
{
	const int nw = w > 1 ? w >> 1 : w, nh = h > 1 ? h >> 1 : h;
	const size_t n = (size_t) nw * nh * images;
	if (!n || (nw == w && nh == h))
		return cudaSuccess;
	const int threads = 256;
	size_t blocks = (n + threads - 1) / threads;
	if (blocks > 148 * 32)
		blocks = 148 * 32;
	mip_reduce_kernel<<<(unsigned) blocks, threads, 0, stream>>>((const uint32_t *) d_in, w, h, (uint32_t *) d_out, nw, nh, images);
	return cudaGetLastError();
}

// =====================================================================================================
// S2TC decode (reference fetch_2d_texel_rgba_dxt1/3/5, s2tc_libtxc_dxtn.cpp:57-140; SURVEY "next" N4): whole images
// instead of one texel per call.  One thread per texel; codes that S3TC would interpolate show as a checkerboard
// of the two endpoints ((x ^ y) & 1), DXT1 code 3 with c1 >= c0 is transparent black.
// =====================================================================================================
__global__ void decode_kernel(int dxt, const uint8_t *__restrict__ blocks, int width, int height, uint32_t *__restrict__ out)
{
	const size_t n = (size_t) width * height, stride = (size_t) gridDim.x * blockDim.x;
	const int bw = (width + 3) >> 2, bs = dxt == kDxt1 ? 8 : 16;
	for (size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
		const int y = (int) (p / width), x = (int) (p - (size_t) y * width);
		const uint8_t *blk = blocks + ((size_t) (y >> 2) * bw + (x >> 2)) * bs;
		const uint2 cw = __ldg(reinterpret_cast<const uint2 *>(blk + (dxt == kDxt1 ? 0 : 8)));
		const uint32_t c0 = cw.x & 0xFFFFu, c1 = cw.x >> 16;
		const uint32_t code = (cw.y >> (2 * ((y & 3) * 4 + (x & 3)))) & 3u;
		uint32_t c = c0, alpha = 255;
		if (code == 1)
			c = c1;
		else if (code == 3 && dxt == kDxt1 && c1 >= c0) {
			c = 0;
			alpha = 0;
		} else if (code >= 2 && ((x ^ y) & 1))
			c = c1;
		const uint32_t r5 = (c >> 11) & 31u, g6 = (c >> 5) & 63u, b5 = c & 31u;
		const uint32_t r = (r5 << 3) | (r5 >> 2), g = (g6 << 2) | (g6 >> 4), b = (b5 << 3) | (b5 >> 2);
		if (dxt == kDxt3) {
			const uint2 aw = __ldg(reinterpret_cast<const uint2 *>(blk));
			const int i = (y & 3) * 4 + (x & 3);
			const uint32_t a4 = ((i < 8 ? aw.x >> (4 * i) : aw.y >> (4 * (i - 8)))) & 15u;
			alpha = a4 | (a4 << 4);
		} else if (dxt == kDxt5) {
			const uint2 aw = __ldg(reinterpret_cast<const uint2 *>(blk));
			const uint32_t a0 = aw.x & 0xFFu, a1 = (aw.x >> 8) & 0xFFu;
			const uint64_t bits = (((uint64_t) aw.y << 32) | aw.x) >> 16;
			const uint32_t ac = (uint32_t) (bits >> (3 * ((y & 3) * 4 + (x & 3)))) & 7u;
			alpha = a0;
			if (ac == 1)
				alpha = a1;
			else if (ac == 6 && a1 >= a0)
				alpha = 0;
			else if (ac == 7 && a1 >= a0)
				alpha = 255;
			else if (ac >= 2 && ((x ^ y) & 1))
				alpha = a1;
		}
		out[p] = r | (g << 8) | (b << 16) | (alpha << 24);
	}
}

cudaError_t launch_decode(int dxt, const void *d_blocks, int width, int height, void *d_rgba, cudaStream_t stream)
{
	const size_t n = (size_t) width * height;
	if (!n)
		return cudaSuccess;
	const int threads = 256;
	size_t blocks = (n + threads - 1) / threads;
	if (blocks > 148 * 32)
		blocks = 148 * 32;
	decode_kernel<<<(unsigned) blocks, threads, 0, stream>>>(dxt, (const uint8_t *) d_blocks, width, height, (uint32_t *) d_rgba);
	return cudaGetLastError();
}

// =====================================================================================================
// Measurement aid: sustained integer min+add rate, the denominator for the search kernels' roofline
// (search modes are bound by the integer pipes, not by HBM).  8 independent chains per thread of
// exactly the two operations the pair scan is made of, in the three instruction mixes the kernels use:
//   MODE 0  s += min(a, b)                       as the compiler schedules it (VIMNMX + IADD3, mostly the ALU pipe)
//   MODE 1  s = min(a, b) * one + s              VIMNMX on the ALU pipe, IMAD on the FMA pipe (search16's 32-bit scan)
//   MODE 2  s = dp2a(vminu2(a, b), 0x0101, s)    VIMNMX.U16x2 + IDP.2A: two mins and two adds per instruction pair
//                                                (the 16-bit packed rows of search16 and pair_search)
// =====================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) int32_peak_kernel(int iters, int seed, uint32_t one, int *sink)
{
	uint32_t a[8], s[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		a[k] = (uint32_t) (seed + threadIdx.x * (k + 1)) & 0x3FFF3FFFu;
		s[k] = 0;
	}
	uint32_t b = (uint32_t) (seed ^ (blockIdx.x << 8)) & 0x3FFF3FFFu;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			if (MODE == 0)
				s[k] += min(a[k], b);
			else if (MODE == 1)
				asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s[k]) : "r"(min(a[k], b)), "r"(one));
			else
				s[k] = __dp2a_lo(__vminu2(a[k], b), 0x0101u, s[k]);
		}
		b = (b + 3u) & 0x3FFF3FFFu;
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k)
		r ^= s[k];
	if (r == 0x7FFFFFFFu)
		*sink = (int) r;
}

cudaError_t launch_int32_peak(int mode, int iters, int ctas, int *d_sink, cudaStream_t stream)
{
	if (mode == 0)
		int32_peak_kernel<0><<<ctas, 256, 0, stream>>>(iters, 12345, 1u, d_sink);
	else if (mode == 1)
		int32_peak_kernel<1><<<ctas, 256, 0, stream>>>(iters, 12345, 1u, d_sink);
	else
		int32_peak_kernel<2><<<ctas, 256, 0, stream>>>(iters, 12345, 1u, d_sink);
	return cudaGetLastError();
}

} // namespace s2tc
