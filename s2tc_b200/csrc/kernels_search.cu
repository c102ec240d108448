// kernels_search.cu -- MODE_NORMAL with random candidates (S2TC_RANDOM_COLORS > 0): the c0/c1 pair search
// (reference reduce_colors_inplace and reduce_colors_inplace_2fixpoints, s2tc_algorithm.cpp:367-478, reached
// from :1004-1006), one warp per 4x4 block.
//
// Per block the reference builds dists[m][n] (m = n gathered colours + nrandom random candidates, n <= 16) and
// scans all m(m-1)/2 pairs for the smallest sum_k min(d[i][k], d[j][k]): 3160 pairs x 16 texels at
// nrandom = 64, >90 % of its run time.  Here:
//   * the warp gathers the block's colours, appends the pre-generated random candidates and fills the
//     distance matrix in shared memory, one row per candidate, columns zero-padded to 16;
//   * metrics whose distances fit 15 bits (AVG, WAVG, W0AVG <= 20681) and alpha (<= 65025) are stored as
//     16-bit halves: a row is 8 words, a pair costs 8 VIMNMX.U16x2 and the 16-term sum is formed with
//     packed three-operand adds (3 x 20681 < 2^16) and three IDP.2A horizontal adds; the other metrics keep
//     32-bit rows (16 VIMNMX + IADD3 tree);
//   * the pair triangle is walked in 16x16 tiles: lane (jj, half) keeps row j = 16b+jj of tile column b in
//     registers and meets rows i = 16a + 8*half + t, t = 0..7, whose loads are warp broadcasts (two distinct
//     addresses per LDS.128), so shared-memory traffic is ~2 wavefronts per 32 pairs; row pitches (12 / 20
//     words) make the per-lane row loads conflict-free;
//   * every lane keeps its minimum by (sum, i, j) and the warp merges lexicographically, which is the
//     reference's "first minimum in (i, j) order";
//   * when a sum went negative (only the SRGB metric can wrap) lane 0 replays the reference's
//     acceptance rule "bestsum < 0 || sum < bestsum" verbatim over the stored matrix;
//   * DXT5 repeats the search for alpha, the two fixed points 0 and 255 folded into every row
//     (min(d[i][k], f[k]) is stored, so the pair loop is unchanged).
// Output: the chosen endpoints, 8 bytes per block; kernels_finish.cu turns them into DXT blocks.
// History: the first version (32-bit rows, lanes striding j with 70 % utilisation) took 198 ms for the
// 16.7 M blocks of config 3 (profiles/r01c).
#define S2TC_USE_SRGB_MIXED_LUT
#include "kernels.cuh"

namespace s2tc {

constexpr int kSearchThreads = 128;
constexpr int kSearchWarps = kSearchThreads / 32;
constexpr int kPitch16 = 12; // words per row, 16-bit distances (8 used): 8 consecutive rows -> 8 distinct 16-byte slots
constexpr int kPitch32 = 20; // words per row, 32-bit distances (16 used)

template <int CD> struct Packs16 { static constexpr bool value = CD == kAVG || CD == kWAVG || CD == kW0AVG; };

// per-warp shared memory: texels [16] | exact rows [(mcap+16)][pitch] | quantised rows [(mcap+16)] x 16 B | c [(mcap+16)] |
// colours [mcap] | features [mcap]
__host__ __device__ inline size_t search_warp_bytes(int mcap, bool pack16, bool has_feat)
{
	const size_t rows = (size_t) (mcap + 16) * (pack16 ? kPitch16 : kPitch32) * 4; // +16: tile loads may touch one tile past m
	size_t b = 64 + rows + (size_t) (mcap + 16) * 20 + (size_t) mcap * 4 + (has_feat ? (size_t) mcap * 12 : 0);
	return (b + 15) & ~(size_t) 15;
}

// one texel row of a block as reduced texels (zeros outside the image)
__device__ __forceinline__ void load_block_row(const ImageView &v, int x0, int y, int w, uint32_t t[4])
{
	t[0] = t[1] = t[2] = t[3] = 0;
	if (y >= v.rows)
		return;
	if (v.fmt != kSrcRGB8) {
		const size_t pitch = (size_t) v.width * 4;
		const uint8_t *row = v.base + (size_t) y * pitch + (size_t) x0 * 4;
		if (w == 4 && ((pitch | (size_t) v.base) & 15) == 0) {
			const uint4 q = __ldg(reinterpret_cast<const uint4 *>(row));
			t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
		} else {
			for (int x = 0; x < w; ++x)
				t[x] = __ldg(reinterpret_cast<const uint32_t *>(row) + x);
		}
		if (v.fmt == kSrcRGBA8)
			for (int x = 0; x < 4; ++x)
				t[x] = reduce_word(t[x], v.alphabits);
		for (int x = w; x < 4; ++x)
			t[x] = 0;
	} else {
		const size_t pitch = (size_t) v.width * 3;
		const uint8_t *row = v.base + (size_t) y * pitch + (size_t) x0 * 3;
		const uint32_t ones = ((1u << v.alphabits) - 1u) << 24;
		for (int x = 0; x < w; ++x) {
			const uint8_t *q = row + x * 3;
			t[x] = (uint32_t) (__ldg(q) >> 3) | ((uint32_t) (__ldg(q + 1) >> 2) << 8) | ((uint32_t) (__ldg(q + 2) >> 3) << 16) | ones;
		}
	}
}

// ---- sum_k min(a[k], b[k]) over one row pair ----------------------------------------------------------
// 16-bit halves.  SUM3: every value <= 21845, so three packed words can be added before widening.
template <bool SUM3>
__device__ __forceinline__ int pair_sum_p16(const uint32_t (&a)[8], const uint32_t (&b)[8])
{
	uint32_t m[8];
#pragma unroll
	for (int q = 0; q < 8; ++q)
		m[q] = __vminu2(a[q], b[q]);
	uint32_t s = 0;
	if (SUM3) {
		s = __dp2a_lo(m[0] + m[1] + m[2], 0x0101u, s);
		s = __dp2a_lo(m[3] + m[4] + m[5], 0x0101u, s);
		s = __dp2a_lo(m[6] + m[7], 0x0101u, s);
	} else {
#pragma unroll
		for (int q = 0; q < 8; ++q)
			s = __dp2a_lo(m[q], 0x0101u, s);
	}
	return (int) s;
}

__device__ __forceinline__ int pair_sum_32(const int (&a)[16], const int (&b)[16])
{
	uint32_t s[4];
#pragma unroll
	for (int q = 0; q < 4; ++q)
		s[q] = (uint32_t) min(a[4 * q], b[4 * q]) + (uint32_t) min(a[4 * q + 1], b[4 * q + 1]) +
				((uint32_t) min(a[4 * q + 2], b[4 * q + 2]) + (uint32_t) min(a[4 * q + 3], b[4 * q + 3]));
	return (int) ((s[0] + s[1]) + (s[2] + s[3]));
}

template <bool PACK16> struct RowRegs;
template <> struct RowRegs<true> {
	uint32_t w[8];
	__device__ __forceinline__ void load(const uint32_t *rows, int r)
	{
		const uint4 a = *reinterpret_cast<const uint4 *>(rows + r * kPitch16), b = *reinterpret_cast<const uint4 *>(rows + r * kPitch16 + 4);
		w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
	}
};
template <> struct RowRegs<false> {
	int w[16];
	__device__ __forceinline__ void load(const uint32_t *rows, int r)
	{
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const int4 a = *reinterpret_cast<const int4 *>(rows + r * kPitch32 + 4 * q);
			w[4 * q] = a.x; w[4 * q + 1] = a.y; w[4 * q + 2] = a.z; w[4 * q + 3] = a.w;
		}
	}
};

// rank of pair (i, j), i < j < m, in the reference's lexicographic scan order
__device__ __forceinline__ int pair_rank(int i, int j, int m) { return i * m - ((i * (i + 1)) >> 1) + (j - i - 1); }

// pair number p (lexicographic over i < j < 16) -> i | j << 8, padded to 128 entries
struct DiagPairs {
	uint16_t v[128];
	constexpr DiagPairs() : v()
	{
		int p = 0;
		for (int i = 0; i < 16; ++i)
			for (int j = i + 1; j < 16; ++j)
				v[p++] = (uint16_t) (i | (j << 8));
		for (; p < 128; ++p)
			v[p] = (uint16_t) (0 | (1 << 8));
	}
};
__device__ const DiagPairs kDiagPairs{};

// ---- pruned scan ---------------------------------------------------------------------------------------
// The exact pair scan costs 16 min + 16 add per pair.  Most pairs are nowhere near the minimum, and a cheap LOWER
// BOUND of a pair's sum shows it:  with q[i][k] = min(255, d[i][k] >> s)  (one byte per texel, 16 bytes per row)
//     sum_k min(d[i][k], d[j][k])  >=  2^s * sum_k min(q[i][k], q[j][k])  =  2^s * (Rq[i] + Rq[j] - SAD(q[i], q[j])) / 2,
// Rq = row sums, SAD = sum of absolute byte differences: four VABSDIFF4.U8.ACC per pair, 16 texels in 4 instructions.
// A pair can only beat (or tie) the best exact sum T found so far if its bound is <= T; everything else is skipped
// without ever being evaluated exactly.  The result is the reference's: every pair whose exact sum could be the
// first minimum in (i, j) order is evaluated exactly and compared by (sum, rank).
//   0. the 120 pairs of the first diagonal tile (the block's own colours) are scanned exactly: T, and from T the
//      shift s (values above T never matter, so 8 bits cover [0, T] as finely as they can);
//   1. every lane quantises rows: q bytes and c[i] = K - Rq[i];
//   2. 16x16 tiles as before: lane (jj, half) keeps the q rows j of up to four tile columns in registers and meets the
//      eight rows i = 16a + 8 half + t (broadcast loads).  acc = c[i] + SAD(q[i], q[j]) = K + Rq[j] - 2 bound, so the pair
//      survives iff acc >= K + Rq[j] - 2 (T >> s), a per-lane constant: the loop only keeps max_t acc (VIMNMX3);
//   3. after each group of tiles the (rare) lanes whose maximum passes walk their eight rows again, evaluate the
//      surviving pairs exactly from the full-precision rows and the warp agrees on the new (T, rank).
// Diagonal tiles go through the same loop (slots with i >= j only raise false alarms that step 3 discards); rows
// beyond m are all-255 (bound = Rq[i], never better than a real pair of row i).
// Measured on a B200 (tools_lab/ubench_sad.cu): VABSDIFF4 issues every other clock per scheduler on the ALU pipe,
// so the bound costs ~9 clocks per pair against ~21 for the exact form.
constexpr uint32_t kBoundBias = 4096; // K > 16 * 255

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
template <int NW> __device__ __forceinline__ uint32_t sad_row(const uint4 &a, const uint4 &b, uint32_t c)
{
	c = sad4(a.x, b.x, c);
	if (NW > 1)
		c = sad4(a.y, b.y, c);
	if (NW > 2)
		c = sad4(a.z, b.z, c);
	if (NW > 3)
		c = sad4(a.w, b.w, c);
	return c;
}

// exact sum_k min(d[i][k], d[j][k]) from the full-precision rows
template <bool PACK16> __device__ __forceinline__ uint32_t exact_pair_sum(const uint32_t *rows, int i, int j)
{
	RowRegs<PACK16> ri, rj;
	ri.load(rows, i);
	rj.load(rows, j);
	if constexpr (PACK16) {
		uint32_t s = 0;
#pragma unroll
		for (int w = 0; w < 8; ++w)
			s = __dp2a_lo(__vminu2(ri.w[w], rj.w[w]), 0x0101u, s);
		return s;
	} else {
		return (uint32_t) pair_sum_32(ri.w, rj.w);
	}
}

struct BestPair { // lexicographic (sum, rank): the reference's first minimum
	uint32_t sum, rank;
	__device__ __forceinline__ void take(uint32_t s, uint32_t r)
	{
		if (s < sum || (s == sum && r < rank)) {
			sum = s;
			rank = r;
		}
	}
	__device__ __forceinline__ void warp_min()
	{
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			const uint32_t os = __shfl_xor_sync(0xFFFFFFFFu, sum, off), orank = __shfl_xor_sync(0xFFFFFFFFu, rank, off);
			take(os, orank);
		}
	}
};

// rows q of one lane for NC tile columns against the eight rows i0 .. i0+7: per-column maximum of acc
template <int NW, int NC>
__device__ __forceinline__ void bound_tile(const uint4 *q8, const uint32_t *cneg, int i0, const uint4 (&rj)[4], uint32_t (&mx)[4])
{
#pragma unroll
	for (int c = 0; c < 4; ++c)
		mx[c] = 0;
#pragma unroll 4
	for (int t = 0; t < 8; t += 2) {
		const uint4 qa = q8[i0 + t], qb = q8[i0 + t + 1];
		const uint32_t ca = cneg[i0 + t], cb = cneg[i0 + t + 1];
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			const uint32_t x = sad_row<NW>(qa, rj[c], ca), y = sad_row<NW>(qb, rj[c], cb);
			mx[c] = max(mx[c], max(x, y));
		}
	}
}

// Returns (i << 16) | j of the winner in every lane.  rows: exact distance rows (16-bit packed, pitch kPitch16, or 32-bit,
// pitch kPitch32; all values >= 0); q8 / cneg: workspace for 16 * ntile quantised rows.  n: columns in use.
template <bool PACK16>
__device__ __forceinline__ uint32_t scan_pruned(const uint32_t *rows, uint4 *q8, uint32_t *cneg, int m, int n, int lane, int sadj)
{
	const int ntile = (m + 15) >> 4;
	const int jj = lane & 15, half = lane >> 4;
	// 0. first diagonal tile, exactly
	BestPair best{0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t ij = __ldg(&kDiagPairs.v[lane + 32 * q]);
		const int i = (int) (ij & 0xFFu), j = (int) (ij >> 8);
		if ((q < 3 || lane < 24) && j < m)
			best.take(exact_pair_sum<PACK16>(rows, i, j), (uint32_t) pair_rank(i, j, m));
	}
	best.warp_min();
	if (best.sum != 0 || best.rank != 0) { // pair (0, 1) with sum 0 cannot be beaten
		// 1. quantise
		const int s = max(0, 32 - __clz(best.sum) - 8 + sadj);
		for (int r = lane; r < 16 * ntile; r += 32) {
			uint4 q = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
			if (r < m) {
				uint32_t h[8];
				if constexpr (PACK16) {
					RowRegs<true> rr;
					rr.load(rows, r);
					const uint32_t mask = (0xFFFFu >> s) * 0x10001u;
#pragma unroll
					for (int w = 0; w < 8; ++w)
						h[w] = __vminu2((rr.w[w] >> s) & mask, 0x00FF00FFu);
				} else {
					RowRegs<false> rr;
					rr.load(rows, r);
#pragma unroll
					for (int w = 0; w < 8; ++w)
						h[w] = min((uint32_t) rr.w[2 * w] >> s, 255u) | (min((uint32_t) rr.w[2 * w + 1] >> s, 255u) << 16);
				}
				q.x = __byte_perm(h[0], h[1], 0x6420);
				q.y = __byte_perm(h[2], h[3], 0x6420);
				q.z = __byte_perm(h[4], h[5], 0x6420);
				q.w = __byte_perm(h[6], h[7], 0x6420);
			}
			q8[r] = q;
			cneg[r] = kBoundBias - sad4(q.w, 0u, sad4(q.z, 0u, sad4(q.y, 0u, sad4(q.x, 0u, 0u))));
		}
		__syncwarp();
		// 2./3. bound scan with exact evaluation of the survivors
		const int nw = (n + 3) >> 2;
		uint32_t tq2 = (best.sum >> s) * 2u;
		for (int a = 0; a < ntile; ++a) {
			const int i0 = 16 * a + 8 * half;
			for (int b0 = a ? a : 1; b0 < ntile; b0 += 4) {
				const int nc = min(4, ntile - b0);
				uint4 rj[4];
				uint32_t rq[4], mx[4];
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const int j = 16 * (b0 + (c < nc ? c : 0)) + jj;
					rj[c] = q8[j];
					rq[c] = 2u * kBoundBias - cneg[j]; // K + Rq[j]
				}
				switch ((nw - 1) * 4 + nc - 1) {
#define S2TC_BT(NW, NC) case (NW - 1) * 4 + NC - 1: bound_tile<NW, NC>(q8, cneg, i0, rj, mx); break;
				S2TC_BT(1, 1) S2TC_BT(1, 2) S2TC_BT(1, 3) S2TC_BT(1, 4)
				S2TC_BT(2, 1) S2TC_BT(2, 2) S2TC_BT(2, 3) S2TC_BT(2, 4)
				S2TC_BT(3, 1) S2TC_BT(3, 2) S2TC_BT(3, 3) S2TC_BT(3, 4)
				S2TC_BT(4, 1) S2TC_BT(4, 2) S2TC_BT(4, 3) default: bound_tile<4, 4>(q8, cneg, i0, rj, mx); break;
#undef S2TC_BT
				}
				bool hit = false;
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					// survives iff acc >= K + Rq[j] - 2 (T >> s); signed: the right-hand side may be negative
					const bool f = c < nc && (int) mx[c] >= (int) (rq[c] - tq2);
					if (__any_sync(0xFFFFFFFFu, f)) {
						hit = true;
						if (f) {
							const int j = 16 * (b0 + c) + jj;
							const int thr = (int) (rq[c] - tq2);
							for (int t = 0; t < 8; ++t) {
								const int i = i0 + t;
								const uint32_t acc = sad_row<4>(q8[i], rj[c], cneg[i]);
								if ((int) acc >= thr && i < j && j < m)
									best.take(exact_pair_sum<PACK16>(rows, i, j), (uint32_t) pair_rank(i, j, m));
							}
						}
					}
				}
				if (hit) {
					best.warp_min();
					tq2 = (best.sum >> s) * 2u;
				}
			}
		}
	}
	// rank -> (i, j): row i of the pair order starts at rank i (m - 1) - i (i - 1) / 2; every lane tests its rows and the
	// one that holds the rank announces itself
	const int rank = (int) best.rank;
	uint32_t mine = 0;
	for (int i = lane; i < m - 1; i += 32) {
		const int start = i * (m - 1) - ((i * (i - 1)) >> 1);
		if (rank >= start && rank < start + (m - 1 - i))
			mine = ((uint32_t) i << 16) | (uint32_t) (i + 1 + rank - start);
	}
	return __reduce_or_sync(0xFFFFFFFFu, mine);
}

// Generic scan: any row width, any sum range.  Returns (i << 16) | j of the winner in every lane.
template <bool PACK16, bool SUM3, bool MAY_BE_NEGATIVE>
__device__ __forceinline__ uint32_t scan_tiles(const uint32_t *rows, int m, int lane)
{
	const int jj = lane & 15, half = lane >> 4;
	const int ntile = (m + 15) >> 4;
	int best = 0x7FFFFFFF;
	uint32_t bij = 1u; // (0,1), the reference's initial besti/bestj
	bool negative = false;
	for (int b = 0; b < ntile; ++b) {
		const int j = 16 * b + jj;
		RowRegs<PACK16> rj;
		rj.load(rows, j); // rows beyond m are inside the allocation (padding tile) and never accepted
		for (int a = 0; a <= b; ++a) {
			const int i0 = 16 * a + 8 * half;
#pragma unroll 4
			for (int t = 0; t < 8; ++t) {
				const int i = i0 + t;
				RowRegs<PACK16> ri;
				ri.load(rows, i);
				int sum;
				if constexpr (PACK16)
					sum = pair_sum_p16<SUM3>(ri.w, rj.w);
				else
					sum = pair_sum_32(ri.w, rj.w);
				const uint32_t ij = ((uint32_t) i << 16) | (uint32_t) j;
				const bool valid = i < j && j < m;
				if (MAY_BE_NEGATIVE)
					negative |= valid && sum < 0;
				if (valid && (sum < best || (sum == best && ij < bij))) {
					best = sum;
					bij = ij;
				}
			}
		}
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) {
		const int ob = __shfl_xor_sync(0xFFFFFFFFu, best, off);
		const uint32_t oij = __shfl_xor_sync(0xFFFFFFFFu, bij, off);
		if (ob < best || (ob == best && oij < bij)) {
			best = ob;
			bij = oij;
		}
	}
	if (MAY_BE_NEGATIVE) { // 32-bit rows only
		if (__any_sync(0xFFFFFFFFu, negative)) { // rare: replay the reference's rule verbatim on one lane (ref :393-410)
			if (lane == 0) {
				int bestsum = -1;
				bij = 1u;
				for (int i = 0; i < m; ++i)
					for (int j = i + 1; j < m; ++j) {
						uint32_t s = 0;
						for (int k = 0; k < 16; ++k)
							s += (uint32_t) min((int) rows[i * kPitch32 + k], (int) rows[j * kPitch32 + k]);
						const int sum = (int) s;
						if (bestsum < 0 || sum < bestsum) {
							bestsum = sum;
							bij = ((uint32_t) i << 16) | (uint32_t) j;
						}
					}
			}
			bij = __shfl_sync(0xFFFFFFFFu, bij, 0);
		}
	}
	return bij;
}

template <int DXT, int CD>
__global__ void __launch_bounds__(kSearchThreads)
pair_search_kernel(ImageView v, int nrandom, int mcap, size_t warp_bytes, int sadj, const uint16_t *__restrict__ cand_c,
		const uint8_t *__restrict__ cand_a, uint2 *__restrict__ ends)
{
	typedef Metric<CD> M;
	typedef typename M::Feat Feat;
	constexpr bool kPack = Packs16<CD>::value;
	constexpr int kPitch = kPack ? kPitch16 : kPitch32;
	extern __shared__ __align__(16) uint8_t smem[];
	const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nblocks = v.blocks_w * v.blocks_h;
	const int t = blockIdx.x * kSearchWarps + wi;
	if (t >= nblocks)
		return;

	uint8_t *wbase = smem + (size_t) wi * warp_bytes;
	uint32_t *px = reinterpret_cast<uint32_t *>(wbase);                                     // [16]
	uint32_t *rows = reinterpret_cast<uint32_t *>(wbase + 64);                              // [(mcap+16)][kPitch]
	uint4 *q8 = reinterpret_cast<uint4 *>(rows + (size_t) (mcap + 16) * kPitch);            // [(mcap+16)] quantised rows
	uint32_t *cneg = reinterpret_cast<uint32_t *>(q8 + (mcap + 16));                        // [(mcap+16)]
	uint32_t *col = cneg + (mcap + 16);                                                     // [mcap]
	Feat *feat = reinterpret_cast<Feat *>(col + mcap);                                      // [mcap] features (32-bit metrics) or scaled colours

	const int by = t / v.blocks_w, bx = t - by * v.blocks_w;
	const int x0 = bx * 4, y0 = by * 4;
	const int w = min(4, v.width - x0), h = min(4, v.rows - y0);

	// 1. texels -> shared
	if (lane < 4) {
		uint32_t r[4];
		load_block_row(v, x0, y0 + lane, w, r);
#pragma unroll
		for (int x = 0; x < 4; ++x)
			px[lane * 4 + x] = r[x];
	}
	__syncwarp();

	// 2. gather in the reference's column-major order (ref :940-959); bit o = x*4+y
	// lane o < 16 looks at texel (x, y) = (o >> 2, o & 3); one ballot gives the mask of gathered texels
	const uint32_t valid = valid_mask(w, h);
	const int ti = (lane & 3) * 4 + ((lane >> 2) & 3); // texel index y * 4 + x of lane o
	const uint32_t mine = px[ti];
	bool use = lane < 16 && ((valid >> ti) & 1u);
	if (DXT == kDxt1)
		use = use && (mine >> 24) != 0;
	const uint32_t usemask = __ballot_sync(0xFFFFFFFFu, use);
	int n = __popc(usemask);
	if (use)
		col[__popc(usemask & ((1u << lane) - 1u))] = mine;
	if (n == 0) {
		if (lane == 0)
			col[0] = 0;
		n = 1;
	}
	{ // ref :962-993, candidates pre-generated by random_candidates_kernel
		const size_t cb = (size_t) t * nrandom;
		for (int k = lane; k < nrandom; k += 32) {
			uint32_t c = from565(cand_c[cb + k]);
			if (DXT == kDxt5)
				c |= (uint32_t) cand_a[cb + k] << 24;
			col[n + k] = c;
		}
	}
	const int m = n + nrandom;
	const int mpad = (m + 15) & ~15; // rows of the last tile beyond m are zero-filled
	__syncwarp();

	// 3. distance matrix, columns zero-padded to 16 (ref :375-392; argument order matters for SRGB)
	if constexpr (kPack) {
		// AVG / WAVG / W0AVG: the metric's feature is the colour with pre-scaled channel bytes and a distance is one
		// per-byte subtraction + IDP.4A (colordist.cuh)
		uint32_t *cvec = reinterpret_cast<uint32_t *>(feat); // [mcap]
		for (int i = lane; i < m; i += 32)
			cvec[i] = M::feat(col[i]).v;
		__syncwarp();
		// lane -> one packed word per row (texel columns 2kq, 2kq+1), 4 rows per step; everything the lane needs from its
		// columns is loop-invariant.  Reads of cvec beyond m return leftovers that the masks discard.
		const int kq = lane & 7;
		const uint32_t ck0 = cvec[2 * kq], ck1 = cvec[2 * kq + 1];
		const uint32_t kmask = (2 * kq < n ? 0x0000FFFFu : 0u) | (2 * kq + 1 < n ? 0xFFFF0000u : 0u);
#pragma unroll 2
		for (int i = lane >> 3; i < mpad; i += 4) {
			const uint32_t ci = cvec[i];
			const uint32_t d0 = (uint32_t) M::dist(FeatBytes{ci}, FeatBytes{ck0}), d1 = (uint32_t) M::dist(FeatBytes{ci}, FeatBytes{ck1});
			rows[i * kPitch + kq] = i < m ? ((d0 | (d1 << 16)) & kmask) : 0u;
		}
	} else {
		for (int i = lane; i < m; i += 32)
			feat[i] = M::feat(col[i]);
		__syncwarp();
		// lane -> texel column k, 2 rows per step
		const int k = lane & 15;
		const Feat fk = feat[k];
		const bool kok = k < n;
		for (int i = lane >> 4; i < mpad; i += 2) {
			int d = 0;
			if (i < m && kok && k != i) {
				const Feat fi = feat[i];
				d = (i < n && k < i) ? M::dist(fk, fi) : M::dist(fi, fk);
			}
			rows[i * kPitch + k] = (uint32_t) d;
		}
	}
	__syncwarp();

	// 4. colour pair scan
	uint32_t cij;
	if constexpr (M::kMayBeNegative) // SRGB: sums can wrap negative, no lower bound to prune with
		cij = scan_tiles<kPack, true, true>(rows, m, lane);
	else
		cij = scan_pruned<kPack>(rows, q8, cneg, m, n, lane, sadj);
	const uint32_t c0 = col[cij >> 16], c1 = col[cij & 0xFFFFu];
	uint32_t a01 = 0;

	if (DXT == kDxt5) { // ref :416-478; alpha rows are always 16-bit, at the 16-bit pitch inside the same buffer
		__syncwarp();
		{ // the fixed points 0 and 255 folded into every row: min(d[i][k], fix[k]) is stored (masked columns: fix = 0)
			const int kq = lane & 7;
			const uint32_t ak0 = col[2 * kq] >> 24, ak1 = col[2 * kq + 1] >> 24;
			const uint32_t f0 = min(ak0 * ak0, (255u - ak0) * (255u - ak0)), f1 = min(ak1 * ak1, (255u - ak1) * (255u - ak1));
			const uint32_t fixw = (2 * kq < n ? f0 : 0u) | (2 * kq + 1 < n ? f1 << 16 : 0u);
#pragma unroll 2
			for (int i = lane >> 3; i < mpad; i += 4) {
				const uint32_t ai = col[i] >> 24;
				const uint32_t t0 = ai - ak0, t1 = (ai - ak1) << 8; // wrapping: the squares are exact mod 2^32
				rows[i * kPitch16 + kq] = i < m ? __vminu2(t1 * t1 + t0 * t0, fixw) : 0u;
			}
		}
		__syncwarp();
		// alpha sums < 16 * 65025 < 2^20, up to 4095 pairs (m <= 90) for the keyed scan
		const uint32_t aij = scan_pruned<true>(rows, q8, cneg, m, n, lane, sadj);
		a01 = (col[aij >> 16] >> 24) | ((col[aij & 0xFFFFu] >> 24) << 8);
	}
	if (lane == 0)
		ends[t] = make_uint2(to565(c0) | (to565(c1) << 16), a01);
}

static size_t search_smem(int cd, int nrandom)
{
	const bool pack = cd == kAVG || cd == kWAVG || cd == kW0AVG;
	return search_warp_bytes(16 + nrandom, pack, true) * kSearchWarps;
}

template <int DXT, int CD>
static cudaError_t launch_search_cd(int nrandom, const ImageView &v, const uint16_t *cand_c, const uint8_t *cand_a,
		uint2 *ends, cudaStream_t stream)
{
	const int nblocks = v.blocks_w * v.blocks_h;
	if (nblocks == 0)
		return cudaSuccess;
	const int mcap = 16 + nrandom;
	const size_t wb = search_warp_bytes(mcap, Packs16<CD>::value, true);
	const size_t smem = wb * kSearchWarps;
	auto kern = pair_search_kernel<DXT, CD>;
	if (smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess)
			return e;
	}
	const dim3 block(kSearchThreads), grid((nblocks + kSearchWarps - 1) / kSearchWarps);
	static const int sadj = [] { const char *e = getenv("S2TC_B200_SADJ"); return e ? atoi(e) : 0; }();
	kern<<<grid, block, smem, stream>>>(v, nrandom, mcap, wb, sadj, cand_c, cand_a, ends);
	return cudaGetLastError();
}

int pair_search_max_nrandom()
{
	// the 32-bit layout is the larger one; 227 KB of shared memory per CTA
	int lo = 0, hi = 1 << 15;
	while (lo < hi) {
		const int mid = (lo + hi + 1) / 2;
		if (search_smem(kRGB, mid) <= 227 * 1024)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

template <int DXT>
static cudaError_t launch_search_dxt(int cd, int nrandom, const ImageView &v, const uint16_t *cand_c,
		const uint8_t *cand_a, uint2 *ends, cudaStream_t stream)
{
	if (nrandom <= 0 || nrandom > pair_search_max_nrandom())
		return cudaErrorInvalidValue; // nrandom <= 0 is served by the fused 16-candidate encoder (search16.inl)
	switch (cd) {
	case kRGB: return launch_search_cd<DXT, kRGB>(nrandom, v, cand_c, cand_a, ends, stream);
	case kYUV: return launch_search_cd<DXT, kYUV>(nrandom, v, cand_c, cand_a, ends, stream);
	case kSRGB: return launch_search_cd<DXT, kSRGB>(nrandom, v, cand_c, cand_a, ends, stream);
	case kSRGB_MIXED: return launch_search_cd<DXT, kSRGB_MIXED>(nrandom, v, cand_c, cand_a, ends, stream);
	case kAVG: return launch_search_cd<DXT, kAVG>(nrandom, v, cand_c, cand_a, ends, stream);
	case kWAVG: return launch_search_cd<DXT, kWAVG>(nrandom, v, cand_c, cand_a, ends, stream);
	case kW0AVG: return launch_search_cd<DXT, kW0AVG>(nrandom, v, cand_c, cand_a, ends, stream);
	case kNORMALMAP: return launch_search_cd<DXT, kNORMALMAP>(nrandom, v, cand_c, cand_a, ends, stream);
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t launch_pair_search(int dxt, int cd, int nrandom, const ImageView &v, const uint16_t *d_cand_c,
		const uint8_t *d_cand_a, uint2 *d_ends, cudaStream_t stream)
{
	switch (dxt) {
	case kDxt1: return launch_search_dxt<kDxt1>(cd, nrandom, v, d_cand_c, d_cand_a, d_ends, stream);
	case kDxt3: return launch_search_dxt<kDxt3>(cd, nrandom, v, d_cand_c, d_cand_a, d_ends, stream);
	default: return launch_search_dxt<kDxt5>(cd, nrandom, v, d_cand_c, d_cand_a, d_ends, stream);
	}
}

S2TC_DEFINE_LUT_INIT(init_luts_search)

} // namespace s2tc
