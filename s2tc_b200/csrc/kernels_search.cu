// kernels_search.cu -- MODE_NORMAL with random candidates (S2TC_RANDOM_COLORS > 0): candidate generation and the
// c0/c1 pair search (reference s2tc_algorithm.cpp:962-993, reduce_colors_inplace and reduce_colors_inplace_2fixpoints
// :367-478, reached from :1004-1006).  One warp owns a chunk of 32 consecutive 4x4 blocks and works through them.
//
// Per block the reference builds dists[m][n] (m = n gathered colours + nrandom random candidates, n <= 16) and scans all
// m(m-1)/2 pairs for the smallest sum_k min(d[i][k], d[j][k]): 3160 pairs x 16 texels at nrandom = 64, > 90 % of its
// run time.  Here, per block:
//   1. gather (one ballot), bounding box (warp reductions);  a block whose gathered colours are all equal has an all-zero
//      matrix and the reference keeps pair (0, 1): answered at once;
//   2. the random candidates come straight from the warp's own copy of glibc's rand() state: lane l holds word l of the
//      31-word window, and the recurrence q[i+31] = q[i] + q[i+28] yields the next 31 values as prefix sums along the
//      three stride-3 chains of the window (shuffle steps).  A chunk's window at its first draw is computed by
//      rand_windows_kernel (polynomial jump-ahead, glibc_rand.cuh); a block always consumes 3 (DXT5: 4) x nrandom draws;
//   3. the exact distance matrix in shared memory, one row per candidate: 16-bit halves for the metrics that fit
//      (AVG, WAVG, W0AVG <= 20681; alpha <= 65025), 32-bit rows otherwise;
//   4. the PRUNED pair scan (below): a 4-instruction lower bound per pair, exact sums only for the pairs that survive it;
//   5. DXT5 repeats 3-4 for alpha with the two fixed points 0 and 255 folded into every row (min(d[i][k], f[k]) is
//      stored, so the pair scan is unchanged).
// Output: the chosen endpoints, 8 bytes per block, written once per chunk; kernels_finish.cu turns them into DXT blocks.
//
// History of this kernel on config 3 (16.7 M DXT1 blocks, WAVG, nrandom = 64): 198 ms (32-bit rows) -> 52.1 ms (round 1:
// exact scan in 16x16 tiles, 8 VIMNMX.U16x2 + 8 IDP.2A per pair, 90 % issue utilisation, + 5.3 ms for a separate
// candidate kernel) -> 39.2 ms (pruned scan, candidates generated here) -> 36.7 ms (survivor groups, the rand() ring in
// its own memory) -> 33.9 ms (diagonal tiles as 4 pairs per lane); the steps and what measured slower are in DESIGN.md 5.1.
#define S2TC_USE_SRGB_MIXED_LUT
#include "kernels.cuh"

namespace s2tc {

constexpr int kSearchThreads = 32; // one warp per CTA: a CTA slot frees as soon as its chunk is done (chunks differ a lot in cost)
constexpr int kSearchWarps = kSearchThreads / 32;
constexpr int kPitch16 = 8;  // words per row, 16-bit distances: rows are only read whole (quantisation, exact sums of survivors)
constexpr int kPitch32 = 20; // words per row, 32-bit distances (16 used)
constexpr int kListCap = 160; // survivor groups waiting for their exact sums: < 8 carried over + at most 128 from one pass of two tile columns
constexpr uint32_t kFlushGroups = 8; // 8 groups = 32 pairs = one exact sum per lane
#ifndef S2TC_PS_FLUSH
#define S2TC_PS_FLUSH 0
#endif
#ifndef S2TC_PS_TWOCOL
#define S2TC_PS_TWOCOL 0
#endif
#ifndef S2TC_PS_MINCTAS
#define S2TC_PS_MINCTAS 32
#endif
#ifndef S2TC_PS_MINFLUSH
#define S2TC_PS_MINFLUSH 1 // groups that must be waiting for a flush at a column end that is not the last one
#endif
#ifndef S2TC_PS_DIAG
#define S2TC_PS_DIAG 1 // diagonal tiles as 4 pairs per lane (circular offsets) instead of 8 rows x 16 columns with half the slots void
#endif
#ifndef S2TC_PS_PREFETCH
#define S2TC_PS_PREFETCH 0 // 1: the next block's rand() draws are generated inside this block's tile loop (measured slower, see pair_search_kernel)
#endif
constexpr int kRing = 256;    // rand() outputs kept per warp (a batch of 32 candidates needs <= 128 + 61)

template <int CD> struct Packs16 { static constexpr bool value = CD == kAVG || CD == kWAVG || CD == kW0AVG; };

// per-warp shared memory, byte offsets (all multiples of 16).  One region is used twice: the survivor list lives where the
// metric features were (the scan starts after the fill).  The rand() ring has its own kilobyte since the draws of the NEXT
// block are generated while this block's rows are being scanned.
struct WarpLayout {
	int rows_cap; // matrix rows: m rounded up to whole 16-row tiles
	uint32_t texels, rows, q8, cneg, col, feat, ring, misc, total;
};
__host__ __device__ inline WarpLayout warp_layout(int mcap, bool pack16)
{
	WarpLayout L;
	L.rows_cap = (mcap + 15) & ~15;
	uint32_t o = 0;
	L.texels = o; o += 64;                                                       // the current block's texels, reduced
	uint32_t rows = (uint32_t) L.rows_cap * (pack16 ? kPitch16 : kPitch32) * 4;
	L.rows = o; o += rows;                                                       // exact distance rows
	L.q8 = o; o += (uint32_t) L.rows_cap * 16;                                   // quantised rows, one byte per texel
	L.cneg = o; o += (uint32_t) L.rows_cap * 4;                                  // K - row sum of the quantised row
	L.col = o; o += ((uint32_t) mcap * 4 + 15) & ~15u;                           // candidate colours
	uint32_t feat = ((uint32_t) mcap * (pack16 ? 8 : 12) + 15) & ~15u;
	if (feat < kListCap * 4)
		feat = kListCap * 4;
	L.feat = o; o += feat;                                                       // metric features | survivor list
	L.ring = o; o += kRing * 4;                                                  // rand() outputs, indexed by draw number mod kRing
	L.misc = o; o += 16;
	L.total = o;
	return L;
}

// one texel row of a block as reduced texels (zeros outside the image); every index a compile-time constant
__device__ __forceinline__ void load_block_row(const ImageView &v, int x0, int y, int w, uint32_t (&t)[4])
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
#pragma unroll
			for (int x = 0; x < 4; ++x)
				if (x < w)
					t[x] = __ldg(reinterpret_cast<const uint32_t *>(row) + x);
		}
		if (v.fmt == kSrcRGBA8) {
#pragma unroll
			for (int x = 0; x < 4; ++x)
				t[x] = x < w ? reduce_word(t[x], v.alphabits) : 0u;
		}
	} else {
		const size_t pitch = (size_t) v.width * 3;
		const uint8_t *row = v.base + (size_t) y * pitch + (size_t) x0 * 3;
		const uint32_t ones = ((1u << v.alphabits) - 1u) << 24;
#pragma unroll
		for (int x = 0; x < 4; ++x)
			if (x < w) {
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

// ---- pruned scan ---------------------------------------------------------------------------------------
// The exact pair scan costs 16 min + 16 add per pair.  Most pairs are nowhere near the minimum, and a cheap LOWER
// BOUND of a pair's sum shows it:  with q[i][k] = min(255, d[i][k] >> s)  (one byte per texel, 16 bytes per row)
//     sum_k min(d[i][k], d[j][k])  >=  2^s * sum_k min(q[i][k], q[j][k])  =  2^s * (Rq[i] + Rq[j] - SAD(q[i], q[j])) / 2,
// Rq = row sums, SAD = sum of absolute byte differences: four VABSDIFF4.U8.ACC per pair, 16 texels in 4 instructions.
// A pair can only beat (or tie) the best exact sum T found so far if its bound is <= T; everything else is skipped
// without ever being evaluated exactly.  The result is the reference's: every pair whose exact sum could be the first
// minimum in (i, j) order is evaluated exactly and compared by (sum, i, j).
//   0. 32 sample pairs among the first 16 rows (the block's own colours) are evaluated exactly: a first T, and from T
//      the shift s (values above T never matter, so 8 bits cover [0, T] as finely as they can);
//   1. every lane quantises rows: q bytes and c[i] = K - Rq[i];
//   2. 16x16 tiles: lane (jj, half) keeps the q row j = 16b + jj in registers and meets the eight rows i = 16a + 8 half + t
//      (broadcast loads).  acc = c[i] + SAD(q[i], q[j]) = K + Rq[j] - 2 bound, so the pair survives iff
//      acc >= K + Rq[j] - 2 (T >> s), a per-lane constant: one test per four pairs on their maximum;
//   3. surviving GROUPS of four pairs (i .. i+3, j) -- the unit of the test -- are appended to a list (rare); when 32
//      pairs are waiting, or at the end of a tile column, the warp sums them exactly, one pair per lane, from the
//      full-precision rows and agrees on the new (T, i, j).
// Diagonal tiles go through the same loop (groups that start at i >= j never enter the list, pairs with i >= j are
// skipped at the flush); rows
// beyond m are all-255 (bound = Rq[i], never better than a real pair of row i).
// Measured on a B200 (tools_lab/ubench_sad.cu): VABSDIFF4 issues every other clock per scheduler on the ALU pipe,
// so the bound costs ~10 clocks per pair against ~21 for the exact form, and the other pipes stay free for the rest.
constexpr uint32_t kBoundBias = 4096; // K > 16 * 255
constexpr uint32_t kDiagGroup = 0x80000000u; // list entry: a group of a diagonal tile (see the flush)

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ uint32_t sad_row(const uint4 &a, const uint4 &b, uint32_t c)
{
	return sad4(a.w, b.w, sad4(a.z, b.z, sad4(a.y, b.y, sad4(a.x, b.x, c))));
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

struct BestPair { // lexicographic (sum, i << 16 | j): the reference's first minimum
	uint32_t sum, ij;
	__device__ __forceinline__ void take(uint32_t s, uint32_t p)
	{
		if (s < sum || (s == sum && p < ij)) {
			sum = s;
			ij = p;
		}
	}
	__device__ __forceinline__ void warp_min() // two REDUX: the smallest sum, then the first pair that has it
	{
		const uint32_t smin = __reduce_min_sync(0xFFFFFFFFu, sum);
		ij = __reduce_min_sync(0xFFFFFFFFu, sum == smin ? ij : 0xFFFFFFFFu);
		sum = smin;
	}
};

// 32 pairs i < j < 16 spread over the 120 of the first tile, (0, 1) first
__device__ const uint32_t kSamplePairs[32] = {0x00001u, 0x00004u, 0x00008u, 0x0000Cu, 0x10002u, 0x10006u, 0x1000Au, 0x1000Du,
		0x20004u, 0x20008u, 0x2000Cu, 0x30004u, 0x30008u, 0x3000Bu, 0x3000Fu, 0x40008u, 0x4000Cu, 0x50006u, 0x5000Au, 0x5000Du, 0x60008u,
		0x6000Cu, 0x70008u, 0x7000Cu, 0x80009u, 0x8000Cu, 0x9000Au, 0x9000Eu, 0xA000Du, 0xB000Du, 0xC000Eu, 0xE000Fu};

// Returns (i << 16) | j of the winner in every lane.  rows: exact distance rows (16-bit packed, pitch kPitch16, or 32-bit,
// pitch kPitch32; all values >= 0, columns >= n zero); q8 / cneg: workspace for 16 * ntile quantised rows; list: kListCap
// words; cnt: one word, zero on entry and on exit.
// bg(): warp-uniform background work issued once per tile (the kernel advances its rand() stream there: shuffles and
// multiply-adds whose latency the VABSDIFF4 stream hides, on pipes the tile loop leaves idle)
template <bool PACK16, class BG>
__device__ __forceinline__ uint32_t pruned_search(const uint32_t *rows, uint4 *q8, uint32_t *cneg, uint32_t *list, uint32_t *cnt,
		int m, int lane, int sadj, BG bg)
{
	const int ntile = (m + 15) >> 4;
	const int jj = lane & 15, half = lane >> 4;
	BestPair best{0xFFFFFFFFu, 0xFFFFFFFFu};
	{ // 0. sample pairs
		const uint32_t p = kSamplePairs[lane]; // (hoisting this load out of the block loop costs a register: a 12-byte spill, +0.3 %)
		const int i = (int) (p >> 16), j = (int) (p & 0xFFFFu);
		if (j < m)
			best.take(exact_pair_sum<PACK16>(rows, i, j), p);
		best.warp_min();
	}
	if (best.sum == 0 && best.ij == 1u) // pair (0, 1) with sum 0 cannot be beaten
		return 1u;
	// 1. quantise
	const int s = max(0, 32 - __clz(best.sum) - 8 + sadj);
	for (int r = lane; r < 16 * ntile; r += 32) {
		uint4 q = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
		if (r < m) {
			uint32_t h[8];
			if constexpr (PACK16) {
				RowRegs<true> rr;
				rr.load(rows, r);
				if (s == 0) {
#pragma unroll
					for (int w = 0; w < 8; ++w)
						h[w] = __vminu2(rr.w[w], 0x00FF00FFu);
				} else {
					const uint32_t mask = (0xFFFFu >> s) * 0x10001u;
#pragma unroll
					for (int w = 0; w < 8; ++w)
						h[w] = __vminu2((rr.w[w] >> s) & mask, 0x00FF00FFu);
				}
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
		cneg[r] = kBoundBias - sad_row(q, make_uint4(0u, 0u, 0u, 0u), 0u);
	}
	__syncwarp();
	// 2./3. bound scan; survivors to the list, exact evaluation in batches.  The list holds GROUPS of four pairs
	// (i .. i+3, j) that share one test: a lane that finds a group appends one word with one atomic, and the flush sums
	// all four pairs of every listed group exactly, one pair per lane (the ones that did not survive cost nothing -- the
	// lanes would be idle -- and cannot win: their bound, hence their sum, is above T).  Round 2a appended pair by pair:
	// four compiler-aggregated atomics per event, ~540 instructions per block in the survivor path (ncu source view).
	uint32_t tq2 = (best.sum >> s) * 2u;
	bool pending = false; // this lane has appended since the list was last emptied
	auto append = [&](uint32_t top, int thr, bool jok, int ig, int j) {
		if ((int) top >= thr && jok && ((S2TC_PS_DIAG && !S2TC_PS_TWOCOL) || ig < j)) { // (tiles below the diagonal: every i < j)
			list[atomicAdd(cnt, 1u)] = ((uint32_t) ig << 16) | (uint32_t) j;
			pending = true;
		}
	};
	auto flush_if = [&](bool last_of_column, bool last_of_all) {
		if (__any_sync(0xFFFFFFFFu, pending)) {
			__syncwarp();
			const uint32_t c = *cnt;
			// S2TC_PS_FLUSH 0: at every column end (T as fresh as possible); 1: after the first column, at the end, and
			// whenever 32 pairs are waiting
			const bool now = c >= kFlushGroups || last_of_all || (S2TC_PS_FLUSH == 0 && last_of_column && c >= S2TC_PS_MINFLUSH);
			if (now) {
				__syncwarp(); // every lane has read the count
				if (lane == 0)
					*cnt = 0;
				for (uint32_t idx = lane; idx < 4u * c; idx += 32) {
					const uint32_t e = list[idx >> 2], t = idx & 3u;
					const uint32_t base = (e >> 16) & 0x7FFFu;
					// ordinary group: rows base .. base + 3; diagonal-tile group (flag): rows base .. base + 3 inside the
					// tile's 16 rows, wrapping around
					const int i = (e & kDiagGroup) ? (int) ((base & ~15u) | ((base + t) & 15u)) : (int) (base + t), j = (int) (e & 0xFFFFu);
					const int lo = min(i, j), hi = max(i, j);
					if (lo < hi && hi < m)
						best.take(exact_pair_sum<PACK16>(rows, lo, hi), ((uint32_t) lo << 16) | (uint32_t) hi);
				}
				best.warp_min();
				tq2 = (best.sum >> s) * 2u;
				pending = false;
				__syncwarp(); // list consumed and count reset before anyone appends again
			}
		}
	};
#if S2TC_PS_TWOCOL
	// two tile columns per pass over the rows i: every broadcast row load meets two rows j
	for (int b = 0; b < ntile; b += 2) {
		const bool has1 = b + 1 < ntile; // warp-uniform
		const int j0 = 16 * b + jj, j1 = has1 ? j0 + 16 : j0;
		const uint4 r0 = q8[j0], r1 = q8[j1];
		const uint32_t rqk0 = 2u * kBoundBias - cneg[j0], rqk1 = 2u * kBoundBias - cneg[j1]; // K + Rq[j]
		const bool jok0 = j0 < m, jok1 = has1 && j1 < m;
		const int alast = has1 ? b + 1 : b;
		for (int a = 0; a <= alast; ++a) {
			bg();
			const int i0 = 16 * a + 8 * half;
			const int thr0 = (int) (rqk0 - tq2), thr1 = (int) (rqk1 - tq2);
			const uint4 *qi = q8 + i0;
			const bool do0 = a <= b; // warp-uniform: column b has no tile below its diagonal
#pragma unroll
			for (int g = 0; g < 8; g += 4) {
				const uint4 cc = *reinterpret_cast<const uint4 *>(cneg + i0 + g);
				const uint4 x0 = qi[g], x1 = qi[g + 1], x2 = qi[g + 2], x3 = qi[g + 3];
				if (do0) {
					const uint32_t t0 = sad_row(x0, r0, cc.x), t1 = sad_row(x1, r0, cc.y), t2 = sad_row(x2, r0, cc.z), t3 = sad_row(x3, r0, cc.w);
					append(max(max(t0, t1), max(t2, t3)), thr0, jok0, i0 + g, j0);
				}
				if (has1) {
					const uint32_t t0 = sad_row(x0, r1, cc.x), t1 = sad_row(x1, r1, cc.y), t2 = sad_row(x2, r1, cc.z), t3 = sad_row(x3, r1, cc.w);
					append(max(max(t0, t1), max(t2, t3)), thr1, jok1, i0 + g, j1);
				}
			}
			flush_if(a == alast, a == alast && (b == 0 || b + 2 >= ntile));
		}
	}
#else
	for (int b = 0; b < ntile; ++b) {
		const int j = 16 * b + jj;
		const uint4 rj = q8[j];
		const uint32_t rqk = 2u * kBoundBias - cneg[j]; // K + Rq[j]
		const bool jok = j < m;
#if S2TC_PS_DIAG
		const int atop = b; // tiles a < b in full; the diagonal tile below
#else
		const int atop = b + 1;
#endif
		for (int a = 0; a < atop; ++a) {
			bg();
			const int i0 = 16 * a + 8 * half;
			const int thr = (int) (rqk - tq2); // survives iff acc >= K + Rq[j] - 2 (T >> s); may be negative
			const uint4 *qi = q8 + i0;
			const uint4 ca = *reinterpret_cast<const uint4 *>(cneg + i0), cb = *reinterpret_cast<const uint4 *>(cneg + i0 + 4);
			uint32_t acc[8];
			acc[0] = sad_row(qi[0], rj, ca.x);
			acc[1] = sad_row(qi[1], rj, ca.y);
			acc[2] = sad_row(qi[2], rj, ca.z);
			acc[3] = sad_row(qi[3], rj, ca.w);
			acc[4] = sad_row(qi[4], rj, cb.x);
			acc[5] = sad_row(qi[5], rj, cb.y);
			acc[6] = sad_row(qi[6], rj, cb.z);
			acc[7] = sad_row(qi[7], rj, cb.w);
#pragma unroll
			for (int g = 0; g < 8; g += 4)
				append(max(max(acc[g], acc[g + 1]), max(acc[g + 2], acc[g + 3])), thr, jok, i0 + g, j);
			flush_if(a == b, a == b && (b == 0 || b == ntile - 1));
		}
#if S2TC_PS_DIAG
		{ // The diagonal tile: 120 pairs among 16 rows.  In the 8 x 16 shape above half of the slots hold i >= j.  Instead lane
		  // (jj, half) meets the four rows at circular offsets 1 + 4 half .. 4 + 4 half from its own row j: every unordered
		  // pair is met once (circular distance 1..7) or twice (distance 8: the same exact sum twice, harmless) -- 16 VABSDIFF4
		  // per lane instead of 32.  The partner rows differ per lane (plain loads, not broadcasts); rows beyond m are
		  // filtered at the flush.
			bg();
			const int thr = (int) (rqk - tq2);
			const uint32_t o0 = (uint32_t) (jj + 1 + 4 * half);
			const int tb = 16 * b;
			uint32_t acc[4];
#pragma unroll
			for (int t = 0; t < 4; ++t) {
				const int i = tb + (int) ((o0 + t) & 15u);
				acc[t] = sad_row(q8[i], rj, cneg[i]);
			}
			const uint32_t top = max(max(acc[0], acc[1]), max(acc[2], acc[3]));
			if ((int) top >= thr && jok) {
				list[atomicAdd(cnt, 1u)] = kDiagGroup | ((uint32_t) (tb + (int) (o0 & 15u)) << 16) | (uint32_t) j;
				pending = true;
			}
			flush_if(true, b == 0 || b == ntile - 1);
		}
#endif
	}
#endif
	__syncwarp();
	return best.ij;
}

// Generic scan: any row width, any sum range, sums may wrap negative (SRGB).  Returns (i << 16) | j of the winner in every lane.
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
		rj.load(rows, j); // rows beyond m are inside the allocation and never accepted
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

// ---- the warp's rand() stream ------------------------------------------------------------------------------
// floor((2^32 - 1) / len) for len = 1 .. 256 (filled at compile time)
struct RcpTable {
	uint32_t v[257];
	constexpr RcpTable() : v()
	{
		v[0] = 0;
		for (uint32_t i = 1; i <= 256; ++i)
			v[i] = 0xFFFFFFFFu / i;
	}
};
__constant__ RcpTable kRcp = RcpTable(); // the index (a channel's box length) is warp-uniform: one constant-bank read

// x % len for x < 2^31 with rcp = floor((2^32 - 1) / len): one multiply-high, one correction
__device__ __forceinline__ uint32_t mod_small(uint32_t x, uint32_t len, uint32_t rcp)
{
	uint32_t r = x - __umulhi(x, rcp) * len;
	if (r >= len)
		r -= len;
	return r;
}

// win: word `lane` of the window q[e .. e+30] (lane 31 unused).  Advances the window by 31 positions and returns the lane's
// new word q[e+31+lane]; the rand() outputs are these words >> 1 in lane order (glibc_rand.cuh: GlibcRand::next).
//   new[l] = w[l] + new[l-3] (l >= 3),  new[l] = w[l] + w[28+l] (l < 3)   ==>   prefix sums along the stride-3 chains
// one: the integer 1 as a kernel argument, so that "x += t if lane >= off" stays one IMAD (t * flag + x) on the FMA pipe
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
struct RandLane {
	uint32_t f0, f3, f6, f12, f24; // 1 where the lane takes part in the step, else 0
	int src0;
	__device__ __forceinline__ RandLane(int lane, uint32_t one)
	{
		f0 = lane < 3 ? one : 0u;
		f3 = lane >= 3 ? one : 0u;
		f6 = lane >= 6 ? one : 0u;
		f12 = lane >= 12 ? one : 0u;
		f24 = lane >= 24 ? one : 0u;
		src0 = (lane + 28) & 31;
	}
	__device__ __forceinline__ uint32_t step31(uint32_t win) const
	{
		uint32_t x = mad_u32(__shfl_sync(0xFFFFFFFFu, win, src0), f0, win);
		x = mad_u32(__shfl_up_sync(0xFFFFFFFFu, x, 3), f3, x);
		x = mad_u32(__shfl_up_sync(0xFFFFFFFFu, x, 6), f6, x);
		x = mad_u32(__shfl_up_sync(0xFFFFFFFFu, x, 12), f12, x);
		x = mad_u32(__shfl_up_sync(0xFFFFFFFFu, x, 24), f24, x);
		return x;
	}
};

template <int DXT, int CD>
__global__ void __launch_bounds__(kSearchThreads, S2TC_PS_MINCTAS)
pair_search_kernel(ImageView v, int nrandom, int mcap, int sadj, uint32_t one, const uint32_t *__restrict__ windows, unsigned nchunks,
		uint2 *__restrict__ ends)
{
	typedef Metric<CD> M;
	typedef typename M::Feat Feat;
	constexpr bool kPack = Packs16<CD>::value;
	constexpr int kPitch = kPack ? kPitch16 : kPitch32;
	constexpr int kDraws = DXT == kDxt5 ? 4 : 3; // per candidate: r, g, b [, a] (ref :986-990)
	extern __shared__ __align__(16) uint8_t smem[];
	const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned chunk = blockIdx.x * kSearchWarps + wi;
	if (chunk >= nchunks)
		return;
	const int nblocks = v.blocks_w * v.blocks_h;
	const int b0 = (int) chunk * kSearchChunkBlocks;
	const int nb = min(kSearchChunkBlocks, nblocks - b0);

	const WarpLayout L = warp_layout(mcap, kPack);
	uint8_t *wbase = smem + (size_t) wi * L.total;
	uint32_t *texels = reinterpret_cast<uint32_t *>(wbase + L.texels); // [16], texel y * 4 + x of the current block
	uint32_t *rows = reinterpret_cast<uint32_t *>(wbase + L.rows);
	uint4 *q8 = reinterpret_cast<uint4 *>(wbase + L.q8);
	uint32_t *cneg = reinterpret_cast<uint32_t *>(wbase + L.cneg);
	uint32_t *col = reinterpret_cast<uint32_t *>(wbase + L.col);
	Feat *feat = reinterpret_cast<Feat *>(wbase + L.feat);
	uint32_t *list = reinterpret_cast<uint32_t *>(wbase + L.feat);
	uint32_t *ring = reinterpret_cast<uint32_t *>(wbase + L.ring);
	uint32_t *cnt = reinterpret_cast<uint32_t *>(wbase + L.misc);

	// texel rows are fetched one block ahead: lane y < 4 holds row y of the next block
	const int by0 = b0 / v.blocks_w, bx0 = b0 - by0 * v.blocks_w;
	uint32_t nextrow[4] = {0u, 0u, 0u, 0u};
	if (lane < 4)
		load_block_row(v, bx0 * 4, by0 * 4 + lane, min(4, v.width - bx0 * 4), nextrow);
	if (lane == 0)
		*cnt = 0;
	uint32_t win = lane < kLag ? __ldg(windows + (size_t) lane * nchunks + chunk) : 0u;
	const RandLane rl(lane, one);
	int gen = 0; // draws of this chunk generated so far
	int pos = 0; // first draw of the current block
	uint2 result = make_uint2(0u, 0u);
	__syncwarp();

	// Prefetch (S2TC_PS_PREFETCH, off): all candidates of a block are built before its first search, so from then on the ring
	// is free for the draws of the NEXT block -- one rand() step per tile, five dependent shuffle + multiply-add pairs whose
	// latency (this phase was 9 % of the instructions but 15 % of the stall samples, ncu source view r02b) could hide behind
	// the VABSDIFF4 stream.  Measured on a B200 it costs more than it hides: 9.49 ms against 9.12 ms per 8192^2 slab (the
	// tile loop is at the ALU pipe's ceiling and pays for every extra test and live register).  The ring must still hold the
	// next block's first draw when that block starts: at most kRing - kLag ahead.  What did pay was the ring's own kilobyte
	// (it used to share the rows' memory: a put-back of the latest window and a range test per step): 9.41 -> 9.12 ms.
	const int draws = kDraws * nrandom;
	int bi = 0;
	auto prefetch = [&]() {
#if S2TC_PS_PREFETCH
		if (bi + 1 < nb && gen < pos + 2 * draws && gen <= pos + draws + (kRing - kLag)) {
			win = rl.step31(win);
			if (lane < kLag)
				ring[(gen + lane) & (kRing - 1)] = win >> 1;
			gen += kLag;
		}
#endif
	};
	int bx = bx0, by = by0;
	for (; bi < nb; ++bi, pos += draws) {
		const int w = min(4, v.width - bx * 4), h = min(4, v.rows - by * 4);
		if (++bx == v.blocks_w) {
			bx = 0;
			++by;
		}
		__syncwarp(); // the previous block's reads of `texels` (a single-colour block leaves the loop body without another barrier)
		if (lane < 4) {
			*reinterpret_cast<uint4 *>(texels + lane * 4) = make_uint4(nextrow[0], nextrow[1], nextrow[2], nextrow[3]);
			if (bi + 1 < nb)
				load_block_row(v, bx * 4, by * 4 + lane, min(4, v.width - bx * 4), nextrow);
		}
		__syncwarp();
		// 1. gather in the reference's column-major order (ref :940-959): lane o < 16 looks at texel (x, y) = (o >> 2, o & 3)
		const uint32_t valid = (w == 4 && h == 4) ? 0xFFFFu : valid_mask(w, h);
		const int ti = (lane & 3) * 4 + ((lane >> 2) & 3);
		const uint32_t mine = texels[ti];
		bool use = lane < 16 && ((valid >> ti) & 1u);
		if (DXT == kDxt1)
			use = use && (mine >> 24) != 0;
		const uint32_t usemask = __ballot_sync(0xFFFFFFFFu, use);
		int n = __popc(usemask);
		const uint32_t first = usemask ? __shfl_sync(0xFFFFFFFFu, mine, __ffs(usemask) - 1) : 0u; // n == 0: black, alpha 0 (ref :952-959)
		const bool flat_c = !__any_sync(0xFFFFFFFFu, use && ((mine ^ first) & 0x00FFFFFFu));
		const bool flat_a = DXT != kDxt5 || !__any_sync(0xFFFFFFFFu, use && ((mine ^ first) >> 24));
		// All gathered colours equal: the box has one colour, every candidate equals it, the matrix is zero and the reference
		// keeps pair (0, 1) -- both endpoints are that colour.  Likewise for alpha.  The block's draws are still consumed.
		uint32_t c01 = to565(first) * 0x10001u, a01 = (first >> 24) * 0x101u;
		if (!(flat_c && flat_a)) {
			if (use)
				col[__popc(usemask & ((1u << lane) - 1u))] = mine;
			if (n == 0) {
				if (lane == 0)
					col[0] = 0;
				n = 1;
			}
			const int m = n + nrandom;
			{ // 2. candidates (ref :962-993): lo + rand() % len per channel over the box of the gathered colours
				uint32_t lo[4], len[4], rcp[4];
#pragma unroll
				for (int ch = 0; ch < kDraws; ++ch) {
					const uint32_t val = (mine >> (8 * ch)) & 0xFFu;
					lo[ch] = usemask ? __reduce_min_sync(0xFFFFFFFFu, use ? val : 255u) : 0u;
					const uint32_t hi = usemask ? __reduce_max_sync(0xFFFFFFFFu, use ? val : 0u) : 0u;
					len[ch] = hi - lo[ch] + 1u;
					rcp[ch] = kRcp.v[len[ch]];
				}
				// ring[d mod kRing] = draw d of the chunk; whatever of this block's draws the previous block's tile loop has
				// not generated already (prefetch, below) is generated here
				for (int k0 = 0; k0 < nrandom; k0 += 32) {
					const int need = pos + min(nrandom, k0 + 32) * kDraws;
					while (gen < need) {
						win = rl.step31(win);
						if (lane < kLag)
							ring[(gen + lane) & (kRing - 1)] = win >> 1;
						gen += kLag;
					}
					__syncwarp();
					const int k = k0 + lane;
					if (k < nrandom) {
						const int d0 = pos + k * kDraws;
						uint32_t c = 0;
#pragma unroll
						for (int ch = 0; ch < kDraws; ++ch)
							c |= (lo[ch] + mod_small(ring[(d0 + ch) & (kRing - 1)], len[ch], rcp[ch])) << (8 * ch);
						col[n + k] = c;
					}
					__syncwarp();
				}
			}

			// 3. distance matrix (ref :375-392; argument order matters for SRGB), columns >= n zero
			if (!flat_c) {
				if constexpr (kPack) {
					// AVG / WAVG / W0AVG: the feature is the colour with pre-scaled channel bytes f (all < 128);
					// d(i, k) = |f_i|^2 + |f_k|^2 - 2 f_i . f_k = two IDP.4A with the negated bytes of f_k and one add
					uint32_t *cvec = reinterpret_cast<uint32_t *>(feat), *nrm = cvec + mcap;
					for (int i = lane; i < m; i += 32) {
						const uint32_t f = M::feat(col[i]).v;
						cvec[i] = f;
						nrm[i] = (uint32_t) __dp4a((int) f, (int) f, 0);
					}
					__syncwarp();
					// lane -> texel columns 4 kq .. 4 kq + 3 of one row, 8 rows per step
					const int kq = lane & 3;
					uint32_t nf[4], nk[4];
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const int k = 4 * kq + c; // k < 16 <= capacity; columns >= n are masked below
						nf[c] = (0x80808080u - cvec[k]) ^ 0x80808080u; // per-byte negation (no borrow: bytes < 128)
						nk[c] = nrm[k];
					}
					const uint32_t m01 = (4 * kq < n ? 0x0000FFFFu : 0u) | (4 * kq + 1 < n ? 0xFFFF0000u : 0u);
					const uint32_t m23 = (4 * kq + 2 < n ? 0x0000FFFFu : 0u) | (4 * kq + 3 < n ? 0xFFFF0000u : 0u);
					for (int i = lane >> 2; i < m; i += 8) {
						const int fi = (int) cvec[i];
						const uint32_t ni = nrm[i];
						uint32_t d[4];
#pragma unroll
						for (int c = 0; c < 4; ++c)
							d[c] = (uint32_t) __dp4a(fi, (int) nf[c], __dp4a(fi, (int) nf[c], (int) (ni + nk[c])));
						*reinterpret_cast<uint2 *>(rows + i * kPitch + 2 * kq) = make_uint2((d[0] | (d[1] << 16)) & m01, (d[2] | (d[3] << 16)) & m23);
					}
				} else {
					for (int i = lane; i < m; i += 32)
						feat[i] = M::feat(col[i]);
					__syncwarp();
					// lane -> texel column k, 2 rows per step
					const int k = lane & 15;
					const Feat fk = feat[k];
					const bool kok = k < n;
					for (int i = lane >> 4; i < m; i += 2) {
						int d = 0;
						if (kok && k != i) {
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
					cij = pruned_search<kPack>(rows, q8, cneg, list, cnt, m, lane, sadj, prefetch);
				c01 = to565(col[cij >> 16]) | (to565(col[cij & 0xFFFFu]) << 16);
			}

			if (DXT == kDxt5 && !flat_a) { // 5. ref :416-478; alpha rows are always 16-bit, at the 16-bit pitch inside the same buffer
				__syncwarp();
				{ // the fixed points 0 and 255 folded into every row: min(d[i][k], fix[k]) is stored (masked columns: fix = 0)
					const int kq = lane & 7;
					const uint32_t ak0 = col[2 * kq] >> 24, ak1 = col[2 * kq + 1] >> 24;
					const uint32_t f0 = min(ak0 * ak0, (255u - ak0) * (255u - ak0)), f1 = min(ak1 * ak1, (255u - ak1) * (255u - ak1));
					const uint32_t fixw = (2 * kq < n ? f0 : 0u) | (2 * kq + 1 < n ? f1 << 16 : 0u);
#pragma unroll 2
					for (int i = lane >> 3; i < m; i += 4) {
						const uint32_t ai = col[i] >> 24;
						const uint32_t t0 = ai - ak0, t1 = (ai - ak1) << 8; // wrapping: the squares are exact mod 2^32
						rows[i * kPitch16 + kq] = __vminu2(t1 * t1 + t0 * t0, fixw);
					}
				}
				__syncwarp();
				const uint32_t aij = pruned_search<true>(rows, q8, cneg, list, cnt, m, lane, sadj, prefetch);
				a01 = (col[aij >> 16] >> 24) | ((col[aij & 0xFFFFu] >> 24) << 8);
			}
			__syncwarp();
		}
		if (lane == bi)
			result = make_uint2(c01, a01);
	}
	if (lane < nb)
		ends[b0 + lane] = result;
}

template <int DXT, int CD>
static cudaError_t launch_search_cd(int nrandom, const ImageView &v, const uint32_t *windows, unsigned nchunks, uint2 *ends, cudaStream_t stream)
{
	if (nchunks == 0)
		return cudaSuccess;
	const int mcap = 16 + nrandom;
	const size_t smem = (size_t) warp_layout(mcap, Packs16<CD>::value).total * kSearchWarps;
	auto kern = pair_search_kernel<DXT, CD>;
	if (smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess)
			return e;
	}
	// quantisation shift relative to "the best sum so far just fits 8 bits": one bit finer measured best on config 3
	// (8192^2: 9.80 / 9.61 / 9.61 / 10.31 ms for 0 / -1 / -2 / +1); saturation at 255 keeps every value a lower bound
	static const int sadj = [] { const char *e = getenv("S2TC_B200_SADJ"); return e ? atoi(e) : -1; }();
	const dim3 block(kSearchThreads), grid((nchunks + kSearchWarps - 1) / kSearchWarps);
	kern<<<grid, block, smem, stream>>>(v, nrandom, mcap, sadj, 1u, windows, nchunks, ends);
	return cudaGetLastError();
}

int pair_search_max_nrandom()
{
	// the 32-bit layout is the larger one; 227 KB of shared memory per CTA
	int lo = 0, hi = 1 << 15;
	while (lo < hi) {
		const int mid = (lo + hi + 1) / 2;
		if ((size_t) warp_layout(16 + mid, false).total * kSearchWarps <= 227 * 1024)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

template <int DXT>
static cudaError_t launch_search_dxt(int cd, int nrandom, const ImageView &v, const uint32_t *windows, unsigned nchunks, uint2 *ends,
		cudaStream_t stream)
{
	if (nrandom <= 0 || nrandom > pair_search_max_nrandom())
		return cudaErrorInvalidValue; // nrandom <= 0 is served by the fused 16-candidate encoder (search16.inl)
	switch (cd) {
	case kRGB: return launch_search_cd<DXT, kRGB>(nrandom, v, windows, nchunks, ends, stream);
	case kYUV: return launch_search_cd<DXT, kYUV>(nrandom, v, windows, nchunks, ends, stream);
	case kSRGB: return launch_search_cd<DXT, kSRGB>(nrandom, v, windows, nchunks, ends, stream);
	case kSRGB_MIXED: return launch_search_cd<DXT, kSRGB_MIXED>(nrandom, v, windows, nchunks, ends, stream);
	case kAVG: return launch_search_cd<DXT, kAVG>(nrandom, v, windows, nchunks, ends, stream);
	case kWAVG: return launch_search_cd<DXT, kWAVG>(nrandom, v, windows, nchunks, ends, stream);
	case kW0AVG: return launch_search_cd<DXT, kW0AVG>(nrandom, v, windows, nchunks, ends, stream);
	case kNORMALMAP: return launch_search_cd<DXT, kNORMALMAP>(nrandom, v, windows, nchunks, ends, stream);
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t launch_pair_search(int dxt, int cd, int nrandom, const ImageView &v, const uint32_t *d_windows, uint2 *d_ends,
		cudaStream_t stream)
{
	const long long nblocks = (long long) v.blocks_w * v.blocks_h;
	const unsigned nchunks = (unsigned) ((nblocks + kSearchChunkBlocks - 1) / kSearchChunkBlocks);
	switch (dxt) {
	case kDxt1: return launch_search_dxt<kDxt1>(cd, nrandom, v, d_windows, nchunks, d_ends, stream);
	case kDxt3: return launch_search_dxt<kDxt3>(cd, nrandom, v, d_windows, nchunks, d_ends, stream);
	default: return launch_search_dxt<kDxt5>(cd, nrandom, v, d_windows, nchunks, d_ends, stream);
	}
}

S2TC_DEFINE_LUT_INIT(init_luts_search)

} // namespace s2tc
