// search16.inl -- MODE_NORMAL without random candidates (S2TC_RANDOM_COLORS = 0, or NORMALMAP with
// the default -1): gather, distance matrix and c0/c1 pair search (and, as a second launch of the same kernel,
// the DXT5 alpha search), one thread per 4x4 block, everything in registers.  It writes the chosen
// endpoints; finish_kernel (kernels_finish.cu) refines and packs them.
//
// Reference path per block: s2tc_algorithm.cpp:938-959 (gather), :997-1001 (single-colour hack), :367-414
// (reduce_colors_inplace), :416-478 (reduce_colors_inplace_2fixpoints).
//
// Why this shape: with at most 16 candidates the search is 120 pairs x 16 texels.  A thread keeps the
// distance matrix in registers (all loops fully unrolled, every index a compile-time constant), so the
// scan has no shared memory, no shuffles and no synchronisation; 32 independent blocks per warp give the
// ILP.  Two code paths:
//   * search_full -- every texel of every block of the warp is a candidate (n == 16): the common case,
//     tuned for the pipes (see there);
//   * search_any  -- fewer colours (DXT1 transparency, ragged edges): the same scan with the missing
//     rows/columns zeroed and the missing pairs masked, which is exactly what the reference's smaller
//     loops compute.
// History (profiles/README.md): a 4-lane group per block with the matrix in shared memory issued 2.6 bank
// conflicts per LDS (r01a, 9.7 ms on config 2); the first register-resident version fused refinement and
// packing into the same kernel (3.1 ms): its scans sat at ~90 % ALU-pipe utilisation with an idle FMA
// pipe, and its refinement tail ran at 3 warps per scheduler.  Moving the adds to the FMA pipe, 16-bit
// packed alpha rows and the split into search + finish kernels brought it to 1.05 (colour) + 0.66 (alpha) +
// 0.63 ms (finish); DESIGN.md 5.1 has the steps and what was measured and rejected.
#include <utility>

#define S2TC_USE_SRGB_MIXED_LUT
#include "kernels.cuh"

namespace s2tc {

__host__ __device__ constexpr int tri(int i, int k) { return i * 16 - i * (i + 1) / 2 + (k - i - 1); } // i < k
// pair number p (lexicographic over i < j < 16) -> i, j
__host__ __device__ constexpr int pair_i(int p)
{
	int i = 0;
	while (p >= 15 - i) {
		p -= 15 - i;
		++i;
	}
	return i;
}
__host__ __device__ constexpr int pair_j(int p) { return p - tri(pair_i(p), pair_i(p) + 1) + pair_i(p) + 1; }

// d[i][k] of the symmetric matrix whose upper triangle is D (compile-time indices only)
template <int I, int K>
__device__ __forceinline__ int sym(const int (&D)[120])
{
	if constexpr (I == K)
		return 0;
	else if constexpr (I < K)
		return D[tri(I, K)];
	else
		return D[tri(K, I)];
}

// sum_k min(d[I][k], d[J][k] [, fix[k]]) for one pair; every index is a template constant so that the
// matrix stays in registers (a runtime-indexed scan put it in local memory: 443 LDL per thread)
template <int I, int J, bool SKIP_SELF, bool FIXED, int... K>
__device__ __forceinline__ int pair_sum16(const int (&D)[120], const int (&fix)[16], std::integer_sequence<int, K...>)
{
	uint32_t s0 = 0, s1 = 0;
	auto term = [&](auto kc) {
		constexpr int k = decltype(kc)::value;
		if constexpr (SKIP_SELF && (k == I || k == J)) {
			// d[i][i] = 0 and distances are >= 0: the term is 0
		} else {
			int m = min(sym<I, k>(D), sym<J, k>(D));
			if constexpr (FIXED)
				m = min(m, fix[k]);
			if constexpr (k & 1)
				s1 += (uint32_t) m;
			else
				s0 += (uint32_t) m;
		}
	};
	(term(std::integral_constant<int, K>{}), ...);
	return (int) (s0 + s1);
}

// one step of the reference's scan (ref :393-410): accept if "bestsum < 0 || sum < bestsum"
template <int P, bool MAY_BE_NEGATIVE, bool FIXED>
__device__ __forceinline__ void pair_step(const int (&D)[120], const int (&fix)[16], int n, int &best, uint32_t &bij)
{
	constexpr int I = pair_i(P), J = pair_j(P);
	const int sum = pair_sum16<I, J, !MAY_BE_NEGATIVE, FIXED>(D, fix, std::make_integer_sequence<int, 16>{});
	bool accept;
	if constexpr (MAY_BE_NEGATIVE)
		accept = J < n && (best < 0 || sum < best); // verbatim: sums of a wrapping metric can be negative
	else
		accept = J < n && (uint32_t) sum < (uint32_t) best; // same rule when sums are >= 0 (best starts at -1)
	if (accept) {
		best = sum;
		bij = (uint32_t) (I << 4 | J);
	}
}

template <bool MAY_BE_NEGATIVE, bool FIXED, int... P>
__device__ __forceinline__ uint32_t scan120(const int (&D)[120], const int (&fix)[16], int n, std::integer_sequence<int, P...>)
{
	int best = -1;
	uint32_t bij = 1u; // (0, 1), the reference's initial besti/bestj
	(pair_step<P, MAY_BE_NEGATIVE, FIXED>(D, fix, n, best, bij), ...);
	return bij;
}

// upper triangle of the distance matrix, entries beyond n zeroed
template <class F, int... P>
__device__ __forceinline__ void fill120(int (&D)[120], int n, F dist, std::integer_sequence<int, P...>)
{
	((D[P] = pair_j(P) < n ? dist(std::integral_constant<int, pair_i(P)>{}, std::integral_constant<int, pair_j(P)>{}) : 0), ...);
}

// ---- the common case: every texel of the block is a candidate (n == 16) in every lane of the warp -----------
// No masking anywhere, and two changes that take work off the ALU pipe, which bounds the scans (ncu: ALU pipe ~90 %
// busy inside the scans with the FMA pipe idle):
//  * the 16-term sums accumulate through IMAD (m * one + s, `one` an opaque kernel argument equal to 1) in
//    S2TC_SEARCH16_CHAINS independent chains (measured on config 2: 1 / 2 / 4 / 7 chains -> 1.99 / 1.93 / 1.75 / 1.72 ms)
//    instead of IADD3 only: a pair costs 20 ALU + 7 FMA-pipe instructions instead of 24 ALU;
//  * distances that fit 16 bits (alpha always; AVG / WAVG / W0AVG colours) are stored as full rows of 16-bit halves,
//    the DXT5 fixed points folded in (min(d[i][k], fix[k]) is stored, as in kernels_search.cu): a pair is
//    8 VIMNMX.U16x2 + 8 IDP.2A, and since the sums stay below 2^20 the whole acceptance rule is one unsigned min over
//    keys (sum << 7) | pair_number ("first minimum in scan order", ref :393-410).
#ifndef S2TC_SEARCH16_CHAINS
#define S2TC_SEARCH16_CHAINS 7 // independent accumulation chains per pair sum
#endif
template <int CD> struct Fits16 { static constexpr bool value = CD == kAVG || CD == kWAVG || CD == kW0AVG; };

// t-th texel column that contributes to pair (I, J): every column, or (SKIP) every column but I and J, whose terms are 0
template <int I, int J, bool SKIP>
__host__ __device__ constexpr int kth_col(int t)
{
	if (!SKIP)
		return t;
	int k = 0;
	for (;; ++k) {
		if (k == I || k == J)
			continue;
		if (t-- == 0)
			return k;
	}
}

// a * b + c that the compiler cannot see through (written as a * one + c it factors `one` out of the whole sum)
__device__ __forceinline__ uint32_t mad_opaque(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

template <int I, int J, bool SKIP, int... T>
__device__ __forceinline__ int pair_sum_full(const int (&D)[120], uint32_t one, std::integer_sequence<int, T...>)
{
	uint32_t s[S2TC_SEARCH16_CHAINS];
	auto term = [&](auto tc) {
		constexpr int t = decltype(tc)::value;
		constexpr int k = kth_col<I, J, SKIP>(t);
		const uint32_t m = (uint32_t) min(sym<I, k>(D), sym<J, k>(D));
		constexpr int ch = t % S2TC_SEARCH16_CHAINS;
		if constexpr (t < S2TC_SEARCH16_CHAINS)
			s[ch] = m;
		else
			s[ch] = mad_opaque(m, one, s[ch]);
	};
	(term(std::integral_constant<int, T>{}), ...);
	// the partial sums meet in plain adds (IADD3): a tree of the same multiply-adds measured slower, it lengthens
	// the live ranges until the matrix spills (2.13 vs 1.72 ms on config 2)
	uint32_t r = s[0];
#pragma unroll
	for (int q = 1; q < S2TC_SEARCH16_CHAINS; ++q)
		r += s[q];
	return (int) r;
}

template <int P, bool MAY_BE_NEGATIVE>
__device__ __forceinline__ void pair_step_full(const int (&D)[120], uint32_t one, int &best, uint32_t &bp)
{
	constexpr int I = pair_i(P), J = pair_j(P);
	constexpr bool SKIP = !MAY_BE_NEGATIVE; // d[i][i] = 0 and distances >= 0: the self terms are 0
	const int sum = pair_sum_full<I, J, SKIP>(D, one, std::make_integer_sequence<int, SKIP ? 14 : 16>{});
	if constexpr (MAY_BE_NEGATIVE) {
		if (best < 0 || sum < best) { // verbatim (ref :404)
			best = sum;
			bp = (uint32_t) P;
		}
	} else {
		// (predicated multiply-adds for the two conditional moves, to take them off the ALU pipe, measured slower:
		// 1.09 vs 1.03 ms for the colour launch of config 2)
		if ((uint32_t) sum < (uint32_t) best) { // the same rule for sums >= 0 (best starts at -1)
			best = sum;
			bp = (uint32_t) P;
		}
	}
}

// returns the number of the winning pair in scan order
template <bool MAY_BE_NEGATIVE, int... P>
__device__ __forceinline__ uint32_t scan120_full(const int (&D)[120], uint32_t one, std::integer_sequence<int, P...>)
{
	int best = -1;
	uint32_t bp = 0; // pair 0 = (0, 1), the reference's initial besti/bestj
	(pair_step_full<P, MAY_BE_NEGATIVE>(D, one, best, bp), ...);
	return bp;
}

template <class F, int... P>
__device__ __forceinline__ void fill120_full_impl(int (&D)[120], F dist, std::integer_sequence<int, P...>)
{
	((D[P] = dist(std::integral_constant<int, pair_i(P)>{}, std::integral_constant<int, pair_j(P)>{})), ...);
}

// 16-bit rows: R[i][q] = {d[i][2q], d[i][2q+1]}
template <int P>
__device__ __forceinline__ void pair_step_packed(const uint32_t (&R)[16][8], uint32_t scale, uint32_t &best)
{
	constexpr int I = pair_i(P), J = pair_j(P);
	uint32_t s0 = 0, s1 = 0;
#pragma unroll
	for (int q = 0; q < 8; q += 2) {
		s0 = __dp2a_lo(__vminu2(R[I][q], R[J][q]), 0x0101u, s0);
		s1 = __dp2a_lo(__vminu2(R[I][q + 1], R[J][q + 1]), 0x0101u, s1);
	}
	// key = (s0 + s1) * 128 + P; scale = 128 is opaque so that the two multiply-adds stay on the FMA pipe
	best = min(best, mad_opaque(s0, scale, mad_opaque(s1, scale, (uint32_t) P)));
}

template <int... P>
__device__ __forceinline__ uint32_t scan120_packed(const uint32_t (&R)[16][8], uint32_t scale, std::integer_sequence<int, P...>)
{
	uint32_t best = 0xFFFFFFFFu;
	(pair_step_packed<P>(R, scale, best), ...);
	return best & 127u;
}

// pair number -> (i << 4) | j
__device__ __forceinline__ uint32_t pair_unrank(uint32_t p)
{
	uint32_t i = 0;
	while (p >= 15u - i) {
		p -= 15u - i;
		++i;
	}
	return (i << 4) | (i + 1u + p);
}

// COLOR / ALPHA: which of the two searches to run (the kernel is launched once per search, see search16_kernel)
template <int DXT, int CD, bool COLOR, bool ALPHA>
__device__ __forceinline__ void search_full(const Block &b, uint32_t one, uint32_t &c0, uint32_t &c1, int &a0, int &a1)
{
	typedef Metric<CD> M;
	typedef typename M::Feat Feat;
	uint32_t cl[16]; // the candidates in the reference's column-major order (ref :940-951), addressable for the final reads
#pragma unroll
	for (int o = 0; o < 16; ++o)
		cl[o] = b.px[(o & 3) * 4 + (o >> 2)];
	const uint32_t scale = one << 7;

	c0 = c1 = 0;
	if constexpr (COLOR) {
	uint32_t bij;
	if constexpr (Fits16<CD>::value) {
		uint32_t R[16][8];
#pragma unroll
		for (int i = 0; i < 16; ++i)
#pragma unroll
			for (int q = 0; q < 8; ++q)
				R[i][q] = 0;
		{
			Feat f[16];
#pragma unroll
			for (int k = 0; k < 16; ++k)
				f[k] = M::feat(cl[k]);
			// these metrics are symmetric: every distance is computed once and dropped into both rows
#pragma unroll
			for (int i = 0; i < 16; ++i)
#pragma unroll
				for (int k = i + 1; k < 16; ++k) {
					const uint32_t d = (uint32_t) M::dist(f[i], f[k]);
					R[i][k >> 1] += d << (16 * (k & 1));
					R[k][i >> 1] += d << (16 * (i & 1));
				}
		}
		bij = pair_unrank(scan120_packed(R, scale, std::make_integer_sequence<int, 120>{}));
	} else {
		int D[120];
		{
			Feat f[16];
#pragma unroll
			for (int k = 0; k < 16; ++k)
				f[k] = M::feat(cl[k]);
			// lower index first: SRGB is not symmetric
			fill120_full_impl(D, [&](auto i, auto k) { return M::dist(f[decltype(i)::value], f[decltype(k)::value]); },
					std::make_integer_sequence<int, 120>{});
		}
		bij = pair_unrank(scan120_full<M::kMayBeNegative>(D, one, std::make_integer_sequence<int, 120>{}));
	}
	c0 = px_rgb(cl[bij >> 4]);
	c1 = px_rgb(cl[bij & 15u]);
	}

	a0 = a1 = 0;
	if constexpr (ALPHA && DXT == kDxt5) { // ref :416-478
		uint32_t R[16][8];
		{
			uint32_t a[16], fixw[8];
#pragma unroll
			for (int k = 0; k < 16; ++k)
				a[k] = cl[k] >> 24; // re-read: the registers that held the texels are dead by now
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				const uint32_t x = a[2 * q], y = a[2 * q + 1];
				fixw[q] = min(x * x, (255u - x) * (255u - x)) | (min(y * y, (255u - y) * (255u - y)) << 16);
			}
#pragma unroll
			for (int i = 0; i < 16; ++i) {
#pragma unroll
				for (int q = 0; q < 8; ++q) {
					const uint32_t t0 = a[i] - a[2 * q], t1 = (a[i] - a[2 * q + 1]) << 8; // wrapping: squares are exact mod 2^32
					R[i][q] = __vminu2(t1 * t1 + t0 * t0, fixw[q]);
				}
			}
		}
		const uint32_t aij = pair_unrank(scan120_packed(R, scale, std::make_integer_sequence<int, 120>{}));
		a0 = (int) (cl[aij >> 4] >> 24);
		a1 = (int) (cl[aij & 15u] >> 24);
	}
}

// ---- any n: DXT1 transparency, ragged edges ------------------------------------------------------------------
template <int DXT, int CD, bool COLOR, bool ALPHA>
__device__ __noinline__ void search_any(const Block &b, uint32_t usemask, uint32_t &c0, uint32_t &c1, int &a0, int &a1)
{
	typedef Metric<CD> M;
	typedef typename M::Feat Feat;
	uint32_t c[16];
	uint32_t cl[16]; // the same list, addressable (two dynamic reads after the search)
	int n = 0;
#pragma unroll
	for (int o = 0; o < 16; ++o)
		cl[o] = 0;
#pragma unroll
	for (int o = 0; o < 16; ++o)
		if ((usemask >> o) & 1u)
			cl[n++] = b.px[(o & 3) * 4 + (o >> 2)];
	if (n == 0)
		n = 1; // black, alpha 0 (ref :952-959)
	if (n == 1) { // ref :997-1001 (and see DESIGN.md on the reference's uninitialised ca[1])
		cl[1] = cl[0];
		n = 2;
	}
#pragma unroll
	for (int o = 0; o < 16; ++o)
		c[o] = cl[o];

	// ---- colour distance matrix: the 120 entries above the diagonal (ref :375-383) -----------------
	int D[120];
	int fix[16];
#pragma unroll
	for (int k = 0; k < 16; ++k)
		fix[k] = 0;
	c0 = c1 = 0;
	if constexpr (COLOR) {
		{
			Feat f[16];
#pragma unroll
			for (int k = 0; k < 16; ++k)
				f[k] = M::feat(c[k]);
			// lower index first: SRGB is not symmetric
			fill120(D, n, [&](auto i, auto k) { return M::dist(f[decltype(i)::value], f[decltype(k)::value]); },
					std::make_integer_sequence<int, 120>{});
		}
		// ---- pair scan in lexicographic (i, j) order (ref :393-410) ---------------------------------------
		const uint32_t bij = scan120<M::kMayBeNegative, false>(D, fix, n, std::make_integer_sequence<int, 120>{});
		c0 = px_rgb(cl[bij >> 4]);
		c1 = px_rgb(cl[bij & 15u]);
	}

	// ---- DXT5: the same search on alpha with the fixed points 0 and 255 (ref :416-478) ----------------
	a0 = a1 = 0;
	if constexpr (ALPHA && DXT == kDxt5) {
		int a[16];
#pragma unroll
		for (int k = 0; k < 16; ++k)
			a[k] = (int) (cl[k] >> 24);
#pragma unroll
		for (int k = 0; k < 16; ++k)
			fix[k] = k < n ? min(a[k] * a[k], (255 - a[k]) * (255 - a[k])) : 0;
		fill120(D, n, [&](auto i, auto k) {
			const int d = a[decltype(i)::value] - a[decltype(k)::value];
			return d * d;
		}, std::make_integer_sequence<int, 120>{});
		const uint32_t aij = scan120<false, true>(D, fix, n, std::make_integer_sequence<int, 120>{});
		a0 = (int) (cl[aij >> 4] >> 24);
		a1 = (int) (cl[aij & 15u] >> 24);
	}
}

#ifndef S2TC_SEARCH16_MINBLOCKS
#define S2TC_SEARCH16_MINBLOCKS 3
#endif
constexpr int kSearch16Threads = 128;
// Writes the chosen endpoints of every block, {c0 | c1 << 16 as RGB565, a0 | a1 << 8}, the layout finish_kernel
// reads.  one: the integer 1 (see search_full).
// DXT5 runs it twice, once per search (COLOR writes the first word, ALPHA the second): each launch walks ~100 KB /
// ~65 KB of straight-line code instead of ~150 KB in one kernel.  The alpha launch does not depend on the metric.
template <int DXT, int CD, bool COLOR, bool ALPHA>
__global__ void __launch_bounds__(kSearch16Threads, S2TC_SEARCH16_MINBLOCKS) search16_kernel(ImageView v, uint32_t one, uint2 *__restrict__ ends)
{
	const int nblocks = v.blocks_w * v.blocks_h * v.images;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nblocks)
		return;
	size_t out_off; // endpoints are indexed by the block's number in the whole batch
	const int ti = select_image(v, t, out_off);
	const int by = ti / v.blocks_w, bx = ti - by * v.blocks_w;
	Block b;
	load_block(v, bx, by, b);

	// ---- which texels are candidates, in the reference's column-major order (ref :940-959) ---------------
	uint32_t usemask = 0; // bit o = x*4 + y
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		bool use = (b.valid >> i) & 1u;
		if (DXT == kDxt1)
			use = use && (b.px[i] >> 24) != 0;
		if (use)
			usemask |= 1u << ((i & 3) * 4 + (i >> 2));
	}
	uint32_t c0, c1;
	int a0, a1;
	if (__all_sync(__activemask(), usemask == 0xFFFFu)) // warp-uniform
		search_full<DXT, CD, COLOR, ALPHA>(b, one, c0, c1, a0, a1);
	else
		search_any<DXT, CD, COLOR, ALPHA>(b, usemask, c0, c1, a0, a1);
	const uint32_t cw = to565(c0) | (to565(c1) << 16), aw = (uint32_t) a0 | ((uint32_t) a1 << 8);
	if (COLOR && (ALPHA || DXT != kDxt5))
		ends[t] = make_uint2(cw, aw);
	else if (COLOR)
		ends[t].x = cw;
	else
		ends[t].y = aw;
}

template <int DXT, int CD>
static void launch_search16_cd(dim3 grid, dim3 block, const ImageView &v, uint2 *d_ends, cudaStream_t stream)
{
#ifdef S2TC_SEARCH16_ONE_LAUNCH // A/B: both searches in one kernel
	search16_kernel<DXT, CD, true, true><<<grid, block, 0, stream>>>(v, 1u, d_ends);
#else
	search16_kernel<DXT, CD, true, false><<<grid, block, 0, stream>>>(v, 1u, d_ends);
	if (DXT == kDxt5)
		search16_kernel<DXT, kRGB, false, true><<<grid, block, 0, stream>>>(v, 1u, d_ends);
#endif
}

template <int DXT>
static cudaError_t launch_search16_dxt(int cd, const ImageView &v, uint2 *d_ends, cudaStream_t stream)
{
	const int nblocks = (int) view_blocks(v);
	if (nblocks == 0)
		return cudaSuccess;
	const dim3 block(kSearch16Threads), grid((nblocks + kSearch16Threads - 1) / kSearch16Threads);
	switch (cd) {
#ifdef S2TC_SEARCH16_ONLY_CD // A/B builds of a single metric (see the Makefile's EXTRA)
	case S2TC_SEARCH16_ONLY_CD: launch_search16_cd<DXT, S2TC_SEARCH16_ONLY_CD>(grid, block, v, d_ends, stream); break;
#else
	case kRGB: launch_search16_cd<DXT, kRGB>(grid, block, v, d_ends, stream); break;
	case kYUV: launch_search16_cd<DXT, kYUV>(grid, block, v, d_ends, stream); break;
	case kSRGB: launch_search16_cd<DXT, kSRGB>(grid, block, v, d_ends, stream); break;
	case kSRGB_MIXED: launch_search16_cd<DXT, kSRGB_MIXED>(grid, block, v, d_ends, stream); break;
	case kAVG: launch_search16_cd<DXT, kAVG>(grid, block, v, d_ends, stream); break;
	case kWAVG: launch_search16_cd<DXT, kWAVG>(grid, block, v, d_ends, stream); break;
	case kW0AVG: launch_search16_cd<DXT, kW0AVG>(grid, block, v, d_ends, stream); break;
	case kNORMALMAP: launch_search16_cd<DXT, kNORMALMAP>(grid, block, v, d_ends, stream); break;
#endif
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

// one translation unit per DXT mode (the fully unrolled scans are slow to compile)
cudaError_t S2TC_SEARCH16_NAME(int cd, const ImageView &v, uint2 *d_ends, cudaStream_t stream)
{
	return launch_search16_dxt<S2TC_SEARCH16_DXT>(cd, v, d_ends, stream);
}

S2TC_DEFINE_LUT_INIT(S2TC_SEARCH16_LUT_INIT)

} // namespace s2tc
