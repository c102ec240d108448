// search16.inl -- MODE_NORMAL without random candidates (S2TC_RANDOM_COLORS = 0, or NORMALMAP with
// the default -1): gather, distance matrix, c0/c1 pair search, DXT5 alpha search, refinement and packing
// fused in ONE kernel, one thread per 4x4 block, everything in registers.
//
// Reference path per block: s2tc_algorithm.cpp:938-959 (gather), :997-1001 (single-colour hack), :367-414
// (reduce_colors_inplace), :416-478 (reduce_colors_inplace_2fixpoints), then :1010-1107.
//
// Why this shape: with at most 16 candidates the search is 120 pairs x 16 texels.  A thread keeps the
// 120 distinct distances of the symmetric matrix in registers (all loops fully unrolled, every index a
// compile-time constant), so the scan is pure VIMNMX/IADD3 with no shared memory, no shuffles and no
// synchronisation; 32 independent blocks per warp give the ILP.  Blocks with fewer than 16 colours
// (DXT1 transparency, ragged edges) run the same code with the missing rows/columns zeroed and the
// missing pairs masked, which is exactly what the reference's smaller loops compute.
// The first version of this path (a 4-lane group per block with the matrix in shared memory, kept in
// kernels_search.cu for nrandom > 0) issued 2.6 bank conflicts per LDS and sat at 41 % issue
// utilisation; see profiles/r01a_pair_search_g4_config2.ncu.txt.
#include <utility>

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

#ifndef S2TC_ENCODE16_MINBLOCKS
#define S2TC_ENCODE16_MINBLOCKS 3
#endif
template <int DXT, int CD>
__global__ void __launch_bounds__(128, S2TC_ENCODE16_MINBLOCKS) encode16_kernel(ImageView v, int refine, uint8_t *out)
{
	typedef Metric<CD> M;
	typedef typename M::Feat Feat;
	const int nblocks = v.blocks_w * v.blocks_h;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nblocks)
		return;
	const int by = t / v.blocks_w, bx = t - by * v.blocks_w;
	Block b;
	load_block(v, bx, by, b);

	// ---- gather in the reference's column-major order (ref :940-959) --------------------------------
	uint32_t usemask = 0; // bit o = x*4 + y
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		bool use = (b.valid >> i) & 1u;
		if (DXT == kDxt1)
			use = use && (b.px[i] >> 24) != 0;
		if (use)
			usemask |= 1u << ((i & 3) * 4 + (i >> 2));
	}
	uint32_t c[16];
	uint32_t cl[16]; // the same list, addressable (two dynamic reads after the search)
	int n;
	if (usemask == 0xFFFFu) {
#pragma unroll
		for (int o = 0; o < 16; ++o)
			c[o] = b.px[(o & 3) * 4 + (o >> 2)];
		n = 16;
	} else {
		n = 0;
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
	}
#pragma unroll
	for (int o = 0; o < 16; ++o)
		cl[o] = c[o];

	// ---- colour distance matrix: the 120 entries above the diagonal (ref :375-383) -----------------
	int D[120];
	int fix[16];
#pragma unroll
	for (int k = 0; k < 16; ++k)
		fix[k] = 0;
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
	const uint32_t c0 = px_rgb(cl[bij >> 4]), c1 = px_rgb(cl[bij & 15u]);

	// ---- DXT5: the same search on alpha with the fixed points 0 and 255 (ref :416-478) ----------------
	int a0 = 0, a1 = 0;
	if (DXT == kDxt5) {
		int a[16];
		bool flat = true; // every gathered alpha equal
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			a[k] = (int) (cl[k] >> 24); // re-read: c[] is dead by now, which keeps the scan's live set at the matrix itself
			flat = flat && (k >= n || a[k] == a[0]);
		}
		// Exact shortcut: with a single alpha value every distance is 0, every pair sums to 0, and the reference keeps
		// its first pair (0,1) (ref :467-477).  Measured on config 2 (75 % flat blocks, 32 % flat warps) it made the kernel
		// SLOWER (4.21 ms vs 3.11 ms without it, same box, A/B builds), so it is off unless S2TC_ENCODE16_FLAT_ALPHA is
		// defined; the split into two branches is kept because it removed the DXT5 register spills (3.57 -> 3.11 ms).
#ifndef S2TC_ENCODE16_FLAT_ALPHA
		flat = false;
#endif
		if (__all_sync(__activemask(), flat)) {
			a0 = a[0];
			a1 = a[1];
		} else {
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

	// ---- refinement and packing (ref :1010-1107) ---------------------------------------------------------
	// the texels are read again (L1/L2 hits) instead of being kept live across the two scans: the scans need
	// 120 registers for the matrix alone
	load_block(v, bx, by, b);
	uint32_t w[4];
	finish_block<DXT, CD>(b, refine, c0, c1, a0, a1, w);
	if (DXT == kDxt1)
		reinterpret_cast<uint2 *>(out)[t] = make_uint2(w[0], w[1]);
	else
		reinterpret_cast<uint4 *>(out)[t] = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int DXT>
static cudaError_t launch_encode16_dxt(int cd, int refine, const ImageView &v, void *d_out, cudaStream_t stream)
{
	const int nblocks = v.blocks_w * v.blocks_h;
	if (nblocks == 0)
		return cudaSuccess;
	const dim3 block(128), grid((nblocks + 127) / 128);
	uint8_t *out = (uint8_t *) d_out;
	switch (cd) {
#ifdef S2TC_ENCODE16_ONLY_CD
	case S2TC_ENCODE16_ONLY_CD: encode16_kernel<DXT, S2TC_ENCODE16_ONLY_CD><<<grid, block, 0, stream>>>(v, refine, out); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
#else
	case kRGB: encode16_kernel<DXT, kRGB><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kYUV: encode16_kernel<DXT, kYUV><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kSRGB: encode16_kernel<DXT, kSRGB><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kSRGB_MIXED: encode16_kernel<DXT, kSRGB_MIXED><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kAVG: encode16_kernel<DXT, kAVG><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kWAVG: encode16_kernel<DXT, kWAVG><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kW0AVG: encode16_kernel<DXT, kW0AVG><<<grid, block, 0, stream>>>(v, refine, out); break;
	case kNORMALMAP: encode16_kernel<DXT, kNORMALMAP><<<grid, block, 0, stream>>>(v, refine, out); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
#endif

// one translation unit per DXT mode (the fully unrolled scans are slow to compile)
cudaError_t S2TC_ENCODE16_NAME(int cd, int refine, const ImageView &v, void *d_out, cudaStream_t stream)
{
	return launch_encode16_dxt<S2TC_ENCODE16_DXT>(cd, refine, v, d_out, stream);
}

} // namespace s2tc
