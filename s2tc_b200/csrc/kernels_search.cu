// kernels_search.cu -- MODE_NORMAL step 2: the c0/c1 pair search (reference reduce_colors_inplace and
// reduce_colors_inplace_2fixpoints, s2tc_algorithm.cpp:367-478, reached from :1004-1006), a group of G
// lanes per 4x4 block.
//
// Per block the reference builds dists[m][n] (m = gathered colours n <= 16 plus nrandom random
// candidates) and scans all m(m-1)/2 pairs for the smallest sum_k min(d[i][k], d[j][k]); this is
// >90 % of its run time for nrandom >= 0.  Here:
//   * the G lanes of a group gather the block's colours, append the pre-generated random candidates,
//     cache the per-colour metric features and fill the distance matrix in shared memory,
//     rows padded to 20 words so that 8 lanes reading 8 consecutive rows with LDS.128 hit 8 distinct
//     16-byte bank groups;
//   * row i is held in registers while the lanes stride over j; every lane keeps its first minimum
//     (strict <, in increasing (i,j) order) and the group merges by (sum, i, j) so that the
//     reference's "lexicographically first minimum" survives;
//   * when a sum went negative (only the SRGB metric can wrap) lane 0 replays the
//     reference's exact acceptance rule "bestsum < 0 || sum < bestsum" over the stored matrix;
//   * DXT5 repeats the search for alpha with the two fixed points 0 and 255 folded into one extra row.
// Output: the chosen endpoints, 8 bytes per block; kernels_finish.cu turns them into DXT blocks.
#include "kernels.cuh"

namespace s2tc {

constexpr int kRowWords = 20; // 16 distances + 4 words of padding
constexpr int kSearchThreads = 128;

__host__ __device__ inline size_t search_group_bytes(int mcap, int groups_per_cta)
{
	size_t b = 64 + (size_t) mcap * 4 + (size_t) mcap * 12 + (size_t) (mcap + 1) * kRowWords * 4;
	b = (b + 15) & ~(size_t) 15;
	if (groups_per_cta > 1)
		while ((b & 127) != 64) // neighbouring groups start 16 banks apart
			b += 16;
	return b;
}

template <int G>
__device__ __forceinline__ unsigned group_mask()
{
	if constexpr (G == 32) {
		return 0xFFFFFFFFu;
	} else {
		const unsigned lane = threadIdx.x & 31u;
		return ((1u << G) - 1u) << (lane & ~(unsigned) (G - 1));
	}
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

// sum_k min(a[k], b[k]) (and a fixed row f for alpha), wrapping like the reference's int accumulate
template <bool FIXED>
__device__ __forceinline__ int pair_sum(const int *ri, const int *rowj, const int *rf)
{
	uint32_t s = 0;
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const int4 bj = *reinterpret_cast<const int4 *>(rowj + 4 * q);
		int v0 = min(ri[4 * q + 0], bj.x), v1 = min(ri[4 * q + 1], bj.y);
		int v2 = min(ri[4 * q + 2], bj.z), v3 = min(ri[4 * q + 3], bj.w);
		if (FIXED) {
			v0 = min(v0, rf[4 * q + 0]);
			v1 = min(v1, rf[4 * q + 1]);
			v2 = min(v2, rf[4 * q + 2]);
			v3 = min(v3, rf[4 * q + 3]);
		}
		s += (uint32_t) v0 + (uint32_t) v1 + (uint32_t) v2 + (uint32_t) v3;
	}
	return (int) s;
}

// Scans all pairs i<j<m of the matrix in `dist`; returns (i << 16) | j of the winner in every lane.
template <int G, bool FIXED, bool MAY_BE_NEGATIVE>
__device__ __forceinline__ uint32_t scan_pairs(const int *dist, int m, int lane, unsigned gmask)
{
	int best = 0x7FFFFFFF;
	uint32_t bestij = 1u; // (0,1), the reference's initial besti/bestj
	bool negative = false;
	int rf[16];
	if (FIXED) {
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const int4 f = *reinterpret_cast<const int4 *>(dist + m * kRowWords + 4 * q);
			rf[4 * q] = f.x; rf[4 * q + 1] = f.y; rf[4 * q + 2] = f.z; rf[4 * q + 3] = f.w;
		}
	}
	for (int i = 0; i + 1 < m; ++i) {
		int ri[16];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const int4 a = *reinterpret_cast<const int4 *>(dist + i * kRowWords + 4 * q);
			ri[4 * q] = a.x; ri[4 * q + 1] = a.y; ri[4 * q + 2] = a.z; ri[4 * q + 3] = a.w;
		}
		for (int j = i + 1 + lane; j < m; j += G) {
			const int sum = pair_sum<FIXED>(ri, dist + j * kRowWords, rf);
			if (MAY_BE_NEGATIVE)
				negative |= sum < 0;
			if (sum < best) {
				best = sum;
				bestij = ((uint32_t) i << 16) | (uint32_t) j;
			}
		}
	}
#pragma unroll
	for (int off = G / 2; off > 0; off >>= 1) {
		const int ob = __shfl_xor_sync(gmask, best, off, G);
		const uint32_t oij = __shfl_xor_sync(gmask, bestij, off, G);
		if (ob < best || (ob == best && oij < bestij)) {
			best = ob;
			bestij = oij;
		}
	}
	if (MAY_BE_NEGATIVE) {
		unsigned neg = negative;
#pragma unroll
		for (int off = G / 2; off > 0; off >>= 1)
			neg |= __shfl_xor_sync(gmask, neg, off, G);
		if (neg) { // rare: replay the reference's rule verbatim on one lane (ref :393-410)
			if (lane == 0) {
				int bestsum = -1;
				bestij = 1u;
				for (int i = 0; i < m; ++i)
					for (int j = i + 1; j < m; ++j) {
						uint32_t s = 0;
						for (int k = 0; k < 16; ++k)
							s += (uint32_t) min(dist[i * kRowWords + k], dist[j * kRowWords + k]);
						const int sum = (int) s;
						if (bestsum < 0 || sum < bestsum) {
							bestsum = sum;
							bestij = ((uint32_t) i << 16) | (uint32_t) j;
						}
					}
			}
			bestij = __shfl_sync(gmask, bestij, 0, G);
		}
	}
	return bestij;
}

template <int DXT, int CD, int G>
__global__ void __launch_bounds__(kSearchThreads)
pair_search_kernel(ImageView v, int nrandom, int mcap, size_t group_bytes, const uint16_t *__restrict__ cand_c,
		const uint8_t *__restrict__ cand_a, uint2 *__restrict__ ends)
{
	typedef Metric<CD> M;
	typedef typename M::Feat Feat;
	extern __shared__ __align__(16) uint8_t smem[];
	constexpr int kGroups = kSearchThreads / G;
	const int gi = threadIdx.x / G, lane = threadIdx.x % G;
	const int nblocks = v.blocks_w * v.blocks_h;
	const int t = blockIdx.x * kGroups + gi;
	if (t >= nblocks)
		return;
	const unsigned gmask = group_mask<G>();

	uint8_t *gbase = smem + (size_t) gi * group_bytes;
	uint32_t *px = reinterpret_cast<uint32_t *>(gbase);         // [16]
	int *dist = reinterpret_cast<int *>(gbase + 64);            // [(mcap+1)][kRowWords]
	uint32_t *col = reinterpret_cast<uint32_t *>(gbase + 64 + (size_t) (mcap + 1) * kRowWords * 4); // [mcap]
	Feat *feat = reinterpret_cast<Feat *>(col + mcap);          // [mcap]

	const int by = t / v.blocks_w, bx = t - by * v.blocks_w;
	const int x0 = bx * 4, y0 = by * 4;
	const int w = min(4, v.width - x0), h = min(4, v.rows - y0);

	// 1. texels -> shared
	for (int y = lane; y < 4; y += G) {
		uint32_t r[4];
		load_block_row(v, x0, y0 + y, w, r);
#pragma unroll
		for (int x = 0; x < 4; ++x)
			px[y * 4 + x] = r[x];
	}
	__syncwarp(gmask);

	// 2. gather in the reference's column-major order (ref :940-959); bit o = x*4+y
	const uint32_t valid = valid_mask(w, h);
	uint32_t usemask = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		const uint32_t p = px[i];
		bool use = (valid >> i) & 1u;
		if (DXT == kDxt1)
			use = use && (p >> 24) != 0;
		if (use)
			usemask |= 1u << ((i & 3) * 4 + (i >> 2));
	}
	int n = __popc(usemask);
	for (int o = lane; o < 16; o += G)
		if ((usemask >> o) & 1u)
			col[__popc(usemask & ((1u << o) - 1u))] = px[(o & 3) * 4 + (o >> 2)];
	if (n == 0) {
		if (lane == 0)
			col[0] = 0;
		n = 1;
	}
	int m = n;
	if (nrandom > 0) { // ref :962-993, candidates pre-generated by kernels_misc.cu
		const size_t cb = (size_t) t * nrandom;
		for (int k = lane; k < nrandom; k += G) {
			uint32_t c = from565(cand_c[cb + k]);
			if (DXT == kDxt5)
				c |= (uint32_t) cand_a[cb + k] << 24;
			col[n + k] = c;
		}
		m = n + nrandom;
	}
	__syncwarp(gmask);
	if (nrandom <= 0 && n == 1) { // ref :997-1001
		if (lane == 0)
			col[1] = col[0];
		m = n = 2;
		__syncwarp(gmask);
	}

	// 3. per-colour features
	for (int i = lane; i < m; i += G)
		feat[i] = M::feat(col[i]);
	__syncwarp(gmask);

	// 4. distance matrix, rows zero-padded to 16 (ref :375-392; argument order matters for SRGB)
	for (int e = lane; e < m * 16; e += G) {
		const int i = e >> 4, k = e & 15;
		int d = 0;
		if (k < n && k != i) {
			if (i < n && k < i)
				d = M::dist(feat[k], feat[i]);
			else
				d = M::dist(feat[i], feat[k]);
		}
		dist[i * kRowWords + k] = d;
	}
	__syncwarp(gmask);

	// 5. colour pair scan
	const uint32_t cij = scan_pairs<G, false, M::kMayBeNegative>(dist, m, lane, gmask);
	const uint32_t c0 = col[cij >> 16], c1 = col[cij & 0xFFFFu];
	uint32_t a01 = 0;

	if (DXT == kDxt5) { // ref :416-478
		__syncwarp(gmask);
		for (int e = lane; e < (m + 1) * 16; e += G) {
			const int i = e >> 4, k = e & 15;
			int d = 0;
			if (k < n) {
				const int ak = (int) (col[k] >> 24);
				if (i < m) {
					const int ai = (int) (col[i] >> 24);
					d = (ai - ak) * (ai - ak);
				} else {
					d = min(ak * ak, (255 - ak) * (255 - ak));
				}
			}
			dist[i * kRowWords + k] = d;
		}
		__syncwarp(gmask);
		const uint32_t aij = scan_pairs<G, true, false>(dist, m, lane, gmask);
		a01 = (col[aij >> 16] >> 24) | ((col[aij & 0xFFFFu] >> 24) << 8);
	}
	if (lane == 0)
		ends[t] = make_uint2(to565(c0) | (to565(c1) << 16), a01);
}

template <int DXT, int CD, int G>
static cudaError_t launch_search_g(int nrandom, const ImageView &v, const uint16_t *cand_c, const uint8_t *cand_a,
		uint2 *ends, cudaStream_t stream)
{
	const int nblocks = v.blocks_w * v.blocks_h;
	if (nblocks == 0)
		return cudaSuccess;
	constexpr int kGroups = kSearchThreads / G;
	const int mcap = 16 + (nrandom > 0 ? nrandom : 0);
	const size_t gb = search_group_bytes(mcap, kGroups);
	const size_t smem = gb * kGroups;
	auto kern = pair_search_kernel<DXT, CD, G>;
	if (smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess)
			return e;
	}
	const dim3 block(kSearchThreads), grid((nblocks + kGroups - 1) / kGroups);
	kern<<<grid, block, smem, stream>>>(v, nrandom, mcap, gb, cand_c, cand_a, ends);
	return cudaGetLastError();
}

int pair_search_max_nrandom()
{
	// one 32-lane group... four per CTA; 227 KB of shared memory per CTA
	int lo = 0, hi = 1 << 16;
	while (lo < hi) {
		int mid = (lo + hi + 1) / 2;
		if (search_group_bytes(16 + mid, kSearchThreads / 32) * (kSearchThreads / 32) <= 227 * 1024)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

template <int DXT, int CD>
static cudaError_t launch_search_cd(int nrandom, const ImageView &v, const uint16_t *cand_c, const uint8_t *cand_a,
		uint2 *ends, cudaStream_t stream)
{
	// group size: 4 lanes for the 120-pair search of nrandom == 0, a full warp once random
	// candidates multiply the pair count
	if (nrandom <= 0)
		return launch_search_g<DXT, CD, 4>(nrandom, v, cand_c, cand_a, ends, stream);
	if (nrandom > pair_search_max_nrandom())
		return cudaErrorInvalidValue;
	return launch_search_g<DXT, CD, 32>(nrandom, v, cand_c, cand_a, ends, stream);
}

template <int DXT>
static cudaError_t launch_search_dxt(int cd, int nrandom, const ImageView &v, const uint16_t *cand_c,
		const uint8_t *cand_a, uint2 *ends, cudaStream_t stream)
{
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

} // namespace s2tc
