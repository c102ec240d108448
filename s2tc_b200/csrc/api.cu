// api.cu -- the s2tc_b200_* C ABI (include/s2tc_b200.h): contexts, workspaces, and the orchestration of
// the kernels in kernels_*.cu.  No encoding arithmetic happens on the host in this file: the host
// side only sizes buffers, computes the rand() jump polynomials for a launch, and moves bytes.
#include "../../include/s2tc_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"

using namespace s2tc;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_last_error = buf;
	return code;
}

#define CU(call)                                                                                           \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess)                                                                             \
			return fail(e_ == cudaErrorMemoryAllocation ? S2TC_B200_ENOMEM : S2TC_B200_ECUDA, "%s: %s", #call, \
					cudaGetErrorString(e_));                                                               \
	} while (0)

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n)
	{
		if (n <= cap)
			return cudaSuccess;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 8 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess)
			cap = want;
		return e;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

enum Family { kFamPrepass = 0, kFamCand, kFamSearch, kFamFinish, kFamFast, kFamTranscode, kNumFam };

struct PendingTiming {
	int fam;
	cudaEvent_t a, b;
};

constexpr int kPlanRing = 64;
constexpr long long kSlabBlocks = 1 << 22; // MODE_NORMAL works through an image in slabs of at most this many blocks

} // namespace

namespace {

// Host-side copies between the caller's pageable memory and the pinned staging buffers, on a few threads: one memcpy
// thread moves ~10 GB/s, the PCIe link takes ~50.  (The driver's own pageable path stages on the calling thread:
// tx_compress_dxtn on malloc'd memory took 145 ms for config 3 against 42 ms from pinned memory.)
class CopyPool {
public:
	static CopyPool &get()
	{
		// never destroyed: its threads wait on the condition variable for the life of the process, and destroying a
		// condition variable with waiters blocks (the process would hang in its exit handlers)
		static CopyPool *pool = new CopyPool;
		return *pool;
	}
	void copy(void *dst, const void *src, size_t n)
	{
		const size_t nth = workers_.size() + 1;
		if (n < (1u << 20) || nth == 1) {
			memcpy(dst, src, n);
			return;
		}
		std::unique_lock<std::mutex> call(call_mu_); // one parallel copy at a time
		const size_t part = ((n + nth - 1) / nth + 4095) & ~(size_t) 4095;
		{
			std::lock_guard<std::mutex> lk(mu_);
			dst_ = (uint8_t *) dst;
			src_ = (const uint8_t *) src;
			n_ = n;
			part_ = part;
			pending_ = (int) workers_.size();
			++gen_;
		}
		cv_.notify_all();
		run_part(0, (uint8_t *) dst, (const uint8_t *) src, n, part);
		std::unique_lock<std::mutex> lk(mu_);
		done_cv_.wait(lk, [&] { return pending_ == 0; });
	}

private:
	CopyPool()
	{
		const char *e = getenv("S2TC_B200_COPY_THREADS");
		int n = e ? atoi(e) : 0;
		if (n <= 0) {
			const unsigned hw = std::thread::hardware_concurrency();
			n = hw >= 16 ? 8 : (hw >= 4 ? (int) hw / 2 : 1);
		}
		for (int i = 1; i < n; ++i)
			workers_.emplace_back([this, i] { loop(i); });
		for (auto &t : workers_)
			t.detach(); // live for the life of the process (the library has no teardown call, like the reference's)
	}
	static void run_part(size_t idx, uint8_t *dst, const uint8_t *src, size_t n, size_t part)
	{
		const size_t a = idx * part;
		if (a < n)
			memcpy(dst + a, src + a, n - a < part ? n - a : part);
	}
	void loop(int idx)
	{
		uint64_t seen = 0;
		for (;;) {
			uint8_t *dst;
			const uint8_t *src;
			size_t n, part;
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [&] { return gen_ != seen; });
				seen = gen_;
				dst = dst_;
				src = src_;
				n = n_;
				part = part_;
			}
			run_part((size_t) idx, dst, src, n, part);
			{
				std::lock_guard<std::mutex> lk(mu_);
				--pending_;
			}
			done_cv_.notify_one();
		}
	}
	std::vector<std::thread> workers_;
	std::mutex mu_, call_mu_;
	std::condition_variable cv_, done_cv_;
	uint8_t *dst_ = nullptr;
	const uint8_t *src_ = nullptr;
	size_t n_ = 0, part_ = 0;
	int pending_ = 0;
	uint64_t gen_ = 0;
};

// page-locked (cudaHostAlloc / cudaHostRegister) or managed: the DMA engines can read / write it directly
bool host_pinned(const void *p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

constexpr size_t kStageChunk = (size_t) 8 << 20; // staging granularity of pageable uploads
constexpr int kStageSlots = 4;

} // namespace

struct s2tc_b200_ctx {
	int device = 0;
	cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
	std::mutex mu;
	DevBuf src, reduced, out, ends, dither_ws, shard_ws, plans, small, mip, rand_ws;
	// second lane (see encode_rows): with random candidates consecutive slabs run on two streams so that the tail of one
	// search launch (its last one-warp CTAs, ~0.3 ms each, on a mostly idle GPU) overlaps the next slab's kernels
	DevBuf reduced1, ends1, dither_ws1, rand_ws1, carries;
	cudaStream_t aux = nullptr, aux2 = nullptr;
	// pageable caller memory is staged through these (allocated on first use)
	uint8_t *h_stage_in = nullptr, *h_stage_out = nullptr;
	size_t h_stage_out_cap = 0;
	cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_prepass = nullptr;
	RandPlan *h_plans = nullptr; // pinned ring
	int plan_next = 0;
	int *h_carry = nullptr; // pinned, 4 ints
	uint64_t *h_summary = nullptr; // pinned, 16 words
	uint8_t *h_block = nullptr;   // pinned, 64 + 16 bytes for the single-block path
	// the dither workspace holds the chunk/tile maps of this texel range, left by a summary call on maps_stream; only
	// s2tc_b200_encode_rows_after_summary_async may reuse them, and every other call forgets them
	const void *maps_src = nullptr;
	size_t maps_npix = 0;
	int maps_comps = 0, maps_abits = 0;
	cudaStream_t maps_stream = nullptr;
	// workspaces are per context: a call on another stream than the previous one first waits for that call's work
	cudaEvent_t last_done = nullptr;
	cudaStream_t last_stream = nullptr;
	bool last_valid = false;
	uint64_t launches = 0;
	bool profiling = false;
	double fam_ms[kNumFam] = {0};
	uint64_t fam_launches[kNumFam] = {0};
	std::vector<PendingTiming> pending;
};

namespace {

struct FamScope { // brackets a group of launches of one family with events when profiling is on
	s2tc_b200_ctx *c;
	cudaStream_t st;
	int fam;
	cudaEvent_t a = nullptr, b = nullptr;
	FamScope(s2tc_b200_ctx *c_, cudaStream_t st_, int fam_, int nlaunch) : c(c_), st(st_), fam(fam_)
	{
		c->launches += nlaunch;
		c->fam_launches[fam] += nlaunch;
		if (c->profiling) {
			cudaEventCreate(&a);
			cudaEventCreate(&b);
			cudaEventRecord(a, st);
		}
	}
	~FamScope()
	{
		if (a) {
			cudaEventRecord(b, st);
			c->pending.push_back({fam, a, b});
		}
	}
};

// Every entry point that enqueues work brackets it with this: the context's scratch buffers (reduced texels, endpoints,
// dither maps, rand() windows, the small slots) are shared by all calls, so work enqueued on a different stream than the
// previous call's is ordered behind it by an event.  Calls on one stream cost one event record.
struct StreamOrder {
	s2tc_b200_ctx *c;
	cudaStream_t st;
	StreamOrder(s2tc_b200_ctx *c_, cudaStream_t st_, bool keep_maps = false) : c(c_), st(st_)
	{
		if (c->last_valid && c->last_stream != st)
			cudaStreamWaitEvent(st, c->last_done, 0);
		if (!keep_maps)
			c->maps_src = nullptr;
	}
	~StreamOrder()
	{
		if (cudaEventRecord(c->last_done, st) == cudaSuccess) {
			c->last_stream = st;
			c->last_valid = true;
		}
	}
};

int settings_normalise(const s2tc_b200_settings *in, s2tc_b200_settings &s)
{
	if (!in)
		return fail(S2TC_B200_EINVAL, "settings is NULL");
	s = *in;
	s.dxt = norm_dxt(s.dxt);
	s.cd = norm_cd(s.cd);
	s.refine = norm_refine(s.refine);
	if (s.dither != kDitherNone && s.dither != kDitherFloyd)
		s.dither = kDitherSimple; // ref s2tc_algorithm.cpp:1419-1431
	return 0;
}

// Encodes the blocks of `v` (a slab whose first block is block number blk0 of the image) into d_dst.
int encode_view(s2tc_b200_ctx *c, const s2tc_b200_settings &s, const ImageView &v, long long blk0, uint64_t cursor0,
		void *d_dst, cudaStream_t st, int lane = 0)
{
	const long long nblocks = view_blocks(v);
	if (nblocks == 0)
		return 0;
	DevBuf &ends = lane ? c->ends1 : c->ends, &rand_ws = lane ? c->rand_ws1 : c->rand_ws;
	if (v.images > 1 && s.nrandom > 0) { // the chunked rand() stream of the search kernel is per image: one image at a time
		ImageView one = v;
		one.images = 1;
		for (int i = 0; i < v.images; ++i) {
			one.base = v.base + (size_t) i * v.image_bytes;
			if (int rc = encode_view(c, s, one, blk0, cursor0, (uint8_t *) d_dst + (size_t) i * v.out_image_bytes, st, lane))
				return rc;
		}
		return 0;
	}
	if (is_fast_mode(s.cd, s.nrandom)) {
		FamScope f(c, st, kFamFast, 1);
		CU(launch_fast_encode(s.dxt, s.cd, s.refine, v, d_dst, st));
		return 0;
	}
	const int nrandom = s.nrandom > 0 ? s.nrandom : 0;
	CU(ends.reserve((size_t) nblocks * sizeof(uint2)));
	if (nrandom == 0) { // <= 16 candidates: the register-resident search, then refinement + packing
		{
			FamScope f(c, st, kFamSearch, s.dxt == kDxt5 ? 2 : 1); // DXT5: colour launch + alpha launch
			CU(launch_search16(s.dxt, s.cd, v, (uint2 *) ends.p, st));
		}
		FamScope f(c, st, kFamFinish, 1);
		CU(launch_finish(s.dxt, s.cd, s.refine, v, (const uint2 *) ends.p, d_dst, st));
		return 0;
	}
	if (nrandom > pair_search_max_nrandom())
		return fail(S2TC_B200_EUNSUPPORTED, "S2TC_RANDOM_COLORS=%d exceeds the %d candidates the search kernel can hold in shared memory",
				nrandom, pair_search_max_nrandom());
	// jump polynomials for this slab: the warp that owns chunk t (kSearchChunkBlocks blocks) starts at
	// cursor0 + (blk0 + kSearchChunkBlocks t) * draws_per_block
	const uint64_t dpb = (uint64_t) draws_per_block(s.dxt, nrandom);
	if (c->plan_next == kPlanRing) { // ring exhausted: wait for the uploads queued so far (on either lane)
		CU(cudaStreamSynchronize(st));
		CU(cudaStreamSynchronize(c->aux));
		CU(cudaStreamSynchronize(c->aux2));
		CU(cudaStreamSynchronize(c->stream));
		c->plan_next = 0;
	}
	RandPlan *hp = &c->h_plans[c->plan_next];
	RandPlan *dp = (RandPlan *) c->plans.p + c->plan_next;
	c->plan_next++;
	rand_plan_init(*hp, cursor0 + (uint64_t) blk0 * dpb, (uint64_t) kSearchChunkBlocks * dpb);
	const RandPlan *hp_dev = nullptr;
	CU(cudaHostGetDevicePointer((void **) &hp_dev, hp, 0));
	CU(rand_ws.reserve(rand_windows_bytes((size_t) nblocks)));
	{
		FamScope f(c, st, kFamCand, 2);
		CU(launch_plan_upload(hp_dev, dp, st));
		CU(launch_rand_windows(dp, (unsigned) ((nblocks + kSearchChunkBlocks - 1) / kSearchChunkBlocks), (uint32_t *) rand_ws.p, st));
	}
	{
		FamScope f(c, st, kFamSearch, 1);
		CU(launch_pair_search(s.dxt, s.cd, nrandom, v, (const uint32_t *) rand_ws.p, (uint2 *) ends.p, st));
	}
	{
		FamScope f(c, st, kFamFinish, 1);
		CU(launch_finish(s.dxt, s.cd, s.refine, v, (const uint2 *) ends.p, d_dst, st));
	}
	return 0;
}

// d_src_rows: texel row 4*row0 of the image.  d_carry: device ints or NULL.
// ready_ws: a dither workspace that already holds the chunk/tile maps of exactly these texels (phase 1 is skipped), or NULL.
int encode_rows(s2tc_b200_ctx *c, const s2tc_b200_settings &s, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t cursor0, int *d_carry, cudaStream_t st,
		void *ready_ws = nullptr, bool src_is_reduced = false, int lane = -1)
{
	// lane -1: everything on st with the first set of workspaces, except that with random candidates the slabs of the range
	// alternate between st and the context's second stream (and workspace set) so that one search launch's tail overlaps
	// the next slab; st waits for the second stream at the end.  lane 0 / 1: the caller pipelines ranges itself
	// (compress_rows_host, the striped shards): this range uses workspace set `lane` on st, and ev_prepass is recorded on
	// st when the 565 pre-pass (the only step that reads and writes the DITHER_SIMPLE carry) has been enqueued.
	DevBuf &reduced = lane == 1 ? c->reduced1 : c->reduced, &dither_ws = lane == 1 ? c->dither_ws1 : c->dither_ws;
	const int bh = (height + 3) / 4, bw = (width + 3) / 4;
	if (width <= 0 || height <= 0 || row0 < 0 || row1 > bh || row0 > row1)
		return fail(S2TC_B200_EINVAL, "bad geometry %dx%d rows [%d,%d)", width, height, row0, row1);
	if (row0 == row1)
		return 0;
	const int comps = srccomps == 3 ? 3 : 4;
	const int abits = alpha_bits(s.dxt);
	const int y0 = row0 * 4, y1 = row1 * 4 < height ? row1 * 4 : height;
	const int rows = y1 - y0;
	const size_t npix = (size_t) width * rows;

	const uint8_t *texels = (const uint8_t *) d_src_rows;
	int fmt = comps == 3 ? kSrcRGB8 : kSrcRGBA8;
	size_t texel_bytes = comps;
	if (src_is_reduced) { // the caller ran the 565 pre-pass itself (s2tc_b200_floyd_rows_device): 4 bytes per texel {r5, g6, b5, a}
		fmt = kSrcReduced;
		texel_bytes = 4;
	} else if (s.dither == kDitherSimple) {
		CU(reduced.reserve(npix * 4));
		CU(dither_ws.reserve(dither_workspace_bytes(npix)));
		int *carry = d_carry;
		if (!carry) {
			carry = (int *) c->small.p;
			CU(cudaMemsetAsync(carry, 0, 4 * sizeof(int), st));
		}
		const bool ready = ready_ws != nullptr;
		FamScope f(c, st, kFamPrepass, prepass_simple_launches(npix, ready));
		CU(launch_prepass_simple(d_src_rows, comps, abits, npix, reduced.p, carry, ready ? ready_ws : dither_ws.p, ready, st));
		texels = (const uint8_t *) reduced.p;
		fmt = kSrcReduced;
		texel_bytes = 4;
	} else if (s.dither == kDitherFloyd) {
		if (row0 != 0 || row1 != bh)
			return fail(S2TC_B200_EUNSUPPORTED, "DITHER_FLOYDSTEINBERG diffuses error between rows: encode the whole image in one call "
					"(block rows [%d,%d) of %d requested)", row0, row1, bh);
		CU(reduced.reserve(npix * 4));
		CU(dither_ws.reserve(floyd_workspace_bytes(width, height)));
		FamScope f(c, st, kFamPrepass, comps == 4 && abits != 8 ? 2 : 1);
		CU(launch_prepass_floyd(d_src_rows, comps, abits, width, height, reduced.p, dither_ws.p, st));
		texels = (const uint8_t *) reduced.p;
		fmt = kSrcReduced;
		texel_bytes = 4;
	}
	if (lane >= 0)
		CU(cudaEventRecord(c->ev_prepass, st));

	// MODE_NORMAL goes slab by slab to bound the candidate/endpoint workspaces
	// only the random-candidate path has per-block workspaces to bound; the fused kernels take the range in one launch
	const bool one_launch = s.nrandom <= 0;
	long long rows_per_slab = one_launch ? (row1 - row0) : (kSlabBlocks / bw > 0 ? kSlabBlocks / bw : 1);
	const int bs = block_bytes(s.dxt);
	// two lanes inside this range?  (not while per-family times are being taken: they assume one stream)
	const bool alternate = lane < 0 && !one_launch && (row1 - row0) >= rows_per_slab + rows_per_slab / 2 && !c->profiling && st != c->aux;
	if (alternate) {
		CU(cudaEventRecord(c->ev_fork, st));
		CU(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
	}
	int k = 0;
	for (long long r = row0, rnext; r < row1; r = rnext, ++k) {
		int r1 = (int) (r + rows_per_slab < row1 ? r + rows_per_slab : row1);
		if (row1 - r1 < rows_per_slab / 2)
			r1 = row1; // a small remainder joins this slab: every launch of the search kernel costs a tail (compress_rows_host)
		const int sy0 = (int) r * 4 - y0, sy1 = (r1 * 4 < height ? r1 * 4 : height) - y0;
		const ImageView v = make_view(texels + (size_t) sy0 * width * texel_bytes, width, sy1 - sy0, fmt, abits);
		const int l = alternate ? (k & 1) : (lane == 1 ? 1 : 0);
		int rc = encode_view(c, s, v, (long long) r * bw, cursor0, (uint8_t *) d_dst + (size_t) (r - row0) * bw * bs,
				alternate && l ? c->aux : st, l);
		if (rc)
			return rc;
		rnext = r1;
	}
	if (alternate) {
		CU(cudaEventRecord(c->ev_join, c->aux));
		CU(cudaStreamWaitEvent(st, c->ev_join, 0));
	}
	return 0;
}

} // namespace

extern "C" {

const char *s2tc_b200_last_error(void) { return g_last_error.c_str(); }

int s2tc_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int s2tc_b200_ctx_create(int device, s2tc_b200_ctx **out)
{
	if (!out)
		return fail(S2TC_B200_EINVAL, "out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		return fail(S2TC_B200_ENODEVICE, "no CUDA device available (%s); this library has no CPU path",
				e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	}
	if (device < 0 || device >= n)
		return fail(S2TC_B200_EINVAL, "device %d out of range (have %d)", device, n);
	CU(cudaSetDevice(device));
	s2tc_b200_ctx *c = new s2tc_b200_ctx();
	c->device = device;
	const int rc = [&]() -> int {
		CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming));
		CU(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&c->aux2, cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&c->ev_prepass, cudaEventDisableTiming));
		CU(c->carries.reserve(4 * sizeof(int) * 256));
		CU(cudaHostAlloc((void **) &c->h_plans, sizeof(RandPlan) * kPlanRing, cudaHostAllocMapped));
		CU(cudaHostAlloc((void **) &c->h_carry, 4 * sizeof(int), cudaHostAllocDefault));
		CU(cudaHostAlloc((void **) &c->h_summary, 16 * sizeof(uint64_t), cudaHostAllocDefault));
		CU(cudaHostAlloc((void **) &c->h_block, 128, cudaHostAllocDefault));
		CU(c->plans.reserve(sizeof(RandPlan) * kPlanRing));
		CU(c->small.reserve(1024));
		CU(init_all_luts(c->stream)); // device-side tables of the metrics (one copy per translation unit), on this device
		CU(cudaStreamSynchronize(c->stream));
		return 0;
	}();
	if (rc) { // keep the message of the failing call; release whatever was created
		const std::string msg = g_last_error;
		s2tc_b200_ctx_destroy(c);
		g_last_error = msg;
		return rc;
	}
	*out = c;
	return 0;
}

void s2tc_b200_ctx_destroy(s2tc_b200_ctx *c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	for (auto &p : c->pending) {
		cudaEventDestroy(p.a);
		cudaEventDestroy(p.b);
	}
	DevBuf *bufs[] = {&c->src, &c->reduced, &c->out, &c->ends, &c->dither_ws, &c->shard_ws, &c->plans, &c->small, &c->mip, &c->rand_ws,
			&c->reduced1, &c->ends1, &c->dither_ws1, &c->rand_ws1, &c->carries};
	for (DevBuf *b : bufs)
		b->release();
	cudaFreeHost(c->h_plans);
	cudaFreeHost(c->h_carry);
	cudaFreeHost(c->h_summary);
	cudaFreeHost(c->h_block);
	if (c->h_stage_in)
		cudaFreeHost(c->h_stage_in);
	if (c->h_stage_out)
		cudaFreeHost(c->h_stage_out);
	for (cudaEvent_t e : c->stage_ev)
		if (e)
			cudaEventDestroy(e);
	if (c->last_done)
		cudaEventDestroy(c->last_done);
	for (cudaEvent_t e : {c->ev_fork, c->ev_join, c->ev_prepass})
		if (e)
			cudaEventDestroy(e);
	if (c->aux)
		cudaStreamDestroy(c->aux);
	if (c->aux2)
		cudaStreamDestroy(c->aux2);
	if (c->stream)
		cudaStreamDestroy(c->stream);
	if (c->copy_in)
		cudaStreamDestroy(c->copy_in);
	if (c->copy_out)
		cudaStreamDestroy(c->copy_out);
	cudaGetLastError();
	delete c;
}

s2tc_b200_ctx *s2tc_b200_default_ctx(void)
{
	static std::mutex mu;
	static s2tc_b200_ctx *ctx = nullptr;
	std::lock_guard<std::mutex> lock(mu);
	if (!ctx) {
		const char *d = getenv("S2TC_B200_DEVICE");
		if (s2tc_b200_ctx_create(d ? atoi(d) : 0, &ctx) != 0)
			ctx = nullptr;
	}
	return ctx;
}

int s2tc_b200_encode_rows_device(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, int *carry, void *stream)
{
	if (!c || !d_src_rows || !d_dst)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	int *d_carry = nullptr;
	if (carry && s.dither == kDitherSimple) {
		d_carry = (int *) c->small.p + 8;
		memcpy(c->h_carry, carry, 4 * sizeof(int));
		CU(cudaMemcpyAsync(d_carry, c->h_carry, 4 * sizeof(int), cudaMemcpyHostToDevice, st));
	}
	if (int rc = encode_rows(c, s, srccomps, width, height, d_src_rows, row0, row1, d_dst, rand_cursor0, d_carry, st))
		return rc;
	if (d_carry) {
		CU(cudaMemcpyAsync(c->h_carry, d_carry, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		memcpy(carry, c->h_carry, 4 * sizeof(int));
	}
	return 0;
}

// Fully asynchronous variant for sharded encodes: the summary stays on the device (d_maps, 128 bytes) so that it can
// be all-gathered and folded there without a host round trip.
int s2tc_b200_dither_summary_async(s2tc_b200_ctx *c, int srccomps, int alphabits, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_maps, void *stream)
{
	if (!c || !d_src_rows || !d_maps)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	const int bh = (height + 3) / 4;
	if (width <= 0 || height <= 0 || row0 < 0 || row1 > bh || row0 >= row1)
		return fail(S2TC_B200_EINVAL, "bad geometry");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	const int comps = srccomps == 3 ? 3 : 4;
	const int y0 = row0 * 4, y1 = row1 * 4 < height ? row1 * 4 : height;
	const size_t npix = (size_t) width * (y1 - y0);
	StreamOrder order(c, st);
	CU(c->dither_ws.reserve(dither_workspace_bytes(npix)));
	FamScope f(c, st, kFamPrepass, kDitherSummaryLaunches);
	CU(launch_dither_summary(d_src_rows, comps, alphabits, npix, (ByteMap *) d_maps, c->dither_ws.p, st));
	c->maps_src = d_src_rows; // for s2tc_b200_encode_rows_after_summary_async
	c->maps_npix = npix;
	c->maps_comps = comps;
	c->maps_abits = alphabits;
	c->maps_stream = st;
	return 0;
}

// Folds the summaries of shards 0 .. rank-1 (d_all_maps: `rank` or more consecutive 128-byte summaries on the device)
// into the carry entering shard `rank`; d_carry: 4 ints on the device.  Asynchronous.
int s2tc_b200_fold_carry_async(s2tc_b200_ctx *c, const void *d_all_maps, int rank, int srccomps, int alphabits, int *d_carry,
		void *stream)
{
	if (!c || !d_all_maps || !d_carry || rank < 0)
		return fail(S2TC_B200_EINVAL, "bad argument");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st, /*keep_maps=*/true);
	FamScope f(c, st, kFamPrepass, 1);
	CU(launch_fold_carry((const ByteMap *) d_all_maps, rank, srccomps == 3 ? 3 : 4, alphabits, d_carry, st));
	return 0;
}

// encode_rows with the DITHER_SIMPLE carry on the device (in/out), no host synchronisation
int s2tc_b200_encode_rows_async(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, int *d_carry, void *stream)
{
	if (!c || !d_src_rows || !d_dst)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	return encode_rows(c, s, srccomps, width, height, d_src_rows, row0, row1, d_dst, rand_cursor0, d_carry, st);
}

// The same, for the caller that has just summarised exactly these texels with s2tc_b200_dither_summary_async on the same
// context and stream and has NOT changed them since: the chunk/tile maps the summary left in the workspace are reused
// (the first of the three DITHER_SIMPLE phases is skipped).  If the record of the last summary does not match this
// range, or anything but s2tc_b200_fold_carry_async ran on the context in between, the maps are simply recomputed.
int s2tc_b200_encode_rows_after_summary_async(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, int *d_carry, void *stream)
{
	if (!c || !d_src_rows || !d_dst)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	const int comps = srccomps == 3 ? 3 : 4;
	const int bh = (height + 3) / 4;
	bool ready = false;
	if (s.dither == kDitherSimple && width > 0 && height > 0 && row0 >= 0 && row1 <= bh && row0 < row1) {
		const int y0 = row0 * 4, y1 = row1 * 4 < height ? row1 * 4 : height;
		ready = c->maps_src == d_src_rows && c->maps_npix == (size_t) width * (y1 - y0) && c->maps_comps == comps &&
				c->maps_abits == alpha_bits(s.dxt) && c->maps_stream == st;
	}
	StreamOrder order(c, st);
	return encode_rows(c, s, srccomps, width, height, d_src_rows, row0, row1, d_dst, rand_cursor0, d_carry, st,
			ready ? c->dither_ws.p : nullptr);
}

int s2tc_b200_dither_summary_device(s2tc_b200_ctx *c, int srccomps, int alphabits, int width, int height,
		const void *d_src_rows, int row0, int row1, uint64_t maps[16], void *stream)
{
	if (!c || !d_src_rows || !maps)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	const int bh = (height + 3) / 4;
	if (width <= 0 || height <= 0 || row0 < 0 || row1 > bh || row0 > row1)
		return fail(S2TC_B200_EINVAL, "bad geometry");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	const int comps = srccomps == 3 ? 3 : 4;
	const int y0 = row0 * 4, y1 = row1 * 4 < height ? row1 * 4 : height;
	const size_t npix = (size_t) width * (y1 - y0);
	ByteMap *d_sum = (ByteMap *) ((uint8_t *) c->small.p + 256);
	if (npix == 0) {
		const int kinds[4] = {kChanShift3, kChanShift2, kChanShift3, alpha_chan_kind(comps, alphabits)};
		for (int ch = 0; ch < 4; ++ch) {
			ByteMap m;
			bmap_identity(m, kinds[ch]);
			memcpy(maps + 4 * ch, m.e, sizeof(m.e));
		}
		return 0;
	}
	StreamOrder order(c, st);
	CU(c->dither_ws.reserve(dither_workspace_bytes(npix)));
	{
		FamScope f(c, st, kFamPrepass, kDitherSummaryLaunches);
		CU(launch_dither_summary(d_src_rows, comps, alphabits, npix, d_sum, c->dither_ws.p, st));
	}
	CU(cudaMemcpyAsync(c->h_summary, d_sum, 4 * sizeof(ByteMap), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	memcpy(maps, c->h_summary, 16 * sizeof(uint64_t));
	return 0;
}

int s2tc_b200_carry_apply(const uint64_t map[4], int channel, int srccomps, int alphabits, int carry_in)
{
	ByteMap m;
	memcpy(m.e, map, sizeof(m.e));
	const int kind = channel == 1 ? kChanShift2 : (channel == 3 ? alpha_chan_kind(srccomps == 3 ? 3 : 4, alphabits) : kChanShift3);
	return bmap_apply(m, kind, carry_in);
}

} // extern "C"

namespace {

// How a shard of a DITHER_SIMPLE image learns the carry entering it (s2tc_b200_compress_host_shard).
struct CarryExchange {
	int rank, nslab;          // this shard's position; summaries every shard contributes
	ByteMap *d_mine, *d_all;  // nslab summaries of this shard; world * nslab summaries after `gather`
	void (*gather)(void *);   // enqueues the all-gather of d_mine into d_all on the compute stream
	void *user;
};

// Block rows [row0, row1) of a width x height image from host memory to host memory.  src_rows: texel row 4 * row0.
// dest: first block of row0; row_bytes apart per block row.  st: the compute stream.
//
// Large ranges go through in slabs of block rows on three streams: while slab s is encoded, slab s+1 is on its way up
// and slab s-1 on its way down (PCIe is full duplex), so the call costs about max(copy, kernels) instead of their sum.
// The two pieces of cross-slab state stay on the device: the DITHER_SIMPLE carry (chained through d_carry on the
// compute stream) and the rand cursor (closed form per block row).
// With a CarryExchange (one shard of several, DITHER_SIMPLE) the work has two phases: every slab is summarised as it
// lands (its chunk/tile maps stay in its own piece of workspace), the shards all-gather the summaries, each folds those
// before it into its carry, and then the slabs are replayed from their maps, encoded and downloaded.
int compress_rows_host(s2tc_b200_ctx *c, const s2tc_b200_settings &s, int comps, int width, int height, const uint8_t *src_rows,
		int row0, int row1, uint8_t *dest, size_t row_bytes, uint64_t cursor, const CarryExchange *ex, cudaStream_t st)
{
	const int bw = (width + 3) / 4, bh = (height + 3) / 4, bs = block_bytes(s.dxt);
	const int nrows = row1 - row0;
	const int ty0 = row0 * 4, ty1 = row1 * 4 < height ? row1 * 4 : height;
	const size_t in_bytes = (size_t) width * (ty1 - ty0) * comps;
	const size_t tight = (size_t) bw * bs, out_bytes = tight * nrows;
	const int abits = alpha_bits(s.dxt);
	const bool exchange = ex && s.dither == kDitherSimple;
	if (s.dither == kDitherFloyd && (row0 != 0 || row1 != bh))
		return fail(S2TC_B200_EUNSUPPORTED, "DITHER_FLOYDSTEINBERG diffuses error between rows: encode the whole image in one call "
				"(block rows [%d,%d) of %d requested)", row0, row1, bh);

	CU(c->src.reserve(in_bytes));
	CU(c->out.reserve(out_bytes));
	static const int slab_mb = [] { const char *e = getenv("S2TC_B200_SLAB_MB"); int n = e ? atoi(e) : 0; return n > 0 ? n : 16; }();
	// ~16 MiB of texels per slab by default (measured on config 2: 32 / 16 / 8 MiB -> e2e 5.9 / 5.5 / 7.2 ms); with random
	// candidates the kernels dominate the copies and every slab costs a jump-ahead plan on the host: 4x larger slabs
	int nslab = (int) (in_bytes / ((size_t) slab_mb << (s.nrandom > 0 ? 22 : 20)));
	nslab = nslab < 1 ? 1 : (nslab > 64 ? 64 : nslab);
	if (exchange)
		nslab = ex->nslab; // every shard contributes the same number of summaries
	if (nslab > nrows)
		nslab = nrows;
	if (row_bytes < tight || s.dither == kDitherFloyd)
		nslab = 1; // overlapping destination rows: single ordered copy at the end; Floyd-Steinberg: one 2-D recurrence
	// Random candidates: every launch of the search kernel ends with a tail in which the last chunks (one warp, ~0.3 ms
	// each) run on a mostly idle GPU -- measured 0.26 ms per slab whatever its size (tools_lab/e2e_c3.py: 8 / 16 / 32 / 64
	// slabs).  So the slabs GROW: the first one is small (the kernels start after 16 MiB are up, not 64), every next one
	// twice as large up to 256 MiB; the upload (~52 GB/s) stays ahead of the kernels (~25 GB/s of texels) from the second
	// slab on.  first_rows > 0 selects this schedule.
	int first_rows = 0, cap_rows = 0;
	if (s.nrandom > 0 && !exchange && nslab > 1 && !getenv("S2TC_B200_SLAB_MB")) {
		const size_t row_in = (size_t) width * 4 * comps; // texel bytes per block row
		static const int cap_mb = [] { const char *e = getenv("S2TC_B200_SLAB_CAP_MB"); int n = e ? atoi(e) : 0; return n >= 16 ? n : 256; }();
		first_rows = (int) (((size_t) 16 << 20) / row_in);
		cap_rows = (int) (((size_t) cap_mb << 20) / row_in);
		if (first_rows < 1 || cap_rows < first_rows) {
			first_rows = 0;
		} else {
			int n = 0, left = nrows, sz = first_rows;
			while (left > 0) {
				left -= sz;
				if (left > 0 && left * 2 < sz) // a small remainder joins the last slab instead of costing a launch of its own
					left = 0;
				sz = sz * 2 < cap_rows ? sz * 2 : cap_rows;
				++n;
			}
			nslab = n;
		}
	}
	int *d_carry = nullptr;
	if (s.dither == kDitherSimple) {
		d_carry = (int *) c->small.p + 8;
		if (!exchange)
			CU(cudaMemsetAsync(d_carry, 0, 4 * sizeof(int), st));
	}
	// S2TC_B200_TRACE=1: print when every slab's upload, kernels and download finished (ms since the call started)
	static const bool trace = getenv("S2TC_B200_TRACE") && atoi(getenv("S2TC_B200_TRACE"));
	const unsigned evflags = trace ? cudaEventDefault : cudaEventDisableTiming;
	std::vector<cudaEvent_t> up(nslab, nullptr), done(nslab, nullptr), down(trace ? nslab : 0, nullptr);
	std::vector<size_t> ws_off(nslab + 1, 0);
	cudaEvent_t t0 = nullptr;
	const bool pipelined = nslab > 1 || exchange;
	// random candidates: consecutive slabs alternate between the compute stream and the context's second stream (each with
	// its own workspaces), so that the tail of one slab's search launch overlaps the next slab's kernels; the DITHER_SIMPLE
	// carry still goes from slab to slab in order (every pre-pass waits for the one before it)
	const bool lanes = s.nrandom > 0 && nslab > 1 && !exchange && !c->profiling && st != c->aux && s.dither != kDitherFloyd;
	// Pageable caller memory (what Mesa and the reference's own tool pass to tx_compress_dxtn): the driver would stage it on
	// this thread, slab after slab, at a fraction of the link's speed and without overlapping anything.  Instead the texels
	// go through a ring of pinned chunks filled by a few copy threads while the GPU works on the previous slab, and the
	// blocks come back through a pinned buffer and are copied out as their slabs finish.
	static const bool no_stage = getenv("S2TC_B200_NO_STAGING") && atoi(getenv("S2TC_B200_NO_STAGING"));
	const bool stage_in = !exchange && !no_stage && in_bytes >= ((size_t) 1 << 20) && !host_pinned(src_rows);
	const bool stage_out = !exchange && !no_stage && row_bytes == tight && out_bytes >= ((size_t) 1 << 18) && !host_pinned(dest);
	if (stage_in && !c->h_stage_in) {
		CU(cudaHostAlloc((void **) &c->h_stage_in, kStageChunk * kStageSlots, cudaHostAllocDefault));
		for (int k = 0; k < kStageSlots; ++k)
			CU(cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
	}
	if (stage_out && c->h_stage_out_cap < out_bytes) {
		if (c->h_stage_out)
			cudaFreeHost(c->h_stage_out);
		c->h_stage_out = nullptr;
		c->h_stage_out_cap = 0;
		CU(cudaHostAlloc((void **) &c->h_stage_out, out_bytes, cudaHostAllocDefault));
		c->h_stage_out_cap = out_bytes;
	}
	std::vector<cudaEvent_t> landed(stage_out ? nslab : 0, nullptr); // slab i's blocks are in the pinned output buffer
	int drained = 0;                                               // slabs [0, drained) have been copied to dest
	size_t staged_chunks = 0;
	auto slab_rows = [&](int i, int &r0, int &r1) {
		if (first_rows > 0) { // growing slabs
			long long a = 0, sz = first_rows;
			for (int k = 0; k < i; ++k) {
				a += sz;
				sz = sz * 2 < cap_rows ? sz * 2 : cap_rows;
			}
			const long long b = i == nslab - 1 ? nrows : a + sz;
			r0 = row0 + (int) (a < nrows ? a : nrows);
			r1 = row0 + (int) (b < nrows ? b : nrows);
			return;
		}
		r0 = row0 + (int) ((long long) nrows * i / nslab);
		r1 = row0 + (int) ((long long) nrows * (i + 1) / nslab);
	};
	auto slab_texels = [&](int r0, int r1, size_t &off, size_t &len) {
		const int y0 = r0 * 4, y1 = r1 * 4 < height ? r1 * 4 : height;
		off = (size_t) (y0 - ty0) * width * comps;
		len = (size_t) (y1 - y0) * width * comps;
	};
	// everything that can fail after the first enqueue runs inside this lambda; the streams are always drained and the
	// events destroyed afterwards, so that no copy touches the caller's buffers once the call has returned
	const int rc = [&]() -> int {
		for (int i = 0; i < nslab; ++i) {
			CU(cudaEventCreateWithFlags(&up[i], evflags));
			CU(cudaEventCreateWithFlags(&done[i], evflags));
			if (trace)
				CU(cudaEventCreate(&down[i]));
		}
		if (pipelined || trace) {
			CU(cudaEventCreateWithFlags(&t0, evflags));
			CU(cudaEventRecord(t0, st));
			if (pipelined) { // the copy streams start behind whatever the compute stream was given before this call
				CU(cudaStreamWaitEvent(c->copy_in, t0, 0));
				CU(cudaStreamWaitEvent(c->copy_out, t0, 0));
			}
		}
		cudaStream_t sin_ = pipelined ? c->copy_in : st, sout = pipelined ? c->copy_out : st;
		if (lanes) // the second lane starts behind whatever the compute stream was given before this call, too
			CU(cudaStreamWaitEvent(c->aux, t0, 0));
		if (exchange) { // phase A: upload + summarise slab by slab
			for (int i = 0; i < nslab; ++i) {
				int r0, r1;
				size_t off, len;
				slab_rows(i, r0, r1);
				slab_texels(r0, r1, off, len);
				ws_off[i + 1] = ws_off[i] + ((dither_workspace_bytes(len / comps) + 255) & ~(size_t) 255);
			}
			CU(c->shard_ws.reserve(ws_off[nslab]));
			for (int i = 0; i < nslab; ++i) {
				int r0, r1;
				size_t off, len;
				slab_rows(i, r0, r1);
				slab_texels(r0, r1, off, len);
				CU(cudaMemcpyAsync((uint8_t *) c->src.p + off, src_rows + off, len, cudaMemcpyHostToDevice, sin_));
				CU(cudaEventRecord(up[i], sin_));
				CU(cudaStreamWaitEvent(st, up[i], 0));
				FamScope f(c, st, kFamPrepass, kDitherSummaryLaunches);
				CU(launch_dither_summary((const uint8_t *) c->src.p + off, comps, abits, len / comps, ex->d_mine + 4 * i,
						(uint8_t *) c->shard_ws.p + ws_off[i], st));
			}
			if (nslab < ex->nslab) { // a shard with fewer block rows than slabs: the missing ranges are empty
				FamScope f(c, st, kFamPrepass, 1);
				CU(launch_identity_maps(ex->d_mine + 4 * nslab, ex->nslab - nslab, comps, abits, st));
			}
			ex->gather(ex->user);
			FamScope f(c, st, kFamPrepass, 1);
			CU(launch_fold_carry(ex->d_all, ex->rank * ex->nslab, comps, abits, d_carry, st));
		}
		for (int i = 0; i < nslab; ++i) {
			int r0, r1;
			size_t off, len;
			slab_rows(i, r0, r1);
			slab_texels(r0, r1, off, len);
			cudaStream_t sl = lanes && (i & 1) ? c->aux : st; // this slab's stream
			if (!exchange) {
				if (stage_in) { // chunk by chunk through the pinned ring; this thread and the pool do the copying
					for (size_t o = 0; o < len; o += kStageChunk, ++staged_chunks) {
						const size_t n = len - o < kStageChunk ? len - o : kStageChunk;
						const int slot = (int) (staged_chunks % kStageSlots);
						if (staged_chunks >= (size_t) kStageSlots)
							CU(cudaEventSynchronize(c->stage_ev[slot])); // the chunk that used this slot is on the device
						uint8_t *h = c->h_stage_in + (size_t) slot * kStageChunk;
						CopyPool::get().copy(h, src_rows + off + o, n);
						CU(cudaMemcpyAsync((uint8_t *) c->src.p + off + o, h, n, cudaMemcpyHostToDevice, sin_));
						CU(cudaEventRecord(c->stage_ev[slot], sin_));
					}
				} else {
					CU(cudaMemcpyAsync((uint8_t *) c->src.p + off, src_rows + off, len, cudaMemcpyHostToDevice, sin_));
				}
				if (pipelined || trace) {
					CU(cudaEventRecord(up[i], sin_));
					CU(cudaStreamWaitEvent(sl, up[i], 0));
				}
			}
			if (lanes && i > 0 && s.dither == kDitherSimple)
				CU(cudaStreamWaitEvent(sl, c->ev_prepass, 0)); // the carry leaving slab i - 1
			uint8_t *d_out = (uint8_t *) c->out.p + (size_t) (r0 - row0) * tight;
			if (int e = encode_rows(c, s, comps, width, height, (const uint8_t *) c->src.p + off, r0, r1, d_out, cursor, d_carry, sl,
						exchange ? (uint8_t *) c->shard_ws.p + ws_off[i] : nullptr, false, lanes ? (i & 1) : -1))
				return e;
			if (pipelined || trace) {
				CU(cudaEventRecord(done[i], sl));
				CU(cudaStreamWaitEvent(sout, done[i], 0));
			}
			if (stage_out) {
				CU(cudaMemcpyAsync(c->h_stage_out + (size_t) (r0 - row0) * tight, d_out, (size_t) (r1 - r0) * tight, cudaMemcpyDeviceToHost, sout));
				CU(cudaEventCreateWithFlags(&landed[i], cudaEventDisableTiming));
				CU(cudaEventRecord(landed[i], sout));
				while (drained < i) { // finished slabs leave while we are here anyway
					if (cudaEventQuery(landed[drained]) != cudaSuccess) {
						cudaGetLastError(); // "not ready" is not an error
						break;
					}
					int a0, a1;
					slab_rows(drained, a0, a1);
					CopyPool::get().copy(dest + (size_t) (a0 - row0) * tight, c->h_stage_out + (size_t) (a0 - row0) * tight, (size_t) (a1 - a0) * tight);
					++drained;
				}
			} else if (row_bytes == tight)
				CU(cudaMemcpyAsync(dest + (size_t) (r0 - row0) * tight, d_out, (size_t) (r1 - r0) * tight, cudaMemcpyDeviceToHost, sout));
			else if (row_bytes > tight)
				CU(cudaMemcpy2DAsync(dest + (size_t) (r0 - row0) * row_bytes, row_bytes, d_out, tight, tight, r1 - r0, cudaMemcpyDeviceToHost, sout));
			if (trace)
				CU(cudaEventRecord(down[i], sout));
		}
		for (; stage_out && drained < nslab; ++drained) { // the rest, in order, each as soon as it has landed
			int a0, a1;
			slab_rows(drained, a0, a1);
			CU(cudaEventSynchronize(landed[drained]));
			CopyPool::get().copy(dest + (size_t) (a0 - row0) * tight, c->h_stage_out + (size_t) (a0 - row0) * tight, (size_t) (a1 - a0) * tight);
		}
		return 0;
	}();
	for (cudaEvent_t e : landed)
		if (e)
			cudaEventDestroy(e);
	cudaError_t e1 = cudaStreamSynchronize(st);
	const cudaError_t e2 = cudaStreamSynchronize(c->copy_in), e3 = cudaStreamSynchronize(c->copy_out), e4 = cudaStreamSynchronize(c->aux);
	if (e1 == cudaSuccess)
		e1 = e4;
	if (trace && !rc && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess) {
		for (int i = 0; i < nslab; ++i) {
			float a = 0, b = 0, d = 0;
			cudaEventElapsedTime(&a, t0, up[i]);
			cudaEventElapsedTime(&b, t0, done[i]);
			cudaEventElapsedTime(&d, t0, down[i]);
			fprintf(stderr, "s2tc_b200 trace: slab %2d/%d  uploaded %8.3f  encoded %8.3f  downloaded %8.3f ms\n", i, nslab, a, b, d);
		}
	}
	for (int i = 0; i < nslab; ++i) {
		if (up[i])
			cudaEventDestroy(up[i]);
		if (done[i])
			cudaEventDestroy(done[i]);
		if (trace && down[i])
			cudaEventDestroy(down[i]);
	}
	if (t0)
		cudaEventDestroy(t0);
	if (rc)
		return rc;
	CU(e1);
	CU(e2);
	CU(e3);
	if (row_bytes < tight) { // rows overlap in dest (stride between width*bs/4 and the padded width): later rows win, as in the reference
		std::vector<uint8_t> tmp(out_bytes);
		CU(cudaMemcpy(tmp.data(), c->out.p, out_bytes, cudaMemcpyDeviceToHost));
		for (int r = 0; r < nrows; ++r)
			memcpy(dest + (size_t) r * row_bytes, tmp.data() + (size_t) r * tight, tight);
	}
	return 0;
}

} // namespace

extern "C" {

int s2tc_b200_compress_host(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int srccomps, int width, int height,
		const uint8_t *src, uint8_t *dest, int dst_row_stride, uint64_t *rand_cursor)
{
	if (!c || !src || !dest)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	if (width <= 0 || height <= 0)
		return 0; // the reference's loops simply do not run
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	const int comps = srccomps == 3 ? 3 : 4;
	const int bw = (width + 3) / 4, bh = (height + 3) / 4, bs = block_bytes(s.dxt);
	const size_t tight = (size_t) bw * bs;
	const uint64_t cursor = rand_cursor ? *rand_cursor : 0;
	// ref s2tc_libtxc_dxtn.cpp:243,261,279: a stride below width*2 (DXT1) / width*4 (DXT3/5) means tight rows
	const size_t row_bytes = dst_row_stride >= width * (bs / 4) ? (size_t) dst_row_stride : tight;
	{
		StreamOrder order(c, c->stream);
		if (int rc = compress_rows_host(c, s, comps, width, height, src, 0, bh, dest, row_bytes, cursor, nullptr, c->stream))
			return rc;
	}
	if (rand_cursor && s.nrandom > 0)
		*rand_cursor = cursor + (uint64_t) bw * bh * draws_per_block(s.dxt, s.nrandom);
	return 0;
}

// One shard of an image that several contexts (GPUs, processes) encode together, host memory to host memory: block rows
// [row0, row1) of a width x height image; src_rows addresses texel row 4 * row0, dest the shard's first block (tight rows).
// rand_cursor0 is the cursor of the IMAGE's first block.  For DITHER_SIMPLE the shards must learn the carry entering
// them: d_maps_mine (nslab * 128 bytes, device) receives this shard's summaries, `gather(user)` is called once and must
// enqueue, on `stream`, an all-gather of every shard's d_maps_mine into d_maps_all (world * nslab * 128 bytes, device, in
// rank order) -- e.g. ncclAllGather / torch.distributed.all_gather_into_tensor.  nslab (1..64, the same on every shard) is
// also the number of pieces the shard is pipelined in.  With another dither mode gather is never called.
int s2tc_b200_compress_host_shard(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int srccomps, int width, int height,
		const uint8_t *src_rows, int row0, int row1, uint8_t *dest, uint64_t rand_cursor0, int rank, int nslab, void *d_maps_mine,
		void *d_maps_all, void (*gather)(void *), void *user, void *stream)
{
	if (!c || !src_rows || !dest)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	const int bh = (height + 3) / 4;
	if (width <= 0 || height <= 0 || row0 < 0 || row1 > bh || row0 > row1 || rank < 0)
		return fail(S2TC_B200_EINVAL, "bad geometry %dx%d rows [%d,%d) rank %d", width, height, row0, row1, rank);
	const bool exchange = s.dither == kDitherSimple && gather;
	if (exchange && (!d_maps_mine || !d_maps_all || nslab < 1 || nslab > 64))
		return fail(S2TC_B200_EINVAL, "DITHER_SIMPLE shards need d_maps_mine, d_maps_all and 1 <= nslab <= 64");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	const int comps = srccomps == 3 ? 3 : 4;
	if (row0 == row1) { // an empty shard still takes part in the exchange
		if (exchange) {
			FamScope f(c, st, kFamPrepass, 1);
			CU(launch_identity_maps((ByteMap *) d_maps_mine, nslab, comps, alpha_bits(s.dxt), st));
			gather(user);
			CU(cudaStreamSynchronize(st));
		}
		return 0;
	}
	CarryExchange ex{rank, nslab, (ByteMap *) d_maps_mine, (ByteMap *) d_maps_all, gather, user};
	const size_t tight = (size_t) ((width + 3) / 4) * block_bytes(s.dxt);
	return compress_rows_host(c, s, comps, width, height, src_rows, row0, row1, dest, tight, rand_cursor0, exchange ? &ex : nullptr, st);
}

// DITHER_FLOYDSTEINBERG for row shards: see include/s2tc_b200.h
int s2tc_b200_floyd_rows_device(s2tc_b200_ctx *c, int srccomps, int alphabits, int width, int height, const void *d_src_rows, int row0,
		int row1, int phase, const int *d_err_in, int *d_err_out, void *d_reduced_rows, void *stream)
{
	if (!c || !d_src_rows || !d_reduced_rows)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	const int bh = (height + 3) / 4;
	if (width <= 0 || height <= 0 || row0 < 0 || row1 > bh || row0 >= row1 || (phase != 0 && phase != 1))
		return fail(S2TC_B200_EINVAL, "bad geometry %dx%d rows [%d,%d) phase %d", width, height, row0, row1, phase);
	const int comps = srccomps == 3 ? 3 : 4;
	if (phase == 1 && (comps != 4 || (alphabits != 1 && alphabits != 4)))
		return fail(S2TC_B200_EINVAL, "the alpha pass exists for 4-component sources and 1- or 4-bit alpha only (ref s2tc_algorithm.cpp:1374-1404)");
	if (alphabits != 1 && alphabits != 4 && alphabits != 8)
		return fail(S2TC_B200_EINVAL, "alphabits %d", alphabits);
	const int y0 = row0 * 4, y1 = row1 * 4 < height ? row1 * 4 : height;
	if (row1 < bh && !d_err_out)
		return fail(S2TC_B200_EINVAL, "rows [%d,%d) of %d have rows below them: d_err_out is needed", row0, row1, bh);
	if (phase == 1 && !d_err_in)
		return fail(S2TC_B200_EINVAL, "the alpha pass needs d_err_in (first rows: the seed the last rows' colour pass left)");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	CU(c->dither_ws.reserve(floyd_workspace_bytes(width, y1 - y0)));
	FamScope f(c, st, kFamPrepass, row1 < bh ? 2 : 1);
	CU(launch_floyd_rows(d_src_rows, comps, alphabits, width, height, y0, y1 - y0, phase, d_err_in, d_err_out, d_reduced_rows,
			c->dither_ws.p, st));
	return 0;
}

// Block rows [row0, row1) of an image whose 565 pre-pass the caller has already run: d_reduced_rows = reduced texels
// {r5, g6, b5, a}, 4 bytes each, of texel row 4 * row0 onwards.  s->dither is ignored.
int s2tc_b200_encode_reduced_rows_device(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int width, int height,
		const void *d_reduced_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, void *stream)
{
	if (!c || !d_reduced_rows || !d_dst)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	return encode_rows(c, s, 4, width, height, d_reduced_rows, row0, row1, d_dst, rand_cursor0, nullptr, st, nullptr, true);
}

void s2tc_b200_stripe_rows(int height, int world, int nwave, const int *wave_weights, int wave, int rank, int *row0, int *row1)
{
	// wave w covers block rows [bh * cum(w) / total, bh * cum(w + 1) / total) and is cut evenly among the shards
	const long long bh = (height + 3) / 4;
	long long total = 0, before = 0, mine = 1;
	for (int w = 0; w < nwave; ++w) {
		const long long x = wave_weights ? (wave_weights[w] > 0 ? wave_weights[w] : 0) : 1;
		if (w < wave)
			before += x;
		if (w == wave)
			mine = x;
		total += x;
	}
	if (total <= 0 || wave < 0 || wave >= nwave) {
		*row0 = *row1 = 0;
		return;
	}
	const long long a = bh * before / total, b = bh * (before + mine) / total;
	*row0 = (int) (a + (b - a) * rank / world);
	*row1 = (int) (a + (b - a) * (rank + 1) / world);
}

// One texture, `world` shards, STRIPED: the block rows are cut into nwave * world stripes of consecutive block rows and
// stripe wave * world + rank belongs to shard `rank` (s2tc_b200_stripe_rows).  Why not one contiguous range per shard
// (s2tc_b200_compress_host_shard): with DITHER_SIMPLE a shard cannot start before every shard above it is uploaded and
// summarised, and all shards upload at the same pace -- nothing is encoded until everything is on the devices.  Striped,
// wave w of every shard lands at about the same time, its summaries are exchanged (128 bytes per shard and wave) and the
// stripes are encoded while wave w + 1 is on its way: uploads, kernels and downloads overlap on every GPU.
int s2tc_b200_compress_host_striped(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int srccomps, int width, int height,
		const uint8_t *const *src_stripes, uint8_t *const *dest_stripes, uint64_t rand_cursor0, int rank, int world, int nwave,
		const int *wave_weights, void *d_maps_mine, void *d_maps_all, void (*gather)(void *, int), void *user, void *stream)
{
	if (!c || !src_stripes || !dest_stripes)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	if (width <= 0 || height <= 0 || world < 1 || rank < 0 || rank >= world || nwave < 1 || nwave > 256)
		return fail(S2TC_B200_EINVAL, "bad geometry %dx%d rank %d of %d, %d waves", width, height, rank, world, nwave);
	if (s.dither == kDitherFloyd)
		return fail(S2TC_B200_EUNSUPPORTED, "DITHER_FLOYDSTEINBERG diffuses error between rows: use s2tc_b200_floyd_rows_device per shard");
	const bool simple = s.dither == kDitherSimple;
	const bool exchange = simple && world > 1;
	if (exchange && (!gather || !d_maps_mine || !d_maps_all))
		return fail(S2TC_B200_EINVAL, "DITHER_SIMPLE stripes of several shards need gather, d_maps_mine and d_maps_all");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	const int comps = srccomps == 3 ? 3 : 4, abits = alpha_bits(s.dxt);
	const size_t tight = (size_t) ((width + 3) / 4) * block_bytes(s.dxt);

	struct Stripe {
		int r0, r1;
		size_t in_off, in_len, out_off, out_len, ws_off;
	};
	std::vector<Stripe> sp(nwave);
	size_t in_total = 0, out_total = 0, ws_total = 0;
	for (int w = 0; w < nwave; ++w) {
		Stripe &q = sp[w];
		s2tc_b200_stripe_rows(height, world, nwave, wave_weights, w, rank, &q.r0, &q.r1);
		const int y0 = q.r0 * 4, y1 = q.r1 * 4 < height ? q.r1 * 4 : height;
		q.in_off = in_total;
		q.in_len = (size_t) width * (y1 - y0) * comps;
		q.out_off = out_total;
		q.out_len = tight * (q.r1 - q.r0);
		q.ws_off = ws_total;
		in_total += (q.in_len + 255) & ~(size_t) 255;
		out_total += (q.out_len + 255) & ~(size_t) 255;
		if (simple && q.in_len)
			ws_total += (dither_workspace_bytes(q.in_len / comps) + 255) & ~(size_t) 255;
		if (q.in_len && (!src_stripes[w] || !dest_stripes[w]))
			return fail(S2TC_B200_EINVAL, "stripe %d has no source or destination", w);
	}
	CU(c->src.reserve(in_total ? in_total : 1));
	CU(c->out.reserve(out_total ? out_total : 1));
	if (ws_total)
		CU(c->shard_ws.reserve(ws_total));
	int *d_wave = (int *) c->small.p + 24;                     // carry entering the current wave's first stripe
	// The control work of every wave (summary, all-gather, folds) runs on st.  With random candidates the stripes themselves
	// alternate between the context's two lane streams, each behind its wave's control work, so that the tail of one
	// stripe's search launch overlaps the next stripe (compress_rows_host) and the control work of wave w + 1 does not wait
	// for the kernels of wave w.  Every stripe has its own carry slot: fold(w + 1) must not overwrite what encode(w) reads.
	const bool lanes = s.nrandom > 0 && nwave > 1 && !c->profiling && st != c->aux && st != c->aux2;
	ByteMap *d_mine = (ByteMap *) d_maps_mine, *d_all = (ByteMap *) d_maps_all;
	std::vector<cudaEvent_t> up(nwave, nullptr), done(nwave, nullptr);
	cudaEvent_t t0 = nullptr;
	const int rc = [&]() -> int {
		for (int w = 0; w < nwave; ++w) {
			CU(cudaEventCreateWithFlags(&up[w], cudaEventDisableTiming));
			CU(cudaEventCreateWithFlags(&done[w], cudaEventDisableTiming));
		}
		CU(cudaEventCreateWithFlags(&t0, cudaEventDisableTiming));
		CU(cudaEventRecord(t0, st));
		CU(cudaStreamWaitEvent(c->copy_in, t0, 0));
		CU(cudaStreamWaitEvent(c->copy_out, t0, 0));
		if (simple)
			CU(cudaMemsetAsync(exchange ? d_wave : (int *) c->small.p + 8, 0, 4 * sizeof(int), st));
		for (int w = 0; w < nwave; ++w) { // all uploads are queued at once; the compute stream takes them as they land
			if (sp[w].in_len)
				CU(cudaMemcpyAsync((uint8_t *) c->src.p + sp[w].in_off, src_stripes[w], sp[w].in_len, cudaMemcpyHostToDevice, c->copy_in));
			CU(cudaEventRecord(up[w], c->copy_in));
		}
		for (int w = 0; w < nwave; ++w) {
			const Stripe &q = sp[w];
			const uint8_t *d_in = (const uint8_t *) c->src.p + q.in_off;
			uint8_t *ws = ws_total ? (uint8_t *) c->shard_ws.p + q.ws_off : nullptr;
			CU(cudaStreamWaitEvent(st, up[w], 0));
			int *d_carry = simple ? (exchange ? (int *) c->carries.p + 4 * w : (int *) c->small.p + 8) : nullptr; // carry entering this stripe
			if (exchange) {
				{
					FamScope f(c, st, kFamPrepass, q.in_len ? kDitherSummaryLaunches : 1);
					if (q.in_len)
						CU(launch_dither_summary(d_in, comps, abits, q.in_len / comps, d_mine + 4 * w, ws, st));
					else
						CU(launch_identity_maps(d_mine + 4 * w, 1, comps, abits, st));
				}
				gather(user, w); // d_mine[w] of every shard -> d_all[w * world ..], on st
				FamScope f(c, st, kFamPrepass, 2);
				const ByteMap *wave_maps = d_all + (size_t) 4 * world * w;
				CU(launch_fold_carry_from(wave_maps, rank, comps, abits, d_wave, d_carry, st));
				CU(launch_fold_carry_from(wave_maps, world, comps, abits, d_wave, d_wave, st));
			}
			if (q.r1 > q.r0) {
				uint8_t *d_out = (uint8_t *) c->out.p + q.out_off;
				cudaStream_t sl = st;
				if (lanes) { // behind this wave's upload and folds
					sl = (w & 1) ? c->aux : c->aux2;
					CU(cudaEventRecord(c->ev_fork, st));
					CU(cudaStreamWaitEvent(sl, c->ev_fork, 0));
				}
				if (lanes && w > 0 && simple && !exchange)
					CU(cudaStreamWaitEvent(sl, c->ev_prepass, 0)); // one shard: the carry is chained from stripe to stripe
				if (int e = encode_rows(c, s, comps, width, height, d_in, q.r0, q.r1, d_out, rand_cursor0, d_carry, sl, exchange ? ws : nullptr,
							false, lanes ? (w & 1) : -1))
					return e;
				CU(cudaEventRecord(done[w], sl));
				CU(cudaStreamWaitEvent(c->copy_out, done[w], 0));
				CU(cudaMemcpyAsync(dest_stripes[w], d_out, q.out_len, cudaMemcpyDeviceToHost, c->copy_out));
			}
		}
		return 0;
	}();
	cudaError_t e1 = cudaStreamSynchronize(st);
	const cudaError_t e2 = cudaStreamSynchronize(c->copy_in), e3 = cudaStreamSynchronize(c->copy_out);
	const cudaError_t e4 = cudaStreamSynchronize(c->aux), e5 = cudaStreamSynchronize(c->aux2);
	if (e1 == cudaSuccess)
		e1 = e4 != cudaSuccess ? e4 : e5;
	for (int w = 0; w < nwave; ++w) {
		if (up[w])
			cudaEventDestroy(up[w]);
		if (done[w])
			cudaEventDestroy(done[w]);
	}
	if (t0)
		cudaEventDestroy(t0);
	if (rc)
		return rc;
	CU(e1);
	CU(e2);
	CU(e3);
	return 0;
}

size_t s2tc_b200_mipchain_bytes(int dxt, int width, int height)
{
	const int bs = block_bytes(norm_dxt(dxt));
	size_t total = 0;
	for (int w = width, h = height; w > 0 && h > 0;) {
		total += (size_t) ((w + 3) / 4) * ((h + 3) / 4) * bs;
		if (w == 1 && h == 1)
			break;
		w = w > 1 ? w >> 1 : w;
		h = h > 1 ? h >> 1 : h;
	}
	return total;
}

int s2tc_b200_mip_reduce_device(s2tc_b200_ctx *c, const void *d_in, int width, int height, void *d_out, void *stream)
{
	if (!c || !d_in || !d_out || width <= 0 || height <= 0)
		return fail(S2TC_B200_EINVAL, "bad argument");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	FamScope f(c, st, kFamPrepass, 1);
	CU(launch_mip_reduce(d_in, width, height, d_out, st));
	return 0;
}

int s2tc_b200_compress_mipchain_device(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int width, int height, void *d_rgba,
		void *d_scratch, void *d_dst, uint64_t *rand_cursor, void *stream)
{
	if (!c || !d_rgba || !d_scratch || !d_dst)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (width <= 0 || height <= 0)
		return fail(S2TC_B200_EINVAL, "bad size %dx%d", width, height);
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	uint64_t cursor = rand_cursor ? *rand_cursor : 0;
	const int bs = block_bytes(s.dxt);
	uint8_t *cur = (uint8_t *) d_rgba, *other = (uint8_t *) d_scratch, *dst = (uint8_t *) d_dst;
	for (int w = width, h = height;;) { // ref s2tc_compress.c:722-733: one tx_compress_dxtn per level, down to 1x1
		const int bw = (w + 3) / 4, bh = (h + 3) / 4;
		if (int rc = encode_rows(c, s, 4, w, h, cur, 0, bh, dst, cursor, nullptr, st))
			return rc;
		dst += (size_t) bw * bh * bs;
		cursor += (uint64_t) bw * bh * draws_per_block(s.dxt, s.nrandom);
		if (w == 1 && h == 1)
			break;
		{
			FamScope f(c, st, kFamPrepass, 1);
			CU(launch_mip_reduce(cur, w, h, other, st));
		}
		uint8_t *t = cur; cur = other; other = t;
		w = w > 1 ? w >> 1 : w;
		h = h > 1 ? h >> 1 : h;
	}
	if (rand_cursor)
		*rand_cursor = cursor;
	return 0;
}

// A batch of equally sized textures, every mip level of all of them per launch (SURVEY "next" N2 / BASELINE config 4):
// ntex RGBA8 textures back to back in d_rgba (not modified); for each of the nset settings the whole chain of every
// texture is encoded.  The chain of texture i under setting k goes to
//     d_dst + off_k + i * mipchain_bytes(dxt_k),   off_0 = 0,  off_{k+1} = (off_k + ntex * mipchain_bytes(dxt_k)) rounded up to 16.
// d_scratch: ntex * (width*height + width*height/4) + 256 bytes (the two mip buffers).  Each texture is its own run of
// the reference tool (s2tc_compress.c:722-733): the DITHER_SIMPLE carry restarts at every level and every texture's
// rand() cursor starts at rand_cursor0.  The 565 pre-pass of a level is shared by all settings that agree on the dither
// mode and the alpha width.  Launches per call: about levels * (6 + nset * 1..3), independent of ntex (nrandom <= 0).
int s2tc_b200_compress_mipchain_batch_device(s2tc_b200_ctx *c, const s2tc_b200_settings *sets, int nset, int width, int height, int ntex,
		const void *d_rgba, void *d_scratch, void *d_dst, uint64_t rand_cursor0, void *stream)
{
	if (!c || !sets || !d_rgba || !d_scratch || !d_dst)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (width <= 0 || height <= 0 || ntex <= 0 || nset <= 0)
		return fail(S2TC_B200_EINVAL, "bad batch %d x %dx%d, %d settings", ntex, width, height, nset);
	std::vector<s2tc_b200_settings> ss(nset);
	std::vector<size_t> set_off(nset + 1, 0), chain(nset);
	for (int k = 0; k < nset; ++k) {
		if (int rc = settings_normalise(&sets[k], ss[k]))
			return rc;
		chain[k] = s2tc_b200_mipchain_bytes(ss[k].dxt, width, height);
		set_off[k + 1] = (set_off[k] + chain[k] * ntex + 15) & ~(size_t) 15; // 16-byte blocks are stored with 128-bit stores
	}
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	const size_t level1 = (size_t) ntex * (width > 1 ? width >> 1 : width) * (height > 1 ? height >> 1 : height) * 4;
	const uint8_t *cur = (const uint8_t *) d_rgba;
	uint8_t *bufs[2] = {(uint8_t *) d_scratch, (uint8_t *) d_scratch + ((level1 + 255) & ~(size_t) 255)};
	int *zero_carry = (int *) c->small.p + 16;
	CU(cudaMemsetAsync(zero_carry, 0, 4 * sizeof(int), st));
	std::vector<size_t> level_off(nset, 0);
	std::vector<uint64_t> cursor(nset, rand_cursor0);
	int flip = 0;
	for (int w = width, h = height;;) {
		const int bw = (w + 3) / 4, bh = (h + 3) / 4;
		const size_t npix = (size_t) w * h;
		int have_dither = -1, have_abits = -1; // what c->reduced currently holds for this level
		for (int k = 0; k < nset; ++k) {
			const s2tc_b200_settings &s = ss[k];
			const int abits = alpha_bits(s.dxt), bs = block_bytes(s.dxt);
			const uint8_t *texels = cur;
			int fmt = kSrcRGBA8;
			if (s.dither != kDitherNone) {
				if (have_dither != s.dither || have_abits != abits) {
					CU(c->reduced.reserve(npix * 4 * ntex));
					if (s.dither == kDitherSimple) {
						CU(c->dither_ws.reserve(dither_workspace_bytes(npix * ntex)));
						FamScope f(c, st, kFamPrepass, prepass_simple_batch_launches(npix, ntex));
						CU(launch_prepass_simple_batch(cur, 4, abits, npix, ntex, c->reduced.p, zero_carry, c->dither_ws.p, st));
						if (npix > 16384 && npix % 16384)
							CU(cudaMemsetAsync(zero_carry, 0, 4 * sizeof(int), st));
					} else {
						CU(c->dither_ws.reserve(floyd_workspace_bytes(w, h)));
						FamScope f(c, st, kFamPrepass, ntex * 2);
						for (int i = 0; i < ntex; ++i)
							CU(launch_prepass_floyd(cur + (size_t) i * npix * 4, 4, abits, w, h, (uint8_t *) c->reduced.p + (size_t) i * npix * 4,
									c->dither_ws.p, st));
					}
					have_dither = s.dither;
					have_abits = abits;
				}
				texels = (const uint8_t *) c->reduced.p;
				fmt = kSrcReduced;
			}
			const ImageView v = make_batch_view(texels, w, h, fmt, abits, ntex, npix * 4, chain[k]);
			if (int rc = encode_view(c, s, v, 0, cursor[k], (uint8_t *) d_dst + set_off[k] + level_off[k], st))
				return rc;
			level_off[k] += (size_t) bw * bh * bs;
			cursor[k] += (uint64_t) bw * bh * draws_per_block(s.dxt, s.nrandom);
		}
		if (w == 1 && h == 1)
			break;
		{
			FamScope f(c, st, kFamPrepass, 1);
			CU(launch_mip_reduce(cur, w, h, bufs[flip], st, ntex));
		}
		cur = bufs[flip];
		flip ^= 1;
		w = w > 1 ? w >> 1 : w;
		h = h > 1 ? h >> 1 : h;
	}
	return 0;
}

int s2tc_b200_compress_mipchain_host(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, int width, int height, const uint8_t *rgba,
		uint8_t *dest, uint64_t *rand_cursor)
{
	if (!c || !rgba || !dest)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (width <= 0 || height <= 0)
		return fail(S2TC_B200_EINVAL, "bad size %dx%d", width, height);
	const size_t in_bytes = (size_t) width * height * 4;
	const size_t out_bytes = s2tc_b200_mipchain_bytes(sin ? sin->dxt : 0, width, height);
	void *d_in, *d_tmp, *d_out;
	{
		std::lock_guard<std::mutex> lock(c->mu);
		CU(cudaSetDevice(c->device));
		CU(c->src.reserve(in_bytes));
		CU(c->mip.reserve(in_bytes / 4 + 64));
		CU(c->out.reserve(out_bytes));
		d_in = c->src.p; d_tmp = c->mip.p; d_out = c->out.p;
		CU(cudaMemcpyAsync(d_in, rgba, in_bytes, cudaMemcpyHostToDevice, c->stream));
	}
	if (int rc = s2tc_b200_compress_mipchain_device(c, sin, width, height, d_in, d_tmp, d_out, rand_cursor, nullptr))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaMemcpyAsync(dest, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

int s2tc_b200_rgb565_host(s2tc_b200_ctx *c, uint8_t *out, const uint8_t *src, int width, int height, int srccomps,
		int alphabits, int dither)
{
	if (!c || !out || !src)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (width <= 0 || height <= 0)
		return 0;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	StreamOrder order(c, st);
	const int comps = srccomps == 3 ? 3 : 4;
	const int abits = (alphabits == 1 || alphabits == 4) ? alphabits : 8; // ref s2tc_algorithm.cpp:1437-1449
	const size_t npix = (size_t) width * height;
	CU(c->src.reserve(npix * comps));
	CU(c->reduced.reserve(npix * 4));
	CU(cudaMemcpyAsync(c->src.p, src, npix * comps, cudaMemcpyHostToDevice, st));
	if (dither == kDitherNone) {
		FamScope f(c, st, kFamPrepass, 1);
		CU(launch_prepass_none(c->src.p, comps, abits, npix, c->reduced.p, st));
	} else if (dither == kDitherFloyd) {
		CU(c->dither_ws.reserve(floyd_workspace_bytes(width, height)));
		FamScope f(c, st, kFamPrepass, 2);
		CU(launch_prepass_floyd(c->src.p, comps, abits, width, height, c->reduced.p, c->dither_ws.p, st));
	} else {
		CU(c->dither_ws.reserve(dither_workspace_bytes(npix)));
		int *carry = (int *) c->small.p;
		CU(cudaMemsetAsync(carry, 0, 4 * sizeof(int), st));
		FamScope f(c, st, kFamPrepass, prepass_simple_launches(npix, false));
		CU(launch_prepass_simple(c->src.p, comps, abits, npix, c->reduced.p, carry, c->dither_ws.p, false, st));
	}
	CU(cudaMemcpyAsync(out, c->reduced.p, npix * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

int s2tc_b200_encode_block_host(s2tc_b200_ctx *c, const s2tc_b200_settings *sin, uint8_t *out, const uint8_t *rgba, int iw,
		int w, int h, uint64_t *rand_cursor)
{
	if (!c || !out || !rgba)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (w < 1 || w > 4 || h < 1 || h > 4)
		return fail(S2TC_B200_EINVAL, "block extent %dx%d", w, h);
	s2tc_b200_settings s;
	if (int rc = settings_normalise(sin, s))
		return rc;
	s.dither = kDitherNone; // input is already reduced
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	StreamOrder order(c, st);
	for (int y = 0; y < h; ++y)
		memcpy(c->h_block + (size_t) y * w * 4, rgba + (size_t) y * iw * 4, (size_t) w * 4);
	uint8_t *d_px = (uint8_t *) c->small.p + 512, *d_out = d_px + 64;
	CU(cudaMemcpyAsync(d_px, c->h_block, (size_t) w * h * 4, cudaMemcpyHostToDevice, st));
	const ImageView v = make_view(d_px, w, h, kSrcReduced, alpha_bits(s.dxt));
	const uint64_t cursor = rand_cursor ? *rand_cursor : 0;
	if (int rc = encode_view(c, s, v, 0, cursor, d_out, st))
		return rc;
	const int bs = block_bytes(s.dxt);
	CU(cudaMemcpyAsync(c->h_block + 64, d_out, bs, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	memcpy(out, c->h_block + 64, bs);
	if (rand_cursor && s.nrandom > 0)
		*rand_cursor = cursor + draws_per_block(s.dxt, s.nrandom);
	return 0;
}

int s2tc_b200_transcode_device(s2tc_b200_ctx *c, int dxt, void *d_blocks, size_t nblocks, void *stream)
{
	if (!c || !d_blocks)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (dxt != kDxt1 && dxt != kDxt3 && dxt != kDxt5)
		return fail(S2TC_B200_EINVAL, "bad dxt %d", dxt);
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	FamScope f(c, st, kFamTranscode, 1);
	CU(launch_transcode(dxt, d_blocks, nblocks, st));
	return 0;
}

int s2tc_b200_transcode_host(s2tc_b200_ctx *c, int dxt, uint8_t *blocks, size_t nblocks)
{
	if (!c || !blocks)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if (dxt != kDxt1 && dxt != kDxt3 && dxt != kDxt5)
		return fail(S2TC_B200_EINVAL, "bad dxt %d", dxt);
	if (!nblocks)
		return 0;
	const size_t bytes = nblocks * block_bytes(dxt);
	{
		std::lock_guard<std::mutex> lock(c->mu);
		CU(cudaSetDevice(c->device));
		CU(c->out.reserve(bytes));
		CU(cudaMemcpyAsync(c->out.p, blocks, bytes, cudaMemcpyHostToDevice, c->stream));
	}
	if (int rc = s2tc_b200_transcode_device(c, dxt, c->out.p, nblocks, nullptr))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaMemcpyAsync(blocks, c->out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

int s2tc_b200_decode_device(s2tc_b200_ctx *c, int dxt, const void *d_blocks, int width, int height, void *d_rgba, void *stream)
{
	if (!c || !d_blocks || !d_rgba)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if ((dxt != kDxt1 && dxt != kDxt3 && dxt != kDxt5) || width <= 0 || height <= 0)
		return fail(S2TC_B200_EINVAL, "bad argument");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
	StreamOrder order(c, st);
	FamScope f(c, st, kFamTranscode, 1);
	CU(launch_decode(dxt, d_blocks, width, height, d_rgba, st));
	return 0;
}

int s2tc_b200_decode_host(s2tc_b200_ctx *c, int dxt, const uint8_t *blocks, int width, int height, uint8_t *rgba)
{
	if (!c || !blocks || !rgba)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	if ((dxt != kDxt1 && dxt != kDxt3 && dxt != kDxt5) || width <= 0 || height <= 0)
		return fail(S2TC_B200_EINVAL, "bad argument");
	const size_t nb = (size_t) ((width + 3) / 4) * ((height + 3) / 4) * block_bytes(dxt), np = (size_t) width * height * 4;
	void *d_b, *d_p;
	{
		std::lock_guard<std::mutex> lock(c->mu);
		CU(cudaSetDevice(c->device));
		CU(c->out.reserve(nb));
		CU(c->src.reserve(np));
		d_b = c->out.p; d_p = c->src.p;
		CU(cudaMemcpyAsync(d_b, blocks, nb, cudaMemcpyHostToDevice, c->stream));
	}
	if (int rc = s2tc_b200_decode_device(c, dxt, d_b, width, height, d_p, nullptr))
		return rc;
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaMemcpyAsync(rgba, d_p, np, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

static std::mutex g_cursor_mu;
static uint64_t g_cursor = 0;
uint64_t s2tc_b200_rand_cursor_get(void)
{
	std::lock_guard<std::mutex> lock(g_cursor_mu);
	return g_cursor;
}
void s2tc_b200_rand_cursor_set(uint64_t draws)
{
	std::lock_guard<std::mutex> lock(g_cursor_mu);
	g_cursor = draws;
}

int s2tc_b200_sync(s2tc_b200_ctx *c)
{
	if (!c)
		return fail(S2TC_B200_EINVAL, "NULL context");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	return 0;
}

uint64_t s2tc_b200_launch_count(s2tc_b200_ctx *c) { return c ? c->launches : 0; }

static int int_peak_mode(s2tc_b200_ctx *c, int mode, double ops_per_step, double *gops)
{
	const int iters = 1 << 14, ctas = 148 * 16;
	int *sink = (int *) c->small.p + 200;
	cudaEvent_t a, b;
	CU(cudaEventCreate(&a));
	CU(cudaEventCreate(&b));
	double best = 0;
	for (int rep = 0; rep < 4; ++rep) { // first repetition warms up
		CU(cudaEventRecord(a, c->stream));
		CU(launch_int32_peak(mode, iters, ctas, sink, c->stream));
		CU(cudaEventRecord(b, c->stream));
		CU(cudaEventSynchronize(b));
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, a, b));
		const double ops = ops_per_step * 8.0 * iters * 256.0 * ctas;
		if (rep && ms > 0 && ops / ms * 1e-6 > best)
			best = ops / ms * 1e-6;
	}
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	*gops = best;
	return 0;
}

int s2tc_b200_int_peaks(s2tc_b200_ctx *c, double *scalar_gops, double *packed16_gops)
{
	if (!c || !scalar_gops || !packed16_gops)
		return fail(S2TC_B200_EINVAL, "NULL argument");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	double plain = 0, dual = 0;
	if (int rc = int_peak_mode(c, 0, 2.0, &plain))
		return rc;
	if (int rc = int_peak_mode(c, 1, 2.0, &dual))
		return rc;
	*scalar_gops = plain > dual ? plain : dual;
	return int_peak_mode(c, 2, 4.0, packed16_gops);
}

int s2tc_b200_int32_peak(s2tc_b200_ctx *c, double *gops)
{
	double packed;
	return s2tc_b200_int_peaks(c, gops, &packed);
}

int s2tc_b200_profile_enable(s2tc_b200_ctx *c, int on)
{
	if (!c)
		return fail(S2TC_B200_EINVAL, "NULL context");
	std::lock_guard<std::mutex> lock(c->mu);
	c->profiling = on != 0;
	return 0;
}

int s2tc_b200_profile_read(s2tc_b200_ctx *c, double ms[6], uint64_t launches[6], int reset)
{
	if (!c)
		return fail(S2TC_B200_EINVAL, "NULL context");
	std::lock_guard<std::mutex> lock(c->mu);
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	for (auto &p : c->pending) {
		float t = 0;
		if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess)
			c->fam_ms[p.fam] += t;
		cudaEventDestroy(p.a);
		cudaEventDestroy(p.b);
	}
	c->pending.clear();
	for (int i = 0; i < kNumFam; ++i) {
		if (ms)
			ms[i] = c->fam_ms[i];
		if (launches)
			launches[i] = c->fam_launches[i];
		if (reset) {
			c->fam_ms[i] = 0;
			c->fam_launches[i] = 0;
		}
	}
	return 0;
}

} // extern "C"
