/*
 * s2tc_b200.h -- C ABI of the B200 (sm_100a) S2TC encoder.
 *
 * This is the boundary a host program binds (dlopen/dlsym, cgo, JNI, ctypes ...): plain pointers,
 * sizes and ints, no C++ or torch types.  It has two layers:
 *
 *  1. the reference's own entry points, re-exported unchanged so that the library is a drop-in for
 *     libtxc_dxtn.so:   tx_compress_dxtn, fetch_2d_texel_*            -> include/txc_dxtn.h
 *                       s2tc_encode_block_func, rgb565_image           -> include/s2tc_algorithm.h
 *  2. the batched/device-pointer calls below (prefix s2tc_b200_), which is what those entry points
 *     are built on and what a caller that already has texels in GPU memory should use.
 *
 * Every function returns 0 on success or a negative S2TC_B200_E* code; s2tc_b200_last_error() gives
 * the text of the most recent failure on the calling thread.  There is NO CPU fallback: without a
 * usable CUDA device the calls fail with S2TC_B200_ENODEVICE.
 *
 * Enumerator values are the reference's (s2tc_algorithm.h:31-63).
 */
#ifndef S2TC_B200_H
#define S2TC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { S2TC_B200_DITHER_NONE = 0, S2TC_B200_DITHER_SIMPLE = 1, S2TC_B200_DITHER_FLOYDSTEINBERG = 2 };
enum { S2TC_B200_DXT1 = 0, S2TC_B200_DXT3 = 1, S2TC_B200_DXT5 = 2 };
enum { S2TC_B200_REFINE_NEVER = 0, S2TC_B200_REFINE_ALWAYS = 1, S2TC_B200_REFINE_LOOP = 2 };
enum {
	S2TC_B200_RGB = 0, S2TC_B200_YUV, S2TC_B200_SRGB, S2TC_B200_SRGB_MIXED,
	S2TC_B200_AVG, S2TC_B200_WAVG, S2TC_B200_W0AVG, S2TC_B200_NORMALMAP
};

enum {
	S2TC_B200_OK = 0,
	S2TC_B200_ENODEVICE = -1, /* no CUDA device / driver, or the device is not usable */
	S2TC_B200_EINVAL = -2,    /* bad argument (format, sizes, null pointer) */
	S2TC_B200_ECUDA = -3,     /* a CUDA call failed; see s2tc_b200_last_error() */
	S2TC_B200_ENOMEM = -4,
	S2TC_B200_EUNSUPPORTED = -5 /* valid request this build cannot run on the device (stated in the message) */
};

/* Settings of one encode.  What tx_compress_dxtn reads from its arguments and from the S2TC_*
 * environment (reference s2tc_libtxc_dxtn.cpp:156-231), made explicit. */
typedef struct s2tc_b200_settings {
	int dxt;     /* S2TC_B200_DXT1/3/5                       (from destformat, ref :218-231) */
	int cd;      /* colour metric, S2TC_COLORDIST_MODE       (ref :175-197, default WAVG) */
	int nrandom; /* S2TC_RANDOM_COLORS                       (ref :199-202, default -1 = fast mode) */
	int refine;  /* S2TC_REFINE_COLORS                       (ref :204-216, default ALWAYS) */
	int dither;  /* S2TC_DITHER_MODE                         (ref :161-173, default SIMPLE) */
} s2tc_b200_settings;

/* One per (thread, device); owns a stream and the workspaces every call uses.  Calls on one context are serialised on the
 * host by a mutex and ordered on the device: a call given a different stream than the previous call first waits (on the
 * device) for that call's work, because the workspaces are shared.  Use one context per concurrent lane of work. */
typedef struct s2tc_b200_ctx s2tc_b200_ctx;

const char *s2tc_b200_last_error(void);
int s2tc_b200_device_count(void);

int s2tc_b200_ctx_create(int device, s2tc_b200_ctx **out);
void s2tc_b200_ctx_destroy(s2tc_b200_ctx *ctx);
/* process-wide lazily created context on device S2TC_B200_DEVICE (default 0): what the
 * handle-less reference entry points use (tx_compress_dxtn has no init/teardown) */
s2tc_b200_ctx *s2tc_b200_default_ctx(void);

/* ---- whole image, host buffers: the path behind tx_compress_dxtn ------------------------------
 * src: width*height*srccomps bytes, tightly packed, top row first (srccomps 3 or 4; anything else is
 * treated as 4, ref s2tc_algorithm.cpp:1455-1464).  dest/dst_row_stride as for tx_compress_dxtn
 * (ref s2tc_libtxc_dxtn.cpp:243,261,279).  *rand_cursor (may be NULL = 0) is the number of rand() values
 * the reference process would have consumed before this call; it is advanced by
 * blocks * 3|4 * nrandom when nrandom > 0 (ref s2tc_algorithm.cpp:984-992).
 * Pinned (cudaHostAlloc / cudaHostRegister) src and dest are copied without staging. */
int s2tc_b200_compress_host(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int srccomps, int width, int height,
		const uint8_t *src, uint8_t *dest, int dst_row_stride, uint64_t *rand_cursor);

/* ---- block rows of an image already in device memory ------------------------------------------
 * Encodes block rows [row0, row1) of a width x height image.  d_src_rows addresses texel row 4*row0;
 * d_dst receives the blocks of those rows, tightly packed.  rand_cursor0 is the cursor of the IMAGE's
 * first block (the row offset is added internally).  carry (4 ints r,g,b,a; may be NULL = zeros) is the
 * DITHER_SIMPLE error carried into texel row 4*row0 and is updated to the carry leaving the range.
 * `stream` is a cudaStream_t (NULL = the context's own stream); the call is asynchronous unless
 * carry != NULL.  This is the sharding primitive: ranks/GPUs take disjoint row ranges. */
int s2tc_b200_encode_rows_device(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, int *carry, void *stream);

/* DITHER_SIMPLE transfer function of the texel rows of block rows [row0,row1): 4 channels x 32 bytes
 * (byte k of a channel = carry state out for carry state k in; 1-bit alpha: byte 0 = sum mod 255).
 * s2tc_b200_carry_apply(map, channel, srccomps, alphabits, carry_in) evaluates one on the host, so a
 * set of shards can resolve their incoming carries with one 128-byte exchange. */
int s2tc_b200_dither_summary_device(s2tc_b200_ctx *ctx, int srccomps, int alphabits, int width, int height,
		const void *d_src_rows, int row0, int row1, uint64_t maps[16], void *stream);
int s2tc_b200_carry_apply(const uint64_t map[4], int channel, int srccomps, int alphabits, int carry_in);

/* Fully asynchronous forms of the sharding primitives (no host synchronisation; everything stays on `stream`):
 *   _dither_summary_async : the 128-byte summary of this shard's texels is written to device memory (d_maps)
 *   _fold_carry_async     : d_all_maps = the shards' summaries in rank order (e.g. the result of an all-gather);
 *                           writes the carry entering shard `rank` to d_carry (4 ints, device)
 *   _encode_rows_async    : s2tc_b200_encode_rows_device with the carry read from / updated in device memory */
int s2tc_b200_dither_summary_async(s2tc_b200_ctx *ctx, int srccomps, int alphabits, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_maps, void *stream);
int s2tc_b200_fold_carry_async(s2tc_b200_ctx *ctx, const void *d_all_maps, int rank, int srccomps, int alphabits, int *d_carry,
		void *stream);
int s2tc_b200_encode_rows_async(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, int *d_carry, void *stream);
/* _encode_rows_async for the caller that has just called _dither_summary_async for exactly these texels on this context
 * and stream and has not modified them since: the maps the summary left in the workspace are reused (one of the three
 * DITHER_SIMPLE phases is skipped).  Anything else than _fold_carry_async in between, or a different range or stream,
 * and the maps are recomputed -- the result is the same either way as long as the texels are unchanged. */
int s2tc_b200_encode_rows_after_summary_async(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int srccomps, int width, int height,
		const void *d_src_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, int *d_carry, void *stream);

/* ---- one shard of an image from / to HOST memory (several GPUs or processes encode one texture together) ----------
 * Block rows [row0, row1) of a width x height image; src_rows addresses texel row 4*row0, dest receives the shard's
 * blocks (tight rows); rand_cursor0 is the cursor of the IMAGE's first block.  Uploads, kernels and downloads are
 * pipelined in nslab pieces.  DITHER_SIMPLE: d_maps_mine (nslab*128 bytes, device) receives this shard's summaries;
 * gather(user) is called once and must enqueue on `stream` an all-gather of every shard's d_maps_mine into d_maps_all
 * (world*nslab*128 bytes, device, rank order) -- ncclAllGather, torch.distributed.all_gather_into_tensor ...; nslab
 * (1..64) must be the same on every shard.  Other dither modes: gather is not called (FLOYDSTEINBERG shards are a chain:
 * s2tc_b200_floyd_rows_device; here S2TC_B200_EUNSUPPORTED).  Returns when the shard's blocks are in dest.  (the whole-image form is
 * s2tc_b200_compress_host.) */
int s2tc_b200_compress_host_shard(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int srccomps, int width, int height,
		const uint8_t *src_rows, int row0, int row1, uint8_t *dest, uint64_t rand_cursor0, int rank, int nslab, void *d_maps_mine,
		void *d_maps_all, void (*gather)(void *user), void *user, void *stream);

/* ---- DITHER_FLOYDSTEINBERG for row shards (reference rgb565_image, s2tc_algorithm.cpp:1350-1412) ------------------------
 * Error diffusion couples every texel row to the row above, so shards of one image run as a chain: a shard needs the
 * error row its upper neighbour sent below its last row.  One call = one pass over the texel rows of block rows
 * [row0, row1) (d_src_rows: texel row 4*row0; d_reduced_rows: the shard's reduced texels, 4 bytes each, both passes write
 * their bytes into it):
 *   phase 0 (r, g, b): d_err_in = [3][width] ints from the shard above (NULL for the image's first rows);
 *                      d_err_out = [3][width] ints for the shard below -- or, from the image's LAST rows, the seed of the
 *                      alpha pass in its first `width` ints (the reference's alpha pass starts from the red scratch row the
 *                      colour pass left behind, :1380,1397);
 *   phase 1 (alpha; srccomps 4 and alphabits 1 or 4 only): d_err_in = [width] ints: for the image's first rows the seed
 *                      from the last rows' phase 0, otherwise the upper neighbour's d_err_out; d_err_out = [width] ints.
 * Chain for N shards: phase 0 on shard 0, 1, ..., N-1 (each handing d_err_out to the next), then -- if there is an alpha
 * pass -- the seed goes from shard N-1 to shard 0 and phase 1 runs down the shards the same way.  The result is then
 * encoded with s2tc_b200_encode_reduced_rows_device.  A chain has no parallelism between shards (the recurrence has none):
 * sharding Floyd-Steinberg distributes memory and the block encoder's work, not the pre-pass. */
int s2tc_b200_floyd_rows_device(s2tc_b200_ctx *ctx, int srccomps, int alphabits, int width, int height, const void *d_src_rows,
		int row0, int row1, int phase, const int *d_err_in, int *d_err_out, void *d_reduced_rows, void *stream);
/* block rows [row0, row1) from texels the caller has already reduced (4 bytes each {r5, g6, b5, a}; s->dither is ignored) */
int s2tc_b200_encode_reduced_rows_device(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int width, int height,
		const void *d_reduced_rows, int row0, int row1, void *d_dst, uint64_t rand_cursor0, void *stream);

/* ---- the same, STRIPED: several GPUs encode one texture and overlap their uploads with their kernels ----------------
 * The block rows are cut into nwave waves -- in proportion to wave_weights[0..nwave) (NULL: equal; small first and last
 * waves shorten the head and the tail of the pipeline) -- and every wave evenly into `world` stripes of consecutive block
 * rows; stripe wave*world + rank belongs to shard `rank` (s2tc_b200_stripe_rows gives its block rows [row0, row1); empty
 * when there are more stripes than block rows).
 * src_stripes[w] / dest_stripes[w]: host memory of this shard's stripe of wave w (texel row 4*row0; tight block rows).
 * DITHER_SIMPLE with world > 1: after wave w is summarised, gather(user, w) is called and must enqueue on `stream` an
 * all-gather of 128 bytes at d_maps_mine + 128*w of every shard into d_maps_all + 128*world*w (rank order); both are
 * device buffers (nwave*128 and nwave*world*128 bytes).  A contiguous shard (s2tc_b200_compress_host_shard) cannot encode
 * anything before every shard above it has been uploaded; striped, wave w is encoded while wave w+1 is uploaded.
 * Each stripe is a contiguous run of the reference's block-row loop (s2tc_libtxc_dxtn.cpp:246-258); carry and rand()
 * cursor cross stripe boundaries exactly as they cross slab boundaries in the whole-image call. */
void s2tc_b200_stripe_rows(int height, int world, int nwave, const int *wave_weights, int wave, int rank, int *row0, int *row1);
int s2tc_b200_compress_host_striped(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int srccomps, int width, int height,
		const uint8_t *const *src_stripes, uint8_t *const *dest_stripes, uint64_t rand_cursor0, int rank, int world, int nwave,
		const int *wave_weights, void *d_maps_mine, void *d_maps_all, void (*gather)(void *user, int wave), void *user, void *stream);

/* ---- whole mip chain of an RGBA8 image on the device (SURVEY "next" N2) -----------------------------------
 * What the reference tool does per file after the DDS header (s2tc_compress.c:722-733): encode the level with
 * one tx_compress_dxtn call, halve it with Image_MipReduce32 (:427-493), repeat down to 1x1.  Every level is its own
 * "call" (the DITHER_SIMPLE carry restarts, the rand cursor continues).  Levels are written back to back, tightly.
 * _device: d_rgba is overwritten (ping-pong with d_scratch, >= width*height bytes); d_dst holds
 * s2tc_b200_mipchain_bytes(dxt, width, height) bytes. */
size_t s2tc_b200_mipchain_bytes(int dxt, int width, int height);
int s2tc_b200_mip_reduce_device(s2tc_b200_ctx *ctx, const void *d_in, int width, int height, void *d_out, void *stream);
int s2tc_b200_compress_mipchain_device(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int width, int height, void *d_rgba,
		void *d_scratch, void *d_dst, uint64_t *rand_cursor, void *stream);
int s2tc_b200_compress_mipchain_host(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, int width, int height, const uint8_t *rgba,
		uint8_t *dest, uint64_t *rand_cursor);
/* A batch of ntex equally sized RGBA8 textures (back to back in d_rgba, not modified) under nset settings: every mip level
 * of all textures goes through each kernel in ONE launch (the per-level kernels of a single chain are tiny from 64x64
 * down), and the 565 pre-pass of a level is shared by the settings that agree on dither mode and alpha width.  The chain
 * of texture i under setting k is written to d_dst + off_k + i*mipchain_bytes(dxt_k), where off_0 = 0 and
 * off_{k+1} = off_k + ntex*mipchain_bytes(dxt_k) rounded up to a multiple of 16 (d_dst itself 16-byte aligned).
 * Each texture is its own run of the reference tool: carries restart per level, every texture's rand() cursor starts
 * at rand_cursor0.  d_scratch: ntex*(width*height + width*height/4) + 256 bytes.  (BASELINE config 4; the reference loop
 * is s2tc_compress.c:722-733 once per file and per S2TC_COLORDIST_MODE.) */
int s2tc_b200_compress_mipchain_batch_device(s2tc_b200_ctx *ctx, const s2tc_b200_settings *sets, int nset, int width, int height,
		int ntex, const void *d_rgba, void *d_scratch, void *d_dst, uint64_t rand_cursor0, void *stream);

/* ---- 565 pre-pass only: backs the exported rgb565_image (ref s2tc_algorithm.h:38) ------------- */
int s2tc_b200_rgb565_host(s2tc_b200_ctx *ctx, uint8_t *out, const uint8_t *src, int width, int height, int srccomps,
		int alphabits, int dither);

/* ---- one block: backs the function pointers s2tc_encode_block_func returns (ref s2tc_algorithm.h:65-66)
 * rgba addresses the block's first texel inside a pre-reduced 4-byte/texel image of row stride iw. */
int s2tc_b200_encode_block_host(s2tc_b200_ctx *ctx, const s2tc_b200_settings *s, uint8_t *out, const uint8_t *rgba, int iw,
		int w, int h, uint64_t *rand_cursor);

/* ---- S3TC -> S2TC transcode of nblocks blocks, in place (ref s2tc_from_s3tc.cpp:254-263) ------ */
int s2tc_b200_transcode_host(s2tc_b200_ctx *ctx, int dxt, uint8_t *blocks, size_t nblocks);
int s2tc_b200_transcode_device(s2tc_b200_ctx *ctx, int dxt, void *d_blocks, size_t nblocks, void *stream);

/* ---- S2TC decode of whole images (SURVEY "next" N4): what calling fetch_2d_texel_rgba_dxt1/3/5
 * (ref s2tc_libtxc_dxtn.cpp:57-140) for every texel gives; blocks tightly packed, RGBA8 out ------------------- */
int s2tc_b200_decode_device(s2tc_b200_ctx *ctx, int dxt, const void *d_blocks, int width, int height, void *d_rgba, void *stream);
int s2tc_b200_decode_host(s2tc_b200_ctx *ctx, int dxt, const uint8_t *blocks, int width, int height, uint8_t *rgba);

/* Settings as tx_compress_dxtn would read them from the S2TC_* environment right now (warnings on stderr for
 * bad values, as in ref s2tc_libtxc_dxtn.cpp:160-216), for callers of the explicit-settings functions. */
void s2tc_b200_settings_from_env(int dxt, s2tc_b200_settings *out);

/* ---- process-wide rand() cursor used by the handle-less entry points -------------------------- */
uint64_t s2tc_b200_rand_cursor_get(void);
void s2tc_b200_rand_cursor_set(uint64_t draws);

/* ---- measurement helpers (bench.py): device-side timing on the context's stream ---------------- */
int s2tc_b200_sync(s2tc_b200_ctx *ctx);
/* number of kernel launches this context has issued so far */
uint64_t s2tc_b200_launch_count(s2tc_b200_ctx *ctx);
/* device milliseconds spent in each kernel family since the last reset, measured with CUDA events when
 * profiling is enabled (s2tc_b200_profile_enable(ctx, 1)): [0] pre-pass, [1] random candidates,
 * [2] pair search, [3] finish, [4] fast encode, [5] transcode; counts in the second array */
int s2tc_b200_profile_enable(s2tc_b200_ctx *ctx, int on);
int s2tc_b200_profile_read(s2tc_b200_ctx *ctx, double ms[6], uint64_t launches[6], int reset);

/* sustained integer (min + add) rates of the device in Gop/s, the roofline denominators of the search kernels:
 * scalar 32-bit operands (best of the compiler's own mix and VIMNMX + IMAD on two pipes), and 16-bit operands packed
 * two to a register (VIMNMX.U16x2 + IDP.2A).  s2tc_b200_int32_peak returns the scalar one. */
int s2tc_b200_int_peaks(s2tc_b200_ctx *ctx, double *scalar_gops, double *packed16_gops);
int s2tc_b200_int32_peak(s2tc_b200_ctx *ctx, double *gops);

#ifdef __cplusplus
}
#endif
#endif
