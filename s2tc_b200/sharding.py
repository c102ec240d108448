"""Host-side sharding arithmetic for encoding one texture (or a batch) on several GPUs.

Blocks are independent except for two pieces of sequential state of the reference:
  * the rand() cursor: block b of an image starts at cursor0 + b * draws_per_block -- closed form,
    handled inside s2tc_b200_encode_rows_device from the image-level cursor;
  * the DITHER_SIMPLE carry, which runs through the image in raster order: every shard computes the
    transfer function of its own texels (s2tc_b200_dither_summary_device), the shards all-gather these
    128-byte summaries, and each folds the summaries of the shards before it.
No other communication exists: each GPU writes its own slice of the output.
"""
from .api import Encoder


def shard_block_rows(total_block_rows, world, rank):
    """Contiguous block-row range [row0, row1) of `rank` (the reference walks block rows top to bottom)."""
    return (total_block_rows * rank) // world, (total_block_rows * (rank + 1)) // world


def fold_carry(summaries, rank, comps, alphabits, carry=(0, 0, 0, 0)):
    """Carry entering shard `rank`, given the transfer-function summaries (16 words each) of all shards."""
    carry = list(carry)
    for r in range(rank):
        carry = Encoder.carry_apply(summaries[r], comps, alphabits, carry)
    return carry


def summary_to_i64(words):
    """uint64 words -> the int64 values a torch / gloo / nccl tensor can carry."""
    return [w - (1 << 64) if w >= (1 << 63) else w for w in words]


def summary_from_i64(vals):
    return [int(v) & ((1 << 64) - 1) for v in vals]


def floyd_steinberg_sharded(enc, dist, src_rows, width, height, comps, alphabits, row0, row1, rank, world, reduced_rows,
                            new_ints, stream=None):
    """DITHER_FLOYDSTEINBERG pre-pass of one image whose block rows [row0, row1) live on this rank (contiguous, non-empty
    shards in rank order).  Error diffusion is a recurrence over rows, so the shards run as a chain
    (s2tc_b200_floyd_rows_device): the colour pass goes down the ranks, each handing the error row below its last texel
    row to the next rank; the reference's alpha pass is seeded with the red leftovers of the image's LAST row
    (s2tc_algorithm.cpp:1380,1397), so the seed travels from the last rank to rank 0 and the alpha pass goes down the same
    way.  dist: torch.distributed (send / recv on the current stream) or anything with the same two calls;
    new_ints(n): a zeroed int32 buffer of n elements on the device the exchange uses.  Writes reduced_rows (4 bytes per
    texel), to be encoded with Encoder.encode_reduced_rows_device.
    Everything is asynchronous on `stream`, which must be the stream `dist` and `new_ints` work on (torch's current stream);
    the exchange buffers are returned and must be kept alive until the stream has finished with them -- if they are
    allocated on another stream, a caching allocator may hand their memory out again while the kernels still use it."""
    if row1 <= row0:
        raise ValueError("Floyd-Steinberg shards must not be empty")
    err_in, err_out = new_ints(3 * width), new_ints(3 * width)
    if rank > 0:
        dist.recv(err_in, src=rank - 1)
    enc.floyd_rows_device(src_rows, width, height, comps, alphabits, row0, row1, 0, err_in if rank > 0 else None, err_out,
                          reduced_rows, stream=stream)
    if rank < world - 1:
        dist.send(err_out, dst=rank + 1)
    if comps != 4 or alphabits == 8:
        return err_in, err_out
    seed, a_out = new_ints(width), new_ints(width)
    if world == 1:
        seed = err_out[:width]
    elif rank == world - 1:
        dist.send(err_out[:width].contiguous(), dst=0)
    if rank == 0 and world > 1:
        dist.recv(seed, src=world - 1)
    if rank > 0:
        dist.recv(seed, src=rank - 1)
    enc.floyd_rows_device(src_rows, width, height, comps, alphabits, row0, row1, 1, seed, a_out, reduced_rows, stream=stream)
    if rank < world - 1:
        dist.send(a_out, dst=rank + 1)
    return err_in, err_out, seed, a_out
