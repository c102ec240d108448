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
