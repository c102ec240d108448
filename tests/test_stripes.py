"""CPU tier: the stripe geometry of the multi-GPU host path (s2tc_b200_stripe_rows, include/s2tc_b200.h).  Pure host
arithmetic in the shared object, no GPU needed.  Property: for any image height, shard count, wave count and wave weights the
stripes, taken in (wave, rank) order, tile the block rows [0, ceil(height / 4)) exactly -- contiguous, disjoint, complete --
which is what lets the DITHER_SIMPLE carry and the rand() cursor run through them as through one whole-image call
(reference loop: s2tc_libtxc_dxtn.cpp:246-258)."""
import random

from s2tc_b200 import Encoder


def _stripes(height, world, nwave, weights):
    return [Encoder.stripe_rows(height, world, nwave, w, r, weights) for w in range(nwave) for r in range(world)]


def test_stripes_tile_the_block_rows():
    rng = random.Random(7)
    cases = [(16384, 8, 6, [1, 2, 4, 4, 3, 2]), (16384, 2, 6, [1, 2, 4, 4, 3, 2]), (58, 2, 8, None), (1, 3, 4, None),
             (7, 8, 6, [1, 2, 4, 4, 3, 2]), (4096, 1, 1, None), (150, 3, 4, [0, 1, 0, 5])]
    for _ in range(300):
        nwave = rng.randint(1, 12)
        weights = None if rng.random() < 0.3 else [rng.randint(0, 9) for _ in range(nwave)]
        if weights is not None and sum(weights) == 0:
            weights[rng.randrange(nwave)] = 1
        cases.append((rng.randint(1, 70000), rng.randint(1, 9), nwave, weights))
    for height, world, nwave, weights in cases:
        bh = (height + 3) // 4
        pos = 0
        for a, b in _stripes(height, world, nwave, weights):
            assert a == pos and b >= a, (height, world, nwave, weights, a, b, pos)
            pos = b
        assert pos == bh, (height, world, nwave, weights)


def test_equal_waves_split_evenly():
    # 4096 block rows, 8 shards, 8 equal waves: every stripe has 64 block rows
    assert {b - a for a, b in _stripes(16384, 8, 8, None)} == {64}
    # weights scale the waves, not the split inside a wave
    rows = _stripes(16384, 2, 6, [1, 2, 4, 4, 3, 2])
    assert [b - a for a, b in rows] == [128, 128, 256, 256, 512, 512, 512, 512, 384, 384, 256, 256]


def test_out_of_range_wave_is_empty():
    assert Encoder.stripe_rows(100, 2, 3, 5, 0) == (0, 0)
