"""CPU tier: the encoder's own host+device source (s2tc_b200/csrc/*.cuh) compiled for the CPU and checked
against the oracle.  This covers, without a GPU, the per-block logic (fast pick, gather, scalar search,
refinement, packing), the eight metrics, the rand() jump-ahead, the dither transfer maps (byte-permute
composition included) and the transcoder.  The cooperative search kernels and the memory paths are
covered by the -m gpu tests."""
import itertools

import numpy as np

import _hostsim as H
import _oracle as O
from s2tc_b200 import synth


def test_reciprocal_division_exhaustive():
    assert H.lib().hostsim_check_division() == 0


def test_metrics_against_oracle():
    rng = np.random.default_rng(0)
    pairs = rng.integers(0, 256, size=(20000, 6))
    pairs[:, [0, 2, 3, 5]] >>= 3
    pairs[:, [1, 4]] >>= 2
    # saturated magenta-vs-green pairs make the SRGB metric wrap to a NEGATIVE distance (1.3e-6 of random pairs)
    sat = np.array([[31, 63, 31, 0, 0, 0], [0, 0, 0, 31, 63, 31], [31, 0, 31, 0, 63, 0], [31, 63, 0, 0, 0, 31],
                    [31, 14, 31, 0, 61, 5], [1, 58, 1, 31, 2, 31], [3, 60, 1, 31, 2, 31], [31, 23, 31, 3, 63, 2]])
    for cd in range(8):
        neg = 0
        for a in np.vstack([pairs[:3000], sat]):
            ca = bytes([a[0], a[1], a[2]]); cb = bytes([a[3], a[4], a[5]])
            want = O.lib().orc_color_dist(cd, ca, cb)
            pa = int(a[0]) | int(a[1]) << 8 | int(a[2]) << 16
            pb = int(a[3]) | int(a[4]) << 8 | int(a[5]) << 16
            assert H.lib().hostsim_color_dist(cd, pa, pb) == want, (cd, a)
            neg += want < 0
        if cd == O.SRGB:
            assert neg > 0   # the int32 wrap of the SRGB metric is exercised


def magenta_green(width, height, seed):
    """Texels drawn from saturated magentas and greens: SRGB distances between them wrap negative, so the
    pair search has to follow the reference's "bestsum < 0 || sum < bestsum" rule."""
    rng = np.random.default_rng(seed)
    pal = np.array([[255, 16, 255, 255], [8, 250, 8, 255], [24, 255, 40, 0], [255, 100, 255, 128], [255, 8, 248, 255],
                    [0, 244, 0, 255], [16, 232, 8, 255], [248, 60, 255, 64]], np.uint8)
    return pal[rng.integers(0, len(pal), size=(height, width))]


def test_srgb_negative_sums_follow_the_reference_rule():
    img = magenta_green(32, 32, 1)
    for dxt, nr, rf in itertools.product((0, 1, 2), (0, 6), (0, 1, 2)):
        a = H.compress(img, dxt, O.SRGB, nr, rf, 0, cursor=2)
        b = O.orc_compress(img, dxt, O.SRGB, nr, rf, 0, cursor=2)
        assert np.array_equal(a, b), (dxt, nr, rf)
        if O.ref_available():
            assert np.array_equal(b, O.ref_compress(img, dxt, O.SRGB, nr, rf, 0, cursor=2)), (dxt, nr, rf)


def test_rand_jump_ahead():
    seq = O.orc_rand(5000)
    assert H.rand(0, 192, 0, 50) == seq[:50]
    assert H.rand(7, 192, 3, 40) == seq[7 + 3 * 192: 7 + 3 * 192 + 40]
    assert H.rand(1000, 96, 25, 10) == seq[1000 + 25 * 96: 1000 + 25 * 96 + 10]
    far = 3 * 10 ** 11
    assert H.rand(far, 256, 1000003, 6) == O.orc_rand(6, far + 256 * 1000003)


def test_whole_pipeline_against_oracle():
    imgs = [synth.synth_rgba(40, 28, seed=3), synth.synth_noise(37, 21, seed=4), synth.synth_noise(24, 20, seed=5, comps=3)]
    for img in imgs:
        for dxt, cd, nr, rf, di in itertools.product((0, 1, 2), range(8), (-1, 0, 5, 37), (0, 1, 2), (0, 1, 2)):
            if (cd + nr + rf + di + dxt) % 3:   # a third of the grid keeps the CPU tier quick
                continue
            a = H.compress(img, dxt, cd, nr, rf, di, cursor=13)
            b = O.orc_compress(img, dxt, cd, nr, rf, di, cursor=13)
            assert np.array_equal(a, b), (img.shape, dxt, cd, nr, rf, di)


def test_dither_simple_large_and_ragged():
    """Chunk/tile boundaries of the carry scan: > 1 tile (16384 texels), ragged last chunk, saturated values."""
    for img in (synth.synth_noise(300, 131, seed=2), synth.synth_rgba(257, 129, seed=3),
                np.full((70, 250, 4), 255, np.uint8), synth.synth_noise(190, 90, seed=6, comps=3),
                synth.synth_noise(1000, 600, seed=7)):   # 37 tiles: two parts of the scan
        h, w, c = img.shape
        for ab in (1, 4, 8):
            got = np.zeros((h, w, 4), np.uint8)
            H.lib().hostsim_prepass(c, ab, 1, h * w, img.ctypes.data_as(H._u8p), got.ctypes.data_as(H._u8p))
            assert np.array_equal(got, O.orc_prepass(img, ab, 1)), (img.shape, ab)


def test_dither_carry_sharding():
    """A carry chain cut into shards: summaries fold to the right incoming carry (what multi-GPU runs exchange)."""
    from s2tc_b200.sharding import fold_carry, shard_block_rows
    img = synth.synth_noise(64, 96, seed=9)
    want = O.orc_prepass(img, 4, 1)
    world = 3
    bh = 96 // 4
    ranges = [shard_block_rows(bh, world, r) for r in range(world)]
    summaries = [H.dither_summary(img[4 * a:4 * b], 4, 4) for a, b in ranges]
    for r, (a, b) in enumerate(ranges):
        carry = fold_carry_host(summaries, r)
        got, _ = H.prepass_range(img[4 * a:4 * b], 4, 4, carry)
        assert np.array_equal(got.reshape(-1, 64, 4), want[4 * a:4 * b]), r


def fold_carry_host(summaries, rank):
    """fold_carry() evaluates maps through the CUDA library's pure-host helper; here the same fold is done
    with numpy so that the test also runs where the library is not built."""
    carry = [0, 0, 0, 0]
    radius = [7, 3, 7, 15]
    for r in range(rank):
        raw = np.array(summaries[r], np.uint64).view(np.uint8).reshape(4, 32)
        carry = [int(raw[ch][carry[ch] + radius[ch]]) - radius[ch] for ch in range(4)]
    return carry


def test_transcode_against_oracle():
    for dxt in (0, 1, 2):
        blocks = synth.synth_s3tc_blocks(2000, dxt, seed=8)
        assert np.array_equal(H.transcode(blocks, dxt), O.orc_transcode(blocks, dxt))
