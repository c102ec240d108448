"""GPU tier at BASELINE.json's full sizes.  The oracle cannot encode 4-17 M blocks in test time, so these check
(1) spot rows against the oracle -- first, middle and last block rows, i.e. with the DITHER_SIMPLE carry and the rand()
cursor (up to 3.2 G draws) having run through the whole image -- and (2) size-independent properties: a 4-way
row-sharded encode equals the whole-image encode, the host path equals the device path, decode error stays bounded."""
import hashlib

import numpy as np
import pytest
import torch

import _oracle as O
from s2tc_b200 import Settings, synth
from s2tc_b200.sharding import fold_carry, shard_block_rows

pytestmark = pytest.mark.gpu


def _spot_rows(bh):
    return [(0, 4), (bh // 2 - 2, bh // 2 + 2), (bh - 4, bh)]


def test_config2_full_size(encoder):
    """DXT5 8192x8192 SRGB_MIXED nrandom=0 LOOP, DITHER_SIMPLE (bench.py's default workload)."""
    w = h = 8192
    img = synth.synth_rgba(w, h, seed=1234)
    st = Settings(O.DXT5, O.SRGB_MIXED, 0, O.LOOP, O.DITHER_SIMPLE)
    got = encoder.compress(img, st)
    bw, bh = w // 4, h // 4
    for r0, r1 in _spot_rows(bh):
        want = O.orc_rows(img, O.DXT5, O.SRGB_MIXED, 0, O.LOOP, O.DITHER_SIMPLE, (r0, r1))
        assert np.array_equal(got[r0 * bw * 16:r1 * bw * 16], want), (r0, r1)
    # sharded == whole
    d_img = torch.from_numpy(img).cuda()
    stream = torch.cuda.Stream()
    parts = []
    with torch.cuda.stream(stream):
        ranges = [shard_block_rows(bh, 4, r) for r in range(4)]
        shards = [d_img[4 * a:4 * b] for a, b in ranges]
        sums = [encoder.dither_summary_device(s, w, h, 4, 8, a, b, stream=stream.cuda_stream) for s, (a, b) in zip(shards, ranges)]
        for r, (a, b) in enumerate(ranges):
            out = torch.empty((b - a) * bw * 16, dtype=torch.uint8, device="cuda")
            encoder.encode_rows_device(shards[r], w, h, 4, a, b, out, st, carry=fold_carry(sums, r, 4, 8), stream=stream.cuda_stream)
            parts.append(out)
        stream.synchronize()
    sharded = np.concatenate([p.cpu().numpy() for p in parts])
    assert hashlib.sha256(sharded.tobytes()).digest() == hashlib.sha256(got.tobytes()).digest()
    # decode round trip: bounded error on the opaque part
    dec = encoder.decode(got, O.DXT5, w, h)
    opaque = img[..., 3] == 255
    err = np.abs(dec[..., :3].astype(np.int16) - img[..., :3].astype(np.int16))[opaque].mean()
    assert err < 8, err


def test_config3_full_size(encoder):
    """DXT1 16384x16384 WAVG nrandom=64 LOOP: the rand() cursor reaches 3.2e9 draws at the last block."""
    w = h = 16384
    img = synth.synth_rgba(w, h, seed=1234)
    st = Settings(O.DXT1, O.WAVG, 64, O.LOOP, O.DITHER_NONE)
    got, cur = encoder.compress(img, st, cursor=5, return_cursor=True)
    bw, bh = w // 4, h // 4
    assert cur == 5 + bw * bh * 192
    for r0, r1 in [(0, 2), (bh // 2, bh // 2 + 2), (bh - 2, bh)]:
        want = O.orc_rows(img, O.DXT1, O.WAVG, 64, O.LOOP, O.DITHER_NONE, (r0, r1), cursor=5)
        assert np.array_equal(got[r0 * bw * 8:r1 * bw * 8], want), (r0, r1)
    # the bench configuration itself (DITHER_SIMPLE): the host call from pageable memory (staging ring, growing slabs, two
    # lanes with the carry chained between them) against the oracle, and the device-resident call (slabs alternating between
    # the lanes inside one range) against the host call
    st = Settings(O.DXT1, O.WAVG, 64, O.LOOP, O.DITHER_SIMPLE)
    got = encoder.compress(img, st, cursor=0)
    for r0, r1 in [(0, 2), (bh // 4 + 1, bh // 4 + 2), (bh - 1, bh)]:
        want = O.orc_rows(img, O.DXT1, O.WAVG, 64, O.LOOP, O.DITHER_SIMPLE, (r0, r1), cursor=0)
        assert np.array_equal(got[r0 * bw * 8:r1 * bw * 8], want), (r0, r1)
    d_img = torch.from_numpy(img).cuda()
    d_out = torch.empty(bw * bh * 8, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    encoder.encode_rows_device(d_img, w, h, 4, 0, bh, d_out, st, cursor0=0, carry=None, stream=stream.cuda_stream)
    stream.synchronize()
    torch.cuda.synchronize()
    assert hashlib.sha256(d_out.cpu().numpy().tobytes()).digest() == hashlib.sha256(got.tobytes()).digest()


def test_config5_full_size(encoder):
    """DXT5 4096x4096 NORMALMAP (nrandom -1 = the 16-candidate search for this metric), REFINE_NEVER, DITHER_SIMPLE on a
    normal map: spot rows against the oracle, 3-way row shards equal the whole image, unit-length decode."""
    w = h = 4096
    img = synth.synth_normal(w, h, seed=7)
    st = Settings(O.DXT5, O.NORMALMAP, -1, O.NEVER, O.DITHER_SIMPLE)
    got = encoder.compress(img, st)
    bw, bh = w // 4, h // 4
    for r0, r1 in _spot_rows(bh):
        want = O.orc_rows(img, O.DXT5, O.NORMALMAP, -1, O.NEVER, O.DITHER_SIMPLE, (r0, r1))
        assert np.array_equal(got[r0 * bw * 16:r1 * bw * 16], want), (r0, r1)
    d_img = torch.from_numpy(img).cuda()
    stream = torch.cuda.Stream()
    parts = []
    with torch.cuda.stream(stream):
        ranges = [shard_block_rows(bh, 3, r) for r in range(3)]
        shards = [d_img[4 * a:4 * b] for a, b in ranges]
        sums = [encoder.dither_summary_device(s, w, h, 4, 8, a, b, stream=stream.cuda_stream) for s, (a, b) in zip(shards, ranges)]
        for r, (a, b) in enumerate(ranges):
            out = torch.empty((b - a) * bw * 16, dtype=torch.uint8, device="cuda")
            encoder.encode_rows_device(shards[r], w, h, 4, a, b, out, st, carry=fold_carry(sums, r, 4, 8), stream=stream.cuda_stream)
            parts.append(out)
        stream.synchronize()
    sharded = np.concatenate([p.cpu().numpy() for p in parts])
    assert hashlib.sha256(sharded.tobytes()).digest() == hashlib.sha256(got.tobytes()).digest()
    dec = encoder.decode(got, O.DXT5, w, h)
    err = np.abs(dec[..., :3].astype(np.int16) - img[..., :3].astype(np.int16)).mean()
    assert err < 8, err


def test_defaults_full_size_and_floyd(encoder):
    """Reference defaults on 8192x8192 (fast mode, DITHER_SIMPLE) and the same with FLOYDSTEINBERG on 2048x2048."""
    w = h = 8192
    img = synth.synth_rgba(w, h, seed=7)
    got = encoder.compress(img, Settings())
    bw, bh = w // 4, h // 4
    for r0, r1 in _spot_rows(bh):
        want = O.orc_rows(img, O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE, (r0, r1))
        assert np.array_equal(got[r0 * bw * 8:r1 * bw * 8], want), (r0, r1)
    small = np.ascontiguousarray(img[:2048, :2048])
    for dxt in (O.DXT1, O.DXT3):
        a = encoder.compress(small, Settings(dxt, O.WAVG, -1, O.ALWAYS, O.DITHER_FS))
        b = O.orc_compress(small, dxt, O.WAVG, -1, O.ALWAYS, O.DITHER_FS)
        assert np.array_equal(a, b), dxt
