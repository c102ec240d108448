"""GPU tier, the drop-in boundary through the REAL exported symbols (ctypes on the shared object, nothing of the Python
mirror in between): the function pointers `s2tc_encode_block_func` / `get_s2tc_encoder` hand out
(reference s2tc_algorithm.cpp:1110-1194, s2tc_algorithm.h:65-66), the exported `rgb565_image` (:1416-1465), and
`tx_compress_dxtn` itself with the S2TC_* environment changed between calls, invalid values, a bad destformat,
dstRowStride cases and pageable caller memory (reference s2tc_libtxc_dxtn.cpp:142-299)."""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

import _oracle as O
import s2tc_b200
from s2tc_b200 import synth

pytestmark = pytest.mark.gpu

# what the block function sees -- 5/6/5 colours, row-major 4x4, alpha still 8 bits
K1 = bytes.fromhex("1c1e1a160438142705331a1c0a2d01961f10181c0d3507621d2819cf1a3a15ad1a060477053f15330c0d0dc41b3012af1e3004b21f120a3613280c6910330f99")
K2 = bytes.fromhex("040a14ff070a13ff0a0a12ff0d0a11ff041113ff071112000a1111ff0d1110ff041812ff071811ff0a1810000d180fff041f11ff071f10ff0a1f0fff0d1f0eff")
K3 = bytes.fromhex("1f3f1f01" * 16)
B4 = {  # (dxt, cd, nrandom, refine) -> outputs for K1, K2, K3
    (O.DXT1, O.WAVG, -1, O.ALWAYS): ("ad610ebe7f5f4f7d", "9043b249555d3000", "0000ffffffffffff"),
    (O.DXT1, O.WAVG, 0, O.LOOP): ("ad610ebe7f5f4f7d", "9043b249555d3000", "00000100ffffffff"),
    (O.DXT1, O.RGB, 0, O.NEVER): ("ad6112de7f5f4f7d", "f03b5251555d3000", "00000100ffffffff"),
    (O.DXT1, O.AVG, 0, O.ALWAYS): ("a75b31d63f5f4f7d", "70439251555c3000", "00000100ffffffff"),
    (O.DXT3, O.SRGB, -1, O.ALWAYS): ("219161ac37ac3b960db3918e14544441", "ffff0ffffff0ffffb053124255550500", "0000000000000000fffffeff00000000"),
    (O.DXT3, O.W0AVG, 0, O.LOOP): ("219161ac37ac3b962fd24f8e54544451", "ffff0ffffff0ffff904bd24955550000", "0000000000000000fffffeff00000000"),
    (O.DXT5, O.SRGB_MIXED, 0, O.LOOP): ("55b0b663244012202fd24f8e54544451", "feffff7fffbfffff904bd24955550000", "0102000000000000fffffeff00000000"),
    (O.DXT5, O.NORMALMAP, -1, O.NEVER): ("62afb6632440122018fa6f8654444451", "feffff7fffbfffffef53315255555000", "0102000000000000fffffeff00000000"),
    (O.DXT5, O.YUV, -1, O.LOOP): ("55b0b663244012208ab2329615544451", "00ffff7fffbfffff7053d23955150100", "0102000000000000fffffeff00000000"),
}


def _reduce_alpha(block, dxt):
    """what rgb565_image DITHER_NONE does to alpha before the block function runs"""
    a = np.frombuffer(block, np.uint8).reshape(4, 4, 4).copy()
    if dxt == O.DXT1:
        a[..., 3] >>= 7
    elif dxt == O.DXT3:
        a[..., 3] >>= 4
    return a


def _call(fn, px, dxt, w=4, h=4, nrandom=0, iw=4):
    out = np.zeros(O.block_bytes(dxt), np.uint8)
    fn(out.ctypes.data, px.ctypes.data, iw, w, h, nrandom)
    return out


@pytest.fixture(scope="module")
def L():
    lib = s2tc_b200.lib()
    assert lib.s2tc_b200_device_count() > 0
    return lib


def test_block_function_pointers_all_reachable_combinations(L):
    """3 formats x 8 metrics x {fast, normal} x 3 refinements minus NORMALMAP-fast = 135 instantiations
    (ref :1110-1194), each through the pointer of BOTH factory names, on K1-K3, against the oracle and SURVEY B.4."""
    combos = 0
    for dxt, cd, nr, rf in itertools.product((O.DXT1, O.DXT3, O.DXT5), range(8), (-1, 0), (O.NEVER, O.ALWAYS, O.LOOP)):
        if cd == O.NORMALMAP and nr < 0:
            continue   # the factory hands out the normal-mode function for NORMALMAP whatever nrandom is (ref :1139)
        combos += 1
        f1 = L.s2tc_encode_block_func(dxt, cd, nr, rf)
        f2 = L.get_s2tc_encoder(dxt, cd, nr, rf)
        for k, raw in enumerate((K1, K2, K3)):
            px = _reduce_alpha(raw, dxt)
            want = O.orc_encode_block(px, 4, 4, dxt, cd, nr, rf)
            for f in (f1, f2):
                got = _call(f, px, dxt, nrandom=nr)
                assert np.array_equal(got, want), (dxt, cd, nr, rf, k, got.tobytes().hex(), want.tobytes().hex())
            if (dxt, cd, nr, rf) in B4:
                assert want.tobytes().hex() == B4[(dxt, cd, nr, rf)][k]
    assert combos == 135
    # NORMALMAP asked for with nrandom = -1 behaves exactly like nrandom = 0
    px = _reduce_alpha(K1, O.DXT5)
    a = _call(L.s2tc_encode_block_func(O.DXT5, O.NORMALMAP, -1, O.NEVER), px, O.DXT5, nrandom=-1)
    b = _call(L.s2tc_encode_block_func(O.DXT5, O.NORMALMAP, 0, O.NEVER), px, O.DXT5, nrandom=0)
    assert np.array_equal(a, b) and a.tobytes().hex() == B4[(O.DXT5, O.NORMALMAP, -1, O.NEVER)][0]
    # out-of-range enumerators fall back like the reference's switch defaults: refine -> ALWAYS, dxt -> DXT5, cd -> WAVG
    px = _reduce_alpha(K2, O.DXT5)
    assert np.array_equal(_call(L.s2tc_encode_block_func(7, 99, 0, 5), px, O.DXT5),
                          O.orc_encode_block(px, 4, 4, O.DXT5, O.WAVG, 0, O.ALWAYS))


def test_block_function_pointer_rand_sequence(L):
    """SURVEY B.4, the rand()-dependent sequence of a fresh process: the process-wide cursor advances by 4 (DXT5) or 3
    draws per candidate and call, and a partial block (w=3, h=2) keeps index 0 for its missing texels."""
    L.s2tc_b200_rand_cursor_set(0)
    f5 = L.s2tc_encode_block_func(O.DXT5, O.WAVG, 4, O.ALWAYS)
    want5 = ("55b0b663244012202fd24f8e54544451", "feffff7fffbfffff304b934155050000", "0102000000000000fffffeff00000000")
    for raw, want in zip((K1, K2, K3), want5):
        assert _call(f5, _reduce_alpha(raw, O.DXT5), O.DXT5, nrandom=4).tobytes().hex() == want
    assert L.s2tc_b200_rand_cursor_get() == 48
    f1 = L.get_s2tc_encoder(O.DXT1, O.WAVG, 4, O.LOOP)
    assert _call(f1, _reduce_alpha(K1, O.DXT1), O.DXT1, w=3, h=2, nrandom=4).tobytes().hex() == "19ed1aed3f0f0000"
    assert L.s2tc_b200_rand_cursor_get() == 60
    # a block inside a larger pre-reduced image: iw is the row stride in texels
    img = O.orc_prepass(synth.synth_rgba(16, 8, seed=2), 8, O.DITHER_NONE)
    L.s2tc_b200_rand_cursor_set(7)
    sub = img[4:, 8:]   # block (2, 1): a view; the pointer passed is that of its first texel
    out = np.zeros(16, np.uint8)
    f5(out.ctypes.data, img.ctypes.data + (4 * 16 + 8) * 4, 16, 4, 4, 4)
    assert np.array_equal(out, O.orc_encode_block(np.ascontiguousarray(sub[:, :4]), 4, 4, O.DXT5, O.WAVG, 4, O.ALWAYS, cursor=7))


def test_exported_rgb565_image(L):
    rng = np.random.default_rng(11)
    for (w, h), comps, abits, dither in itertools.product(((37, 21), (64, 64), (1, 1)), (3, 4), (1, 4, 8), (0, 1, 2)):
        src = rng.integers(0, 256, size=(h, w, comps), dtype=np.uint8)
        out = np.zeros((h, w, 4), np.uint8)
        L.rgb565_image(out.ctypes.data, src.ctypes.data, w, h, comps, abits, dither)
        assert np.array_equal(out, O.orc_prepass(src, abits, dither)), (w, h, comps, abits, dither)


def _tx(L, img, fmt, stride=0, pad=0):
    h, w, comps = img.shape
    bs = 8 if fmt in (0x83F0, 0x83F1) else 16
    tight = ((w + 3) // 4) * bs
    rb = stride if stride >= w * (bs // 4) else tight
    dest = np.full(((h + 3) // 4) * rb + tight + pad, 0xCD, np.uint8)
    L.tx_compress_dxtn(comps, w, h, img.ctypes.data, fmt, dest.ctypes.data, stride)
    return dest


def test_tx_compress_dxtn_environment_between_calls(L, capfd):
    """The reference reads the S2TC_* variables on EVERY call (ref :160-216); invalid values warn on stderr and keep the
    default.  Source and destination are plain numpy (pageable) memory."""
    img = synth.synth_rgba(52, 36, seed=9)
    saved = {k: os.environ.pop(k, None) for k in ("S2TC_DITHER_MODE", "S2TC_COLORDIST_MODE", "S2TC_RANDOM_COLORS", "S2TC_REFINE_COLORS")}
    try:
        # defaults: SIMPLE / WAVG / -1 / ALWAYS (ref :156-159)
        got = _tx(L, img, 0x83F1)
        want = O.orc_compress(img, O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE)
        assert np.array_equal(got[:want.size], want) and (got[want.size:] == 0xCD).all()
        seq = [({"S2TC_COLORDIST_MODE": "SRGB_MIXED", "S2TC_RANDOM_COLORS": "0", "S2TC_REFINE_COLORS": "LOOP"}, 0x83F3,
                (O.DXT5, O.SRGB_MIXED, 0, O.LOOP, O.DITHER_SIMPLE)),
               ({"S2TC_DITHER_MODE": "NONE", "S2TC_COLORDIST_MODE": "normalmap", "S2TC_REFINE_COLORS": "never"}, 0x83F2,
                (O.DXT3, O.NORMALMAP, 0, O.NEVER, O.DITHER_NONE)),
               ({"S2TC_DITHER_MODE": "FLOYDSTEINBERG", "S2TC_COLORDIST_MODE": "YUV", "S2TC_RANDOM_COLORS": "-1", "S2TC_REFINE_COLORS": "ALWAYS"},
                0x83F0, (O.DXT1, O.YUV, -1, O.ALWAYS, O.DITHER_FS))]
        for env, fmt, (dxt, cd, nr, rf, di) in seq:
            os.environ.update(env)
            got = _tx(L, img, fmt)
            want = O.orc_compress(img, dxt, cd, nr, rf, di)
            assert np.array_equal(got[:want.size], want), env
        capfd.readouterr()
        # invalid values: one warning each on stderr, defaults kept (ref :171,195,214)
        os.environ.update({"S2TC_DITHER_MODE": "bogus", "S2TC_COLORDIST_MODE": "nope", "S2TC_REFINE_COLORS": "sometimes",
                           "S2TC_RANDOM_COLORS": "-1"})
        got = _tx(L, img, 0x83F1)
        err = capfd.readouterr().err
        want = O.orc_compress(img, O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE)
        assert np.array_equal(got[:want.size], want)
        assert err.count("nvalid") >= 3, err
        # random colours: the process-wide rand() cursor continues from call to call (ref :984-992, never seeded)
        for k in ("S2TC_DITHER_MODE", "S2TC_COLORDIST_MODE", "S2TC_REFINE_COLORS"):
            os.environ.pop(k)
        os.environ["S2TC_RANDOM_COLORS"] = "5"
        L.s2tc_b200_rand_cursor_set(0)
        a = _tx(L, img, 0x83F3)
        blocks = 13 * 9
        assert L.s2tc_b200_rand_cursor_get() == blocks * 20
        b = _tx(L, img, 0x83F1)
        assert L.s2tc_b200_rand_cursor_get() == blocks * 20 + blocks * 15
        wa = O.orc_compress(img, O.DXT5, O.WAVG, 5, O.ALWAYS, O.DITHER_SIMPLE, cursor=0)
        wb = O.orc_compress(img, O.DXT1, O.WAVG, 5, O.ALWAYS, O.DITHER_SIMPLE, cursor=blocks * 20)
        assert np.array_equal(a[:wa.size], wa) and np.array_equal(b[:wb.size], wb)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def test_tx_compress_dxtn_bad_format_and_strides(L, capfd):
    img = synth.synth_noise(23, 14, seed=4)    # 6 x 4 blocks, ragged
    os.environ.pop("S2TC_RANDOM_COLORS", None)
    # bad destformat: message on stderr, dest untouched (ref :232-235)
    capfd.readouterr()
    got = _tx(L, img, 0x1234)
    assert (got == 0xCD).all()
    assert "Bad dstFormat" in capfd.readouterr().err
    for fmt, dxt in ((0x83F0, O.DXT1), (0x83F1, O.DXT1), (0x83F2, O.DXT3), (0x83F3, O.DXT5)):
        bs = O.block_bytes(dxt)
        tight = 6 * bs
        # 0 / below width*bs/4 -> tight rows; tight; padded; and a stride between width*bs/4 and the padded width, where block
        # rows overlap in dest and later rows win (ref :243,261,279 + the loop :246-258)
        for stride in (0, 23 * (bs // 4) - 1, tight, tight + 24, 23 * (bs // 4)):
            got = _tx(L, img, fmt, stride, pad=64)
            want = O.orc_compress(img, dxt, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE, stride=stride)
            rb = stride if stride >= 23 * (bs // 4) else tight
            if rb >= tight:
                for r in range(4):
                    assert np.array_equal(got[r * rb:r * rb + tight], want[r * rb:r * rb + tight]), (fmt, stride, r)
                    if r < 3:    # the padding between block rows is not written
                        assert (got[r * rb + tight:(r + 1) * rb] == 0xCD).all(), (fmt, stride, r)
            else:                # overlapping rows: the bytes every row leaves visible
                n = 3 * rb + tight
                assert np.array_equal(got[:n], want[:n]), (fmt, stride)
    # srccomps: 3 is RGB, anything else is treated as 4 (ref s2tc_algorithm.cpp:1455-1464)
    rgb = synth.synth_noise(16, 8, seed=8, comps=3)
    got = _tx(L, rgb, 0x83F3)
    want = O.orc_compress(rgb, O.DXT5, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE)
    assert np.array_equal(got[:want.size], want)
    # zero-sized image: the reference's loops do not run, nothing is written
    dest = np.full(64, 0xCD, np.uint8)
    L.tx_compress_dxtn(4, 0, 0, rgb.ctypes.data, 0x83F1, dest.ctypes.data, 0)
    assert (dest == 0xCD).all()


def test_contexts_on_two_streams_do_not_race(encoder):
    """Two asynchronous calls on ONE context with different streams share its workspaces: the second is ordered behind
    the first on the device.  Results must equal the single-stream ones."""
    import torch
    from s2tc_b200 import Settings
    img = synth.synth_rgba(1024, 1024, seed=5)
    d = torch.from_numpy(img).cuda()
    st = Settings(O.DXT5, O.WAVG, 0, O.LOOP, O.DITHER_SIMPLE)
    ref_a = encoder.compress(img, st)
    img2 = synth.synth_noise(1024, 1024, seed=6)
    d2 = torch.from_numpy(img2).cuda()
    ref_b = encoder.compress(img2, st)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    oa = torch.empty(ref_a.size, dtype=torch.uint8, device="cuda")
    ob = torch.empty(ref_b.size, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        encoder.encode_rows_device(d, 1024, 1024, 4, 0, 256, oa, st, stream=s1.cuda_stream)
        encoder.encode_rows_device(d2, 1024, 1024, 4, 0, 256, ob, st, stream=s2.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(oa.cpu().numpy(), ref_a) and np.array_equal(ob.cpu().numpy(), ref_b)


def test_summary_maps_are_not_reused_after_the_texels_changed(encoder):
    """the maps a summary leaves behind may only be reused by the explicit after-summary call; a plain encode
    of a buffer that was rewritten in place after a summary must not see stale maps."""
    import torch
    from s2tc_b200 import Settings
    a = synth.synth_rgba(512, 256, seed=1)
    b = synth.synth_noise(512, 256, seed=2)
    st = Settings(O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE)
    d = torch.from_numpy(a).cuda()
    out = torch.empty(128 * 64 * 8, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        encoder.dither_summary_device(d, 512, 256, 4, 1, 0, 64, stream=stream.cuda_stream)
        d.copy_(torch.from_numpy(b).cuda())          # same address, new texels
        encoder.encode_rows_device(d, 512, 256, 4, 0, 64, out, st, stream=stream.cuda_stream)
        stream.synchronize()
    assert np.array_equal(out.cpu().numpy(), O.orc_compress(b, O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE))


@pytest.mark.parametrize("dxt,cd,nr,rf,dither", [(O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE), (O.DXT5, O.SRGB_MIXED, 0, O.LOOP, O.DITHER_SIMPLE),
                                                 (O.DXT1, O.WAVG, 7, O.LOOP, O.DITHER_SIMPLE), (O.DXT3, O.RGB, -1, O.NEVER, O.DITHER_NONE)])
def test_pageable_buffers_are_staged(encoder, dxt, cd, nr, rf, dither):
    """Large malloc'd source and destination (what tx_compress_dxtn's callers pass): the texels travel through the pinned
    staging ring in 8 MiB chunks filled by the copy threads, the blocks come back through the pinned output buffer.  Same
    bytes as from pinned memory; first and last block rows against the oracle; 3-component source; the staging switch."""
    import torch
    from s2tc_b200 import Settings
    width, height = 4096, 2500 if nr <= 0 else 1100   # 39 MiB / 17 MiB of texels: several chunks, the last one partial
    img = synth.synth_rgba(width, height, seed=61)
    st = Settings(dxt, cd, nr, rf, dither)
    bw, bs = (width + 3) // 4, O.block_bytes(dxt)
    got = encoder.compress(img, st, cursor=5)                       # numpy in, numpy out: pageable both ways
    pinned_src = torch.from_numpy(img).pin_memory()
    pinned_dst = torch.empty(got.size, dtype=torch.uint8).pin_memory()
    encoder.compress(pinned_src, st, cursor=5, out=pinned_dst)
    assert np.array_equal(got, pinned_dst.numpy())
    bh = (height + 3) // 4
    for a, b in ((0, 2), (bh - 2, bh)):
        want = O.orc_rows(img, dxt, cd, nr, rf, dither, (a, b), cursor=5)
        assert np.array_equal(got[a * bw * bs:b * bw * bs], want), (a, b)
    rgb = np.ascontiguousarray(img[:1000, :, :3])                   # 12 MiB, 3 components
    if nr < 0:
        assert np.array_equal(encoder.compress(rgb, st, cursor=5), O.orc_compress(rgb, dxt, cd, nr, rf, dither, cursor=5))
