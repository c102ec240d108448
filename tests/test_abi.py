"""CPU tier: the C-ABI library loads and exports every symbol the headers under include/ declare, the
host-side settings logic mirrors the reference's, and -- with no GPU -- every compute entry point fails
loudly instead of falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

import s2tc_b200
from s2tc_b200 import Settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def declared_symbols():
    names = set()
    for fn in os.listdir(INCLUDE):
        text = open(os.path.join(INCLUDE, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text):
            name = m.group(1)
            if name not in ("defined", "sizeof", "void", "int") and not name.endswith("_t"):
                names.add(name)
    return names


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert {"tx_compress_dxtn", "fetch_2d_texel_rgba_dxt5", "s2tc_encode_block_func", "get_s2tc_encoder", "rgb565_image",
            "s2tc_b200_compress_host", "s2tc_b200_encode_rows_device", "s2tc_b200_transcode_device"} <= syms
    for lib in ("libs2tc_b200.so", "libtxc_dxtn.so"):
        L = ctypes.CDLL(s2tc_b200.lib_path(lib))
        missing = [s for s in sorted(syms) if not hasattr(L, s)]
        assert not missing, (lib, missing)


def test_reference_symbol_set_is_covered():
    """`nm -D` of the reference .so."""
    L = s2tc_b200.lib()
    for s in ("tx_compress_dxtn", "fetch_2d_texel_rgb_dxt1", "fetch_2d_texel_rgba_dxt1", "fetch_2d_texel_rgba_dxt3",
              "fetch_2d_texel_rgba_dxt5", "s2tc_encode_block_func", "rgb565_image"):
        assert hasattr(L, s)


def test_settings_from_env_mirrors_reference_parsing():
    env = {"S2TC_COLORDIST_MODE": "srgb_mixed", "S2TC_RANDOM_COLORS": "64", "S2TC_REFINE_COLORS": "Loop", "S2TC_DITHER_MODE": "none"}
    s = s2tc_b200.settings_from_env(s2tc_b200.DXT5, env)
    assert (s.dxt, s.cd, s.nrandom, s.refine, s.dither) == (2, s2tc_b200.SRGB_MIXED, 64, s2tc_b200.REFINE_LOOP, s2tc_b200.DITHER_NONE)
    s = s2tc_b200.settings_from_env(env={"S2TC_COLORDIST_MODE": "bogus", "S2TC_RANDOM_COLORS": "12abc", "S2TC_REFINE_COLORS": ""})
    assert (s.cd, s.nrandom, s.refine, s.dither) == (s2tc_b200.WAVG, 12, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_SIMPLE)
    assert s2tc_b200.settings_from_env(env={}) == Settings()   # WAVG, -1, ALWAYS, SIMPLE: ref s2tc_libtxc_dxtn.cpp:156-159


def test_texel_fetchers_match_oracle_decode():
    """The decode half of the libtxc_dxtn ABI is host code (as in the reference); check it on random blocks."""
    import _oracle as O
    rng = np.random.default_rng(3)
    L = s2tc_b200.lib()
    u8p = ctypes.POINTER(ctypes.c_ubyte)
    for dxt, fn, rgb in ((0, L.fetch_2d_texel_rgb_dxt1, 1), (0, L.fetch_2d_texel_rgba_dxt1, 0), (1, L.fetch_2d_texel_rgba_dxt3, 0),
                         (2, L.fetch_2d_texel_rgba_dxt5, 0)):
        data = rng.integers(0, 256, size=6 * 16, dtype=np.uint8)   # 3 x 2 blocks of a 12 x 8 image
        for j in range(8):
            for i in range(12):
                got = np.zeros(4, np.uint8)
                want = np.zeros(4, np.uint8)
                fn(12, data.ctypes.data, i, j, got.ctypes.data)
                O.lib().orc_fetch_texel(dxt, rgb, 12, data.ctypes.data_as(u8p), i, j, want.ctypes.data_as(u8p))
                assert np.array_equal(got, want), (dxt, i, j)


def test_no_gpu_means_loud_failure_not_fallback():
    if s2tc_b200.lib().s2tc_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(s2tc_b200.S2TCError) as e:
        s2tc_b200.Encoder(0)
    assert e.value.code == -1 and "no CPU path" in str(e.value)
    # the reference-facing entry point reports on stderr and leaves dest untouched (as upstream does for bad formats)
    src = np.zeros((4, 4, 4), np.uint8)
    dest = np.full(8, 0xAB, np.uint8)
    s2tc_b200.tx_compress_dxtn(4, 4, 4, src, 0x83F1, dest, 8)
    assert (dest == 0xAB).all()


def test_pure_host_helpers_work_without_gpu():
    ident = [int.from_bytes(bytes(range(8 * i, 8 * i + 8)), "little") for i in range(4)]
    maps = ident * 4
    assert s2tc_b200.Encoder.carry_apply(maps, 4, 4, [3, -2, -7, 11]) == [3, -2, -7, 11]
    assert s2tc_b200.draws_per_block(s2tc_b200.DXT5, 64) == 256 and s2tc_b200.draws_per_block(s2tc_b200.DXT1, 64) == 192
