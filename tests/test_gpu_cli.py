"""GPU tier, CLI level: our tools against the reference's own tools (compiled into oracle/_ref), and the
reference's s2tc_compress driving OUR libtxc_dxtn.so through its -l seam (the drop-in boundary)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import _oracle as O
import s2tc_b200
from s2tc_b200 import synth

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(s2tc_b200.lib_path()), "..", "bin")
OURS = os.path.join(BIN, "s2tc_compress")
OUR_LIB = s2tc_b200.lib_path("libtxc_dxtn.so")
REF = O.ref_path("s2tc_compress_ref")
REF_LIB = O.ref_path("libtxc_dxtn_ref.so")
need_ref_tools = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(REF_LIB)), reason="oracle/_ref tools not present")


def write_tga(path, img, bottom_up=True, rle=False):
    h, w = img.shape[:2]
    alpha = img.shape[2] == 4
    bgr = img[..., [2, 1, 0, 3]] if alpha else img[..., [2, 1, 0]]
    rows = bgr[::-1] if bottom_up else bgr
    hdr = struct.pack("<BBBHHBHHHHBB", 0, 0, 10 if rle else 2, 0, 0, 0, 0, 0, w, h, 32 if alpha else 24,
                      (8 if alpha else 0) | (0 if bottom_up else 0x20))
    body = bytearray()
    if not rle:
        body += rows.tobytes()
    else:
        bp = rows.shape[2]
        for row in rows:
            x = 0
            while x < w:   # raw packets of up to 128 texels, a run packet wherever two neighbours match
                if x + 1 < w and (row[x] == row[x + 1]).all():
                    n = 2
                    while x + n < w and n < 128 and (row[x + n] == row[x]).all():
                        n += 1
                    body += bytes([0x80 | (n - 1)]) + row[x].tobytes()
                else:
                    n = 1
                    while x + n < w and n < 128 and not (x + n + 1 < w and (row[x + n] == row[x + n + 1]).all()):
                        n += 1
                    body += bytes([n - 1]) + row[x:x + n].tobytes()
                x += n
        assert bp in (3, 4)
    with open(path, "wb") as f:
        f.write(hdr + bytes(body))


def run(cmd, env_extra=None, stdin=None):
    env = dict(os.environ)
    for k in ("S2TC_DITHER_MODE", "S2TC_COLORDIST_MODE", "S2TC_RANDOM_COLORS", "S2TC_REFINE_COLORS"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run(cmd, env=env, input=stdin, capture_output=True, check=True).stdout


SETTINGS = [
    {},                                                                                                   # library defaults
    {"S2TC_COLORDIST_MODE": "SRGB_MIXED", "S2TC_RANDOM_COLORS": "0", "S2TC_REFINE_COLORS": "LOOP"},
    {"S2TC_DITHER_MODE": "NONE", "S2TC_COLORDIST_MODE": "WAVG", "S2TC_RANDOM_COLORS": "8", "S2TC_REFINE_COLORS": "LOOP"},
    {"S2TC_COLORDIST_MODE": "normalmap", "S2TC_REFINE_COLORS": "never", "S2TC_DITHER_MODE": "simple"},
]


@need_ref_tools
@pytest.mark.parametrize("fmt", ["DXT1", "DXT3", "DXT5"])
def test_s2tc_compress_matches_reference_tool(tmp_path, fmt):
    imgs = {"a.tga": (synth.synth_rgba(100, 60, seed=1), True, False), "b.tga": (synth.synth_noise(40, 24, seed=2, comps=3), False, True)}
    for name, (img, bottom_up, rle) in imgs.items():
        path = str(tmp_path / name)
        write_tga(path, img, bottom_up, rle)
        for env in SETTINGS:
            want = run([REF, "-l", REF_LIB, "-t", fmt, "-i", path], env)
            got = run([OURS, "-t", fmt, "-i", path], env)
            # The 1x1 mip level of a DXT5 file is a single-texel block; in normal mode without random colours
            # the reference computes its alpha endpoints from uninitialised memory (known
            # divergence), so the last block is not comparable for exactly those settings.
            fast = "S2TC_RANDOM_COLORS" not in env and env.get("S2TC_COLORDIST_MODE", "").upper() != "NORMALMAP"
            keep = len(want) if fmt != "DXT5" or fast or int(env.get("S2TC_RANDOM_COLORS", "-1")) > 0 else len(want) - 16
            assert len(got) == len(want) and got[:keep] == want[:keep], (name, fmt, env)
            # the reference binary with OUR library behind its dlopen seam
            seam = run([REF, "-l", OUR_LIB, "-t", fmt, "-i", path], env)
            assert seam[:keep] == want[:keep], ("seam", name, fmt, env)


@need_ref_tools
def test_stdin_stdout_and_bad_arguments(tmp_path):
    img = synth.synth_rgba(16, 16, seed=9)
    path = str(tmp_path / "x.tga")
    write_tga(path, img)
    data = open(path, "rb").read()
    assert run([OURS, "-t", "dxt5"], stdin=data) == run([REF, "-l", REF_LIB, "-t", "dxt5"], stdin=data)
    assert subprocess.run([OURS, "-t", "DXT7"], input=data, capture_output=True).returncode == 1


@pytest.mark.skipif(not os.path.exists(O.ref_path("s2tc_from_s3tc_ref")), reason="oracle/_ref tools not present")
def test_s2tc_from_s3tc_matches_reference_tool(tmp_path):
    for dxt, cc in ((0, b"DXT1"), (1, b"DXT3"), (2, b"DXT5")):
        blocks = synth.synth_s3tc_blocks(3000, dxt, seed=4)
        hdr = bytearray(128)
        hdr[0:4] = b"DDS "
        hdr[84:88] = cc
        data = bytes(hdr) + blocks.tobytes() + b"\x01\x02\x03"   # trailing partial block is dropped by both
        want = run([O.ref_path("s2tc_from_s3tc_ref")], stdin=data)
        got = run([os.path.join(BIN, "s2tc_from_s3tc")], stdin=data)
        assert got == want, dxt
