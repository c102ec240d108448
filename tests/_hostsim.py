"""ctypes access to tests/hostsim/libhostsim.so: the encoder's host+device headers compiled for the CPU
(TEST INFRASTRUCTURE; see tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_u8p = C.POINTER(C.c_ubyte)
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libhostsim.so")
        subprocess.run(["make", "-s", "-C", HERE], check=True)
        L = C.CDLL(path)
        L.hostsim_compress.argtypes = [C.c_int] * 3 + [_u8p] + [C.c_int] * 5 + [C.c_uint64, _u8p]
        L.hostsim_prepass.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, _u8p, _u8p]
        L.hostsim_prepass_range.argtypes = [C.c_int, C.c_int, C.c_size_t, _u8p, C.POINTER(C.c_int), _u8p]
        L.hostsim_dither_summary.argtypes = [C.c_int, C.c_int, C.c_size_t, _u8p, C.POINTER(C.c_uint64)]
        L.hostsim_transcode.argtypes = [C.c_int, _u8p, C.c_size_t]
        L.hostsim_rand.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
        L.hostsim_color_dist.argtypes = [C.c_int, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def compress(img, dxt, cd, nr, rf, di, cursor=0):
    img = np.ascontiguousarray(img)
    h, w, c = img.shape
    bs = 8 if dxt == 0 else 16
    out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * bs, np.uint8)
    rc = lib().hostsim_compress(c, w, h, img.ctypes.data_as(_u8p), dxt, cd, nr, rf, di, cursor, out.ctypes.data_as(_u8p))
    assert rc == 0
    return out


def prepass_range(texels, comps, abits, carry):
    """DITHER_SIMPLE over a flat run of texels with carry in; returns (reduced, carry out)."""
    t = np.ascontiguousarray(texels).reshape(-1, comps)
    out = np.zeros((t.shape[0], 4), np.uint8)
    c = (C.c_int * 4)(*carry)
    lib().hostsim_prepass_range(comps, abits, t.shape[0], t.ctypes.data_as(_u8p), c, out.ctypes.data_as(_u8p))
    return out, list(c)


def dither_summary(texels, comps, abits):
    t = np.ascontiguousarray(texels).reshape(-1, comps)
    m = (C.c_uint64 * 16)()
    lib().hostsim_dither_summary(comps, abits, t.shape[0], t.ctypes.data_as(_u8p), m)
    return list(m)


def transcode(blocks, dxt):
    b = np.array(blocks, np.uint8, copy=True).reshape(-1)
    lib().hostsim_transcode(dxt, b.ctypes.data_as(_u8p), b.size // (8 if dxt == 0 else 16))
    return b


def rand(cursor, stride, t, n):
    out = (C.c_int * n)()
    lib().hostsim_rand(cursor, stride, t, n, out)
    return list(out)
