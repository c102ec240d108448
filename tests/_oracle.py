"""ctypes access to the parity checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

`orc_*`  = our plain-C restatement (oracle/s2tc_oracle.c), always available.
`ref_*`  = the UNMODIFIED upstream sources compiled into oracle/_ref/ by oracle/Makefile.  Built in
           the authoring container (where /root/reference exists); on the GPU box the prebuilt
           files that travelled with the snapshot are used.  Nothing here reads /root/reference
           at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

DXT1, DXT3, DXT5 = 0, 1, 2
NEVER, ALWAYS, LOOP = 0, 1, 2
RGB, YUV, SRGB, SRGB_MIXED, AVG, WAVG, W0AVG, NORMALMAP = range(8)
DITHER_NONE, DITHER_SIMPLE, DITHER_FS = 0, 1, 2
GL_FORMAT = {DXT1: 0x83F1, DXT3: 0x83F2, DXT5: 0x83F3}
CD_NAMES = ["RGB", "YUV", "SRGB", "SRGB_MIXED", "AVG", "WAVG", "W0AVG", "NORMALMAP"]
REFINE_NAMES = ["NEVER", "ALWAYS", "LOOP"]
DITHER_NAMES = ["NONE", "SIMPLE", "FLOYDSTEINBERG"]
DXT_NAMES = ["DXT1", "DXT3", "DXT5"]

_u8p = C.POINTER(C.c_ubyte)


def _ptr(a):
    return a.ctypes.data_as(_u8p)


def build():
    """(Re)build oracle/liboracle.so and, where the upstream checkout is present, oracle/_ref/."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


class RandState(C.Structure):
    _fields_ = [("win", C.c_uint32 * 31), ("head", C.c_int), ("draws", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_rand_init.argtypes = [C.POINTER(RandState)]
        L.orc_rand_seek.argtypes = [C.POINTER(RandState), C.c_uint64]
        L.orc_rand_next.argtypes = [C.POINTER(RandState)]
        L.orc_rand_next.restype = C.c_int
        L.orc_color_dist.argtypes = [C.c_int, C.c_char_p, C.c_char_p]
        L.orc_color_dist.restype = C.c_int
        L.orc_rgb565_image.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_encode_block.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(RandState)]
        L.orc_compress_image.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_uint, _u8p, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.POINTER(RandState)]
        L.orc_compress_image.restype = C.c_int
        L.orc_transcode_blocks.argtypes = [_u8p, C.c_size_t, C.c_int]
        L.orc_fetch_texel.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _u8p]
        L.orc_mip_reduce.argtypes = [_u8p, _u8p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]
        L.refh_open.argtypes = [C.c_char_p]
        L.refh_open.restype = C.c_void_p
        L.refh_has_seek.argtypes = [C.c_void_p]
        L.refh_close.argtypes = [C.c_void_p]
        L.refh_encode_block.argtypes = [C.c_void_p, _u8p, _u8p] + [C.c_int] * 7
        L.refh_prepass.argtypes = [C.c_void_p, _u8p, _u8p] + [C.c_int] * 5
        L.refh_seek.argtypes = [C.c_void_p, C.c_uint64]
        L.refh_encode_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p, C.c_uint, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_uint64, _u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_double)]
        L.refh_encode_mt.restype = C.c_int
        L.refh_tx_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p, C.c_uint, _u8p, C.c_int]
        _lib = L
    return _lib


def block_bytes(dxt):
    return 8 if dxt == DXT1 else 16


def out_size(w, h, dxt):
    return ((w + 3) // 4) * ((h + 3) // 4) * block_bytes(dxt)


def draws_per_block(dxt, nrandom):
    return 0 if nrandom <= 0 else nrandom * (4 if dxt == DXT5 else 3)


# --------------------------------------------------------------------------- restatement

def orc_compress(img, dxt, cd=WAVG, nrandom=-1, refine=ALWAYS, dither=DITHER_SIMPLE, cursor=0, stride=0):
    """img: (H, W, 3|4) uint8.  Returns the encoded bytes as a uint8 array."""
    img = np.ascontiguousarray(img)
    h, w, comps = img.shape
    rows = (h + 3) // 4
    tight = ((w + 3) // 4) * block_bytes(dxt)
    row_bytes = stride if stride >= w * (block_bytes(dxt) // 4) else tight
    out = np.zeros(max(rows * row_bytes, 1) + tight, np.uint8)
    st = RandState()
    lib().orc_rand_seek(C.byref(st), cursor)
    rc = lib().orc_compress_image(comps, w, h, _ptr(img), GL_FORMAT[dxt], _ptr(out), stride, dither, cd, nrandom,
                                  refine, C.byref(st))
    assert rc == 0
    if stride == 0:
        return out[:rows * tight].copy()
    return out


def orc_rows(img, dxt, cd, nrandom, refine, dither, rows, cursor=0):
    """Block rows [rows[0], rows[1]) of the whole-image encode: the pre-pass runs over the full image (the carry chain
    needs it), only the requested block rows are searched.  Makes full-size spot checks affordable."""
    img = np.ascontiguousarray(img)
    h, w, comps = img.shape
    L = lib()
    L.orc_encode_block_rows.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_uint64, _u8p, C.c_int]
    red = orc_prepass(img, {DXT1: 1, DXT3: 4, DXT5: 8}[dxt], dither)
    bw = (w + 3) // 4
    full = np.zeros(((h + 3) // 4) * bw * block_bytes(dxt), np.uint8)
    L.orc_encode_block_rows(_ptr(red), w, h, rows[0], rows[1], dxt, cd, nrandom, refine, cursor, _ptr(full), 0)
    return full[rows[0] * bw * block_bytes(dxt):rows[1] * bw * block_bytes(dxt)].copy()


def orc_prepass(img, alphabits, dither):
    img = np.ascontiguousarray(img)
    h, w, comps = img.shape
    out = np.zeros((h, w, 4), np.uint8)
    lib().orc_rgb565_image(_ptr(out), _ptr(img), w, h, comps, alphabits, dither)
    return out


def orc_encode_block(px, w, h, dxt, cd, nrandom, refine, cursor=0, iw=4):
    """px: pre-reduced (rows, iw, 4) uint8 block storage."""
    px = np.ascontiguousarray(px, np.uint8)
    out = np.zeros(block_bytes(dxt), np.uint8)
    st = RandState()
    lib().orc_rand_seek(C.byref(st), cursor)
    lib().orc_encode_block(_ptr(out), _ptr(px), iw, w, h, dxt, cd, nrandom, refine, C.byref(st))
    return out


def orc_transcode(blocks, dxt):
    b = np.array(blocks, np.uint8, copy=True).reshape(-1)
    lib().orc_transcode_blocks(_ptr(b), b.size // block_bytes(dxt), dxt)
    return b


def orc_rand(n, start=0):
    st = RandState()
    lib().orc_rand_seek(C.byref(st), start)
    return [lib().orc_rand_next(C.byref(st)) for _ in range(n)]


# --------------------------------------------------------------------------- compiled reference

def ref_path(name="libtxc_dxtn_ref.so"):
    return os.path.join(REF_DIR, name)


def ref_available():
    return os.path.exists(ref_path()) and os.path.exists(ref_path("libtxc_dxtn_ref_tls.so"))


_refs = {}


def ref_handle(tls=True):
    """tls=True: the build whose rand() is the seekable thread-local replica (bit-identical stream,
    but positionable); tls=False: the plain build using libc rand()."""
    if tls not in _refs:
        h = lib().refh_open(ref_path("libtxc_dxtn_ref_tls.so" if tls else "libtxc_dxtn_ref.so").encode())
        assert h, "cannot open compiled reference"
        _refs[tls] = C.c_void_p(h)
    return _refs[tls]


def ref_compress(img, dxt, cd=WAVG, nrandom=-1, refine=ALWAYS, dither=DITHER_SIMPLE, cursor=0, stride=0,
                 threads=1, rows=None, times=None):
    img = np.ascontiguousarray(img)
    h, w, comps = img.shape
    nrows = (h + 3) // 4
    tight = ((w + 3) // 4) * block_bytes(dxt)
    row_bytes = stride if stride >= w * (block_bytes(dxt) // 4) else tight
    out = np.zeros(max(nrows * row_bytes, 1) + tight, np.uint8)
    r0, r1 = rows if rows is not None else (0, nrows)
    t = (C.c_double * 2)()
    rc = lib().refh_encode_mt(ref_handle(True), comps, w, h, _ptr(img), GL_FORMAT[dxt], dither, cd, nrandom, refine,
                              cursor, _ptr(out), stride, r0, r1, threads, t)
    assert rc == 0, rc
    if times is not None:
        times[:] = [t[0], t[1]]
    if stride == 0:
        return out[:nrows * tight].copy()
    return out


def ref_prepass(img, alphabits, dither):
    img = np.ascontiguousarray(img)
    h, w, comps = img.shape
    out = np.zeros((h, w, 4), np.uint8)
    lib().refh_prepass(ref_handle(True), _ptr(out), _ptr(img), w, h, comps, alphabits, dither)
    return out


def ref_encode_block(px, w, h, dxt, cd, nrandom, refine, cursor=0, iw=4):
    px = np.ascontiguousarray(px, np.uint8)
    out = np.zeros(block_bytes(dxt), np.uint8)
    hdl = ref_handle(True)
    lib().refh_seek(hdl, cursor)
    lib().refh_encode_block(hdl, _ptr(out), _ptr(px), iw, w, h, dxt, cd, nrandom, refine)
    return out


_tc = None


def ref_transcode(blocks, dxt):
    """Upstream convert_dxt1a / convert_dxt1 / convert_dxt5 applied per block as its main() does
    (s2tc_from_s3tc.cpp:254-263)."""
    global _tc
    if _tc is None:
        _tc = C.CDLL(ref_path("libfrom_s3tc_ref.so"))
    b = np.array(blocks, np.uint8, copy=True).reshape(-1)
    bs = block_bytes(dxt)
    base = b.ctypes.data
    f1a, f1, f5 = _tc._Z13convert_dxt1aPh, _tc._Z12convert_dxt1Ph, _tc._Z12convert_dxt5Ph
    for k in range(b.size // bs):
        p = C.c_void_p(base + k * bs)
        if dxt == DXT1:
            f1a(p)
        else:
            f1(C.c_void_p(base + k * bs + 8))
        if dxt == DXT5:
            f5(p)
    return b
