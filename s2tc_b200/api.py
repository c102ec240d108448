"""ctypes binding of include/s2tc_b200.h, include/s2tc_b200_txc_dxtn.h and include/s2tc_b200_algorithm.h.

Names, argument meaning and error behaviour follow the reference interface: settings are the
(DxtMode, ColorDistMode, nrandom, RefinementMode, DitherMode) tuple of s2tc_algorithm.h:31-66 and
`tx_compress_dxtn` reads them from the S2TC_* environment on every call (s2tc_libtxc_dxtn.cpp:156-216).
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIBDIR = os.environ.get("S2TC_B200_LIBDIR") or os.path.join(_HERE, "lib")   # override: A/B builds of the kernels

# enumerators: values of the reference (s2tc_algorithm.h:31-63)
DITHER_NONE, DITHER_SIMPLE, DITHER_FLOYDSTEINBERG = 0, 1, 2
DXT1, DXT3, DXT5 = 0, 1, 2
REFINE_NEVER, REFINE_ALWAYS, REFINE_LOOP = 0, 1, 2
RGB, YUV, SRGB, SRGB_MIXED, AVG, WAVG, W0AVG, NORMALMAP = range(8)

GL_FORMAT = {DXT1: 0x83F1, DXT3: 0x83F2, DXT5: 0x83F3}
_CD_NAMES = ["RGB", "YUV", "SRGB", "SRGB_MIXED", "AVG", "WAVG", "W0AVG", "NORMALMAP"]
_REFINE_NAMES = ["NEVER", "ALWAYS", "LOOP"]
_DITHER_NAMES = ["NONE", "SIMPLE", "FLOYDSTEINBERG"]

_u8p = C.POINTER(C.c_ubyte)
GATHER_FN = C.CFUNCTYPE(None, C.c_void_p)   # the exchange callback of s2tc_b200_compress_host_shard
GATHER_WAVE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)   # ... of s2tc_b200_compress_host_striped (user, wave)


class S2TCError(RuntimeError):
    """A call into the CUDA library failed (code = S2TC_B200_E*)."""

    def __init__(self, code, msg):
        super().__init__(f"s2tc_b200 error {code}: {msg}")
        self.code = code


class _Settings(C.Structure):
    _fields_ = [("dxt", C.c_int), ("cd", C.c_int), ("nrandom", C.c_int), ("refine", C.c_int), ("dither", C.c_int)]


@dataclass
class Settings:
    """Defaults are the reference's (s2tc_libtxc_dxtn.cpp:156-159)."""
    dxt: int = DXT1
    cd: int = WAVG
    nrandom: int = -1
    refine: int = REFINE_ALWAYS
    dither: int = DITHER_SIMPLE

    def c(self):
        return _Settings(self.dxt, self.cd, self.nrandom, self.refine, self.dither)


def settings_from_env(dxt=DXT1, env=None):
    """The reference's environment parsing (case-insensitive, bad values keep the default)."""
    env = os.environ if env is None else env
    s = Settings(dxt=dxt)

    def pick(var, names, cur):
        v = env.get(var)
        if v is None:
            return cur
        for i, n in enumerate(names):
            if v.upper() == n:
                return i
        return cur

    s.dither = pick("S2TC_DITHER_MODE", _DITHER_NAMES, s.dither)
    s.cd = pick("S2TC_COLORDIST_MODE", _CD_NAMES, s.cd)
    s.refine = pick("S2TC_REFINE_COLORS", _REFINE_NAMES, s.refine)
    if "S2TC_RANDOM_COLORS" in env:
        s.nrandom = _atoi(env["S2TC_RANDOM_COLORS"])
    return s


def _atoi(v):
    v = v.strip()
    n = 0
    sign = 1
    i = 0
    if i < len(v) and v[i] in "+-":
        sign = -1 if v[i] == "-" else 1
        i += 1
    while i < len(v) and v[i].isdigit():
        n = n * 10 + int(v[i])
        i += 1
    return sign * n


def block_bytes(dxt):
    return 8 if dxt == DXT1 else 16


def draws_per_block(dxt, nrandom):
    return 0 if nrandom <= 0 else nrandom * (4 if dxt == DXT5 else 3)


def lib_path(name="libs2tc_b200.so"):
    return os.path.join(_LIBDIR, name)


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _CSRC, "-j", str(os.cpu_count() or 4), "all"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building s2tc_b200/csrc failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib():
    """The loaded shared object.  Missing library = ImportError: the encoder has no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C {_CSRC}` (or s2tc_b200.build()); "
                          "s2tc_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    sp = C.POINTER(_Settings)
    L.s2tc_b200_last_error.restype = C.c_char_p
    L.s2tc_b200_device_count.restype = i32
    L.s2tc_b200_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.s2tc_b200_ctx_destroy.argtypes = [vp]
    L.s2tc_b200_default_ctx.restype = vp
    L.s2tc_b200_compress_host.argtypes = [vp, sp, i32, i32, i32, vp, vp, i32, C.POINTER(u64)]
    L.s2tc_b200_encode_rows_device.argtypes = [vp, sp, i32, i32, i32, vp, i32, i32, vp, u64, C.POINTER(i32), vp]
    L.s2tc_b200_dither_summary_device.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, C.POINTER(u64), vp]
    L.s2tc_b200_dither_summary_async.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, vp, vp]
    L.s2tc_b200_fold_carry_async.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.s2tc_b200_encode_rows_async.argtypes = [vp, sp, i32, i32, i32, vp, i32, i32, vp, u64, vp, vp]
    L.s2tc_b200_encode_rows_after_summary_async.argtypes = [vp, sp, i32, i32, i32, vp, i32, i32, vp, u64, vp, vp]
    L.s2tc_b200_compress_host_shard.argtypes = [vp, sp, i32, i32, i32, vp, i32, i32, vp, u64, i32, i32, vp, vp, GATHER_FN, vp, vp]
    L.s2tc_b200_compress_host_striped.argtypes = [vp, sp, i32, i32, i32, C.POINTER(vp), C.POINTER(vp), u64, i32, i32, i32,
                                                  C.POINTER(i32), vp, vp, GATHER_WAVE_FN, vp, vp]
    L.s2tc_b200_floyd_rows_device.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp, vp, vp]
    L.s2tc_b200_encode_reduced_rows_device.argtypes = [vp, sp, i32, i32, vp, i32, i32, vp, u64, vp]
    L.s2tc_b200_stripe_rows.argtypes = [i32, i32, i32, C.POINTER(i32), i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.s2tc_b200_stripe_rows.restype = None
    L.s2tc_b200_carry_apply.argtypes = [C.POINTER(u64), i32, i32, i32, i32]
    L.s2tc_b200_mipchain_bytes.argtypes = [i32, i32, i32]
    L.s2tc_b200_mipchain_bytes.restype = C.c_size_t
    L.s2tc_b200_mip_reduce_device.argtypes = [vp, vp, i32, i32, vp, vp]
    L.s2tc_b200_compress_mipchain_device.argtypes = [vp, sp, i32, i32, vp, vp, vp, C.POINTER(u64), vp]
    L.s2tc_b200_compress_mipchain_host.argtypes = [vp, sp, i32, i32, vp, vp, C.POINTER(u64)]
    L.s2tc_b200_compress_mipchain_batch_device.argtypes = [vp, sp, i32, i32, i32, i32, vp, vp, vp, u64, vp]
    L.s2tc_b200_decode_device.argtypes = [vp, i32, vp, i32, i32, vp, vp]
    L.s2tc_b200_decode_host.argtypes = [vp, i32, vp, i32, i32, vp]
    L.s2tc_b200_rgb565_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32]
    L.s2tc_b200_encode_block_host.argtypes = [vp, sp, vp, vp, i32, i32, i32, C.POINTER(u64)]
    L.s2tc_b200_transcode_host.argtypes = [vp, i32, vp, C.c_size_t]
    L.s2tc_b200_transcode_device.argtypes = [vp, i32, vp, C.c_size_t, vp]
    L.s2tc_b200_rand_cursor_get.restype = u64
    L.s2tc_b200_rand_cursor_set.argtypes = [u64]
    L.s2tc_b200_sync.argtypes = [vp]
    L.s2tc_b200_launch_count.argtypes = [vp]
    L.s2tc_b200_launch_count.restype = u64
    L.s2tc_b200_profile_enable.argtypes = [vp, i32]
    L.s2tc_b200_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64), i32]
    L.s2tc_b200_int32_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.s2tc_b200_int_peaks.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tx_compress_dxtn.argtypes = [i32, i32, i32, vp, C.c_uint, vp, i32]
    L.tx_compress_dxtn.restype = None
    L.rgb565_image.argtypes = [vp, vp, i32, i32, i32, i32, i32]
    L.rgb565_image.restype = None
    blockfn = C.CFUNCTYPE(None, vp, vp, i32, i32, i32, i32)
    L.s2tc_encode_block_func.argtypes = [i32, i32, i32, i32]
    L.s2tc_encode_block_func.restype = blockfn
    L.get_s2tc_encoder.argtypes = [i32, i32, i32, i32]
    L.get_s2tc_encoder.restype = blockfn
    for f in ("fetch_2d_texel_rgb_dxt1", "fetch_2d_texel_rgba_dxt1", "fetch_2d_texel_rgba_dxt3", "fetch_2d_texel_rgba_dxt5"):
        getattr(L, f).argtypes = [i32, vp, i32, i32, vp]
        getattr(L, f).restype = None
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise S2TCError(rc, lib().s2tc_b200_last_error().decode(errors="replace"))


def _addr(x):
    """Host/device address of a numpy array, torch tensor or int."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class Encoder:
    """One encoder context on one GPU (owns a CUDA stream and its workspaces)."""

    FAMILIES = ("prepass", "candidates", "search", "finish", "fast", "transcode")

    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        _check(lib().s2tc_b200_ctx_create(int(device), C.byref(self._ctx)))
        self.device = int(device)

    def close(self):
        if self._ctx:
            lib().s2tc_b200_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers -------------------------------------------------------------------------
    def compress(self, img, settings, cursor=0, stride=0, out=None, return_cursor=False):
        """img: (H, W, 3|4) uint8 numpy array (or a pinned torch CPU tensor).  Returns the encoded bytes
        exactly as tx_compress_dxtn lays them out for dstRowStride = `stride`."""
        h, w, comps = img.shape
        bs = block_bytes(settings.dxt)
        bw, bh = (w + 3) // 4, (h + 3) // 4
        tight = bw * bs
        row_bytes = stride if stride >= w * (bs // 4) else tight
        nbytes = max(bh * row_bytes, 1) + (tight if stride else 0)
        if out is None:
            out = np.zeros(nbytes, np.uint8)
        cur = C.c_uint64(cursor)
        s = settings.c()
        _check(lib().s2tc_b200_compress_host(self._ctx, C.byref(s), comps, w, h, _addr(img), _addr(out), stride,
                                             C.byref(cur)))
        if stride == 0 and isinstance(out, np.ndarray):
            out = out[:bh * tight]
        return (out, cur.value) if return_cursor else out

    def compress_mipchain(self, img, settings, cursor=0, return_cursor=False):
        """img: (H, W, 4) uint8.  Every mip level down to 1x1, encoded back to back (the payload of the DDS file the
        reference's s2tc_compress writes)."""
        h, w, comps = img.shape
        assert comps == 4
        out = np.zeros(lib().s2tc_b200_mipchain_bytes(settings.dxt, w, h), np.uint8)
        cur = C.c_uint64(cursor)
        s = settings.c()
        _check(lib().s2tc_b200_compress_mipchain_host(self._ctx, C.byref(s), w, h, _addr(np.ascontiguousarray(img)), _addr(out),
                                                      C.byref(cur)))
        return (out, cur.value) if return_cursor else out

    def compress_mipchain_device(self, rgba, scratch, dst, width, height, settings, cursor=0, stream=None):
        cur = C.c_uint64(cursor)
        s = settings.c()
        _check(lib().s2tc_b200_compress_mipchain_device(self._ctx, C.byref(s), width, height, _addr(rgba), _addr(scratch), _addr(dst),
                                                        C.byref(cur), stream))
        return cur.value

    def compress_mipchain_batch_device(self, rgba, scratch, dst, width, height, ntex, settings_list, cursor0=0, stream=None):
        """ntex textures back to back in `rgba` (device), every mip level of all of them per launch, once per entry of
        settings_list; chains are written setting-major, texture-minor into `dst` (see include/s2tc_b200.h)."""
        arr = (_Settings * len(settings_list))(*[s.c() for s in settings_list])
        _check(lib().s2tc_b200_compress_mipchain_batch_device(self._ctx, arr, len(settings_list), width, height, ntex, _addr(rgba),
                                                              _addr(scratch), _addr(dst), cursor0, stream))

    def rgb565_image(self, img, alphabits, dither):
        h, w, comps = img.shape
        out = np.zeros((h, w, 4), np.uint8)
        _check(lib().s2tc_b200_rgb565_host(self._ctx, _addr(out), _addr(np.ascontiguousarray(img)), w, h, comps, alphabits,
                                           dither))
        return out

    def encode_block(self, px, w, h, settings, cursor=0, iw=4):
        """px: pre-reduced texels, row stride iw."""
        px = np.ascontiguousarray(px, np.uint8)
        out = np.zeros(block_bytes(settings.dxt), np.uint8)
        cur = C.c_uint64(cursor)
        s = settings.c()
        _check(lib().s2tc_b200_encode_block_host(self._ctx, C.byref(s), _addr(out), _addr(px), iw, w, h, C.byref(cur)))
        return out

    def decode(self, blocks, dxt, width, height):
        """S2TC blocks -> (H, W, 4) RGBA8, the result of the reference's per-texel fetchers applied to every texel."""
        b = np.ascontiguousarray(blocks, np.uint8).reshape(-1)
        out = np.zeros((height, width, 4), np.uint8)
        _check(lib().s2tc_b200_decode_host(self._ctx, dxt, _addr(b), width, height, _addr(out)))
        return out

    def transcode(self, blocks, dxt):
        b = np.array(blocks, np.uint8, copy=True).reshape(-1)
        _check(lib().s2tc_b200_transcode_host(self._ctx, dxt, _addr(b), b.size // block_bytes(dxt)))
        return b

    # ---- device buffers (torch tensors or raw addresses) ----------------------------------------
    def encode_rows_device(self, src_rows, width, height, comps, row0, row1, dst, settings, cursor0=0, carry=None,
                           stream=None):
        """src_rows: device buffer holding texel rows 4*row0 .. of a width x height image.
        carry: None, or a list of 4 ints updated in place (DITHER_SIMPLE carry in/out; synchronises)."""
        s = settings.c()
        cptr = None
        if carry is not None:
            arr = (C.c_int * 4)(*carry)
            cptr = arr
        _check(lib().s2tc_b200_encode_rows_device(self._ctx, C.byref(s), comps, width, height, _addr(src_rows), row0, row1,
                                                  _addr(dst), cursor0, cptr, stream))
        if carry is not None:
            carry[:] = list(arr)

    def sharded_encode_async(self, src_rows, width, height, comps, row0, row1, dst, settings, maps_mine, all_gather, maps_all,
                             rank, carry_dev, cursor0=0, stream=None):
        """One shard of a DITHER_SIMPLE image without host synchronisation: summary -> all_gather() (a callable that
        gathers `maps_mine` (16 int64, device) of every rank into `maps_all` on the same stream) -> fold -> encode."""
        s = settings.c()
        abits = {DXT1: 1, DXT3: 4, DXT5: 8}[settings.dxt]
        if settings.dither == DITHER_SIMPLE:
            _check(lib().s2tc_b200_dither_summary_async(self._ctx, comps, abits, width, height, _addr(src_rows), row0, row1,
                                                        _addr(maps_mine), stream))
            all_gather()
            _check(lib().s2tc_b200_fold_carry_async(self._ctx, _addr(maps_all), rank, comps, abits, _addr(carry_dev), stream))
        _check(lib().s2tc_b200_encode_rows_after_summary_async(self._ctx, C.byref(s), comps, width, height, _addr(src_rows), row0,
                                                               row1, _addr(dst), cursor0,
                                                               _addr(carry_dev) if settings.dither == DITHER_SIMPLE else None, stream))

    def compress_shard(self, src_rows, width, height, row0, row1, dst, settings, rank, nslab, maps_mine, maps_all, all_gather,
                       cursor0=0, stream=None):
        """Host to host: block rows [row0, row1) of a width x height image (src_rows: the shard's texel rows, a numpy
        array or pinned torch tensor of shape (rows, width, comps)); pipelined in `nslab` pieces.  all_gather: a callable
        that gathers `maps_mine` (nslab * 16 int64, device) of every rank into `maps_all` on `stream` (DITHER_SIMPLE)."""
        s = settings.c()
        comps = src_rows.shape[2]
        cb = GATHER_FN(lambda _user: all_gather())
        _check(lib().s2tc_b200_compress_host_shard(self._ctx, C.byref(s), comps, width, height, _addr(src_rows), row0, row1,
                                                   _addr(dst), cursor0, rank, nslab, _addr(maps_mine), _addr(maps_all), cb, None,
                                                   stream))

    def floyd_rows_device(self, src_rows, width, height, comps, alphabits, row0, row1, phase, err_in, err_out, reduced_rows,
                          stream=None):
        """One DITHER_FLOYDSTEINBERG pass (0: colour, 1: alpha) over the texel rows of block rows [row0, row1); err_in /
        err_out: device int32 tensors ([3 * width] / [width]) or None (s2tc_b200_floyd_rows_device)."""
        _check(lib().s2tc_b200_floyd_rows_device(self._ctx, comps, alphabits, width, height, _addr(src_rows), row0, row1, phase,
                                                 None if err_in is None else _addr(err_in),
                                                 None if err_out is None else _addr(err_out), _addr(reduced_rows), stream))

    def encode_reduced_rows_device(self, reduced_rows, width, height, row0, row1, dst, settings, cursor0=0, stream=None):
        s = settings.c()
        _check(lib().s2tc_b200_encode_reduced_rows_device(self._ctx, C.byref(s), width, height, _addr(reduced_rows), row0, row1,
                                                          _addr(dst), cursor0, stream))

    @staticmethod
    def stripe_rows(height, world, nwave, wave, rank, weights=None):
        """Block rows [row0, row1) of stripe wave * world + rank (s2tc_b200_stripe_rows); weights: relative wave sizes."""
        a, b = C.c_int(), C.c_int()
        wts = (C.c_int * nwave)(*weights) if weights is not None else None
        lib().s2tc_b200_stripe_rows(height, world, nwave, wts, wave, rank, C.byref(a), C.byref(b))
        return a.value, b.value

    def compress_striped(self, src_stripes, width, height, dst_stripes, settings, rank, world, nwave, maps_mine=None,
                         maps_all=None, all_gather=None, comps=4, cursor0=0, stream=None, weights=None):
        """Host to host, one texture on `world` shards in nwave * world stripes (s2tc_b200_compress_host_striped).
        src_stripes[w] / dst_stripes[w]: this shard's stripe of wave w (numpy arrays or pinned torch tensors; None for an
        empty stripe).  all_gather(w): gathers maps_mine[16 w : 16 w + 16] (int64, device) of every rank into
        maps_all[16 world w : 16 world (w + 1)] on `stream` (DITHER_SIMPLE, world > 1)."""
        s = settings.c()
        srcs = (C.c_void_p * nwave)(*[None if a is None else _addr(a) for a in src_stripes])
        dsts = (C.c_void_p * nwave)(*[None if a is None else _addr(a) for a in dst_stripes])
        cb = GATHER_WAVE_FN((lambda _user, w: all_gather(w)) if all_gather else (lambda _user, w: None))
        wts = (C.c_int * nwave)(*weights) if weights is not None else None
        _check(lib().s2tc_b200_compress_host_striped(self._ctx, C.byref(s), comps, width, height, srcs, dsts, cursor0, rank, world,
                                                     nwave, wts, None if maps_mine is None else _addr(maps_mine),
                                                     None if maps_all is None else _addr(maps_all), cb, None, stream))

    def dither_summary_device(self, src_rows, width, height, comps, alphabits, row0, row1, stream=None):
        maps = (C.c_uint64 * 16)()
        _check(lib().s2tc_b200_dither_summary_device(self._ctx, comps, alphabits, width, height, _addr(src_rows), row0, row1,
                                                     maps, stream))
        return list(maps)

    @staticmethod
    def carry_apply(maps, comps, alphabits, carry):
        """Carry leaving a texel range with transfer maps `maps` (16 words) for an incoming carry (4 ints)."""
        out = []
        for ch in range(4):
            m = (C.c_uint64 * 4)(*maps[4 * ch:4 * ch + 4])
            out.append(lib().s2tc_b200_carry_apply(m, ch, comps, alphabits, carry[ch]))
        return out

    def transcode_device(self, blocks, dxt, nblocks, stream=None):
        _check(lib().s2tc_b200_transcode_device(self._ctx, dxt, _addr(blocks), nblocks, stream))

    # ---- measurement ------------------------------------------------------------------------------
    def sync(self):
        _check(lib().s2tc_b200_sync(self._ctx))

    def launch_count(self):
        return int(lib().s2tc_b200_launch_count(self._ctx))

    def int32_peak_gops(self):
        g = C.c_double()
        _check(lib().s2tc_b200_int32_peak(self._ctx, C.byref(g)))
        return g.value

    def int_peaks_gops(self):
        """(scalar 32-bit, 16-bit packed) sustained min+add rates in Gop/s"""
        a, b = C.c_double(), C.c_double()
        _check(lib().s2tc_b200_int_peaks(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile(self, on=True):
        _check(lib().s2tc_b200_profile_enable(self._ctx, 1 if on else 0))

    def profile_read(self, reset=True):
        ms = (C.c_double * 6)()
        n = (C.c_uint64 * 6)()
        _check(lib().s2tc_b200_profile_read(self._ctx, ms, n, 1 if reset else 0))
        return {f: (ms[i], int(n[i])) for i, f in enumerate(self.FAMILIES)}


def tx_compress_dxtn(srccomps, width, height, src, destformat, dest, dst_row_stride):
    """The libtxc_dxtn entry point itself (reads the S2TC_* environment; errors go to stderr)."""
    lib().tx_compress_dxtn(srccomps, width, height, _addr(src), destformat, _addr(dest), dst_row_stride)
