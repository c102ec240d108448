"""Deterministic synthetic textures.  numpy only.

The same host buffer feeds the GPU encoder and the CPU checker, so the generator's own float
behaviour never enters a parity comparison.  Images are built in row strips to keep peak memory
near the size of the result even at 16384 x 16384.
"""
import numpy as np

_STRIP = 512


def synth_rgba(width, height, seed=1234):
    """Low-frequency colour gradient + uniform noise of +-16; alpha plane in 32x32 patches:
    50 % opaque, 25 % fully transparent, 25 % a smooth ramp (exercises DXT1 transparency, DXT3
    nibbles and the DXT5 alpha search)."""
    rng = np.random.default_rng(seed)
    out = np.empty((height, width, 4), np.uint8)
    px = (width + 31) // 32
    py = (height + 31) // 32
    kind = rng.integers(0, 4, size=(py, px), dtype=np.uint8)  # 0,1 opaque; 2 transparent; 3 ramp
    x = np.arange(width, dtype=np.int32)[None, :]
    for y0 in range(0, height, _STRIP):
        y1 = min(height, y0 + _STRIP)
        y = np.arange(y0, y1, dtype=np.int32)[:, None]
        n = rng.integers(-16, 17, size=(y1 - y0, width, 3), dtype=np.int16)
        base = np.empty((y1 - y0, width, 3), np.int16)
        base[..., 0] = (x * 255) // max(width, 1)
        base[..., 1] = (y * 255) // max(height, 1)
        base[..., 2] = ((x + y) * 127) // max(width, 1)
        out[y0:y1, :, :3] = np.clip(base + n, 0, 255).astype(np.uint8)
        k = kind[(y // 32).ravel()][:, (x // 32).ravel()]
        ramp = ((x * 7 + y * 3) // 4) & 255
        a = np.where(k <= 1, 255, np.where(k == 2, 0, ramp)).astype(np.uint8)
        out[y0:y1, :, 3] = a
    return out


def synth_noise(width, height, seed=99, comps=4):
    """iid uniform bytes: the worst case, and the input that makes the SRGB metric wrap."""
    rng = np.random.default_rng(seed)
    out = np.empty((height, width, comps), np.uint8)
    for y0 in range(0, height, _STRIP):
        y1 = min(height, y0 + _STRIP)
        out[y0:y1] = rng.integers(0, 256, size=(y1 - y0, width, comps), dtype=np.uint8)
    return out


def synth_normal(width, height, seed=7):
    """Tangent-space normals of a sum-of-8-sines height field, round((n*0.5+0.5)*255); alpha = height."""
    rng = np.random.default_rng(seed)
    fx = rng.uniform(1.0, 24.0, 8)
    fy = rng.uniform(1.0, 24.0, 8)
    ph = rng.uniform(0.0, 2 * np.pi, 8)
    amp = rng.uniform(0.2, 1.0, 8)
    out = np.empty((height, width, 4), np.uint8)
    u = (np.arange(width, dtype=np.float32) / max(width, 1))[None, :]
    for y0 in range(0, height, _STRIP):
        y1 = min(height, y0 + _STRIP)
        v = (np.arange(y0, y1, dtype=np.float32) / max(height, 1))[:, None]
        hgt = np.zeros((y1 - y0, width), np.float32)
        dx = np.zeros_like(hgt)
        dy = np.zeros_like(hgt)
        for k in range(8):
            arg = (2 * np.pi) * (fx[k] * u + fy[k] * v) + ph[k]
            hgt += amp[k] * np.sin(arg)
            c = amp[k] * np.cos(arg)
            dx += c * fx[k]
            dy += c * fy[k]
        scale = 0.05
        nx, ny, nz = -dx * scale, -dy * scale, np.ones_like(hgt)
        inv = 1.0 / np.sqrt(nx * nx + ny * ny + nz * nz)
        out[y0:y1, :, 0] = np.rint((nx * inv * 0.5 + 0.5) * 255)
        out[y0:y1, :, 1] = np.rint((ny * inv * 0.5 + 0.5) * 255)
        out[y0:y1, :, 2] = np.rint((nz * inv * 0.5 + 0.5) * 255)
        out[y0:y1, :, 3] = np.clip(np.rint((hgt / amp.sum() * 0.5 + 0.5) * 255), 0, 255)
    return out


def synth_s3tc_blocks(nblocks, dxt, seed=5):
    """Random DXT1/DXT3/DXT5 blocks that use every S3TC index code (interpolated ones included),
    in both endpoint orders: input for the s2tc_from_s3tc transcode path."""
    rng = np.random.default_rng(seed)
    bs = 8 if dxt == 0 else 16
    blocks = rng.integers(0, 256, size=(nblocks, bs), dtype=np.uint8)
    # force a share of equal-endpoint blocks, the c1 >= c0 boundary of the transcoder
    eq = rng.random(nblocks) < 0.05
    off = 0 if dxt == 0 else 8
    blocks[eq, off + 2] = blocks[eq, off + 0]
    blocks[eq, off + 3] = blocks[eq, off + 1]
    if dxt == 2:
        eqa = rng.random(nblocks) < 0.05
        blocks[eqa, 1] = blocks[eqa, 0]
    return blocks
