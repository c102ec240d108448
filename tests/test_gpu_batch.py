"""GPU tier: batches of mip-mapped textures (BASELINE config 4; reference loop s2tc_compress.c:722-733 once per texture file
and per S2TC_COLORDIST_MODE).  s2tc_b200_compress_mipchain_batch_device pushes every mip level of ALL textures through
each kernel in one launch; the result must be, texture by texture and level by level, what one tx_compress_dxtn call per
level gives -- checked against the oracle on small textures (odd / non-square sizes, all three carry-restart paths of the
batched DITHER_SIMPLE, the 1-bit alpha channel, rand() candidates) and, at the size BASELINE names, for one 2048x2048
chain under all 8 metrics."""
import hashlib

import numpy as np
import pytest
import torch

import _oracle as O
import s2tc_b200
from s2tc_b200 import Settings, synth

pytestmark = pytest.mark.gpu


def _orc_chain(img, st, cursor=0):
    from test_oracle import orc_mip_reduce
    out, level = [], img
    while True:
        out.append(O.orc_compress(level, st.dxt, st.cd, st.nrandom, st.refine, st.dither, cursor=cursor))
        cursor += ((level.shape[1] + 3) // 4) * ((level.shape[0] + 3) // 4) * O.draws_per_block(st.dxt, st.nrandom)
        if level.shape[0] == 1 and level.shape[1] == 1:
            return np.concatenate(out)
        level = orc_mip_reduce(level)


def _run_batch(enc, texs, sets, cursor0=0):
    h, w = texs[0].shape[:2]
    ntex = len(texs)
    d = torch.from_numpy(np.stack(texs)).cuda()
    scratch = torch.empty(ntex * (w * h + w * h // 4) + 4096, dtype=torch.uint8, device="cuda")
    sizes = [s2tc_b200.lib().s2tc_b200_mipchain_bytes(s.dxt, w, h) for s in sets]
    offs = [0]
    for sz in sizes:
        offs.append((offs[-1] + sz * ntex + 15) & ~15)     # every setting's region starts 16-byte aligned
    dst = torch.full((offs[-1],), 0xEE, dtype=torch.uint8, device="cuda")
    before = d.clone()
    enc.compress_mipchain_batch_device(d, scratch, dst, w, h, ntex, sets, cursor0=cursor0)
    torch.cuda.synchronize()
    assert torch.equal(d, before), "the source textures must not be modified"
    out = dst.cpu().numpy()
    return [[out[offs[k] + i * sz:offs[k] + (i + 1) * sz] for i in range(ntex)] for k, sz in enumerate(sizes)]


SETS = [
    Settings(O.DXT3, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE),
    Settings(O.DXT3, O.NORMALMAP, -1, O.ALWAYS, O.DITHER_SIMPLE),     # shares the pre-pass with the one before
    Settings(O.DXT1, O.SRGB, 0, O.LOOP, O.DITHER_SIMPLE),             # other alpha width: 1-bit channel, new pre-pass
    Settings(O.DXT5, O.RGB, 3, O.NEVER, O.DITHER_NONE),               # rand() candidates: cursor restarts per texture
    Settings(O.DXT1, O.YUV, -1, O.ALWAYS, O.DITHER_FS),
    Settings(O.DXT5, O.SRGB_MIXED, 0, O.ALWAYS, O.DITHER_SIMPLE),
]


@pytest.mark.parametrize("shape", [(24, 40), (129, 37), (64, 64), (256, 256)])
def test_batch_equals_per_texture_oracle(encoder, shape):
    """(24,40)/(129,37): sizes that are not whole dither tiles (image-by-image path, odd levels drop a row/column);
    (64,64): one fused CTA per image from level 0; (256,256): 4 tiles per image, the carry restarts inside one launch."""
    h, w = shape
    texs = [synth.synth_rgba(w, h, seed=40 + i) if i % 2 == 0 else synth.synth_noise(w, h, seed=50 + i) for i in range(3)]
    got = _run_batch(encoder, texs, SETS, cursor0=5)
    for k, st in enumerate(SETS):
        for i, t in enumerate(texs):
            assert np.array_equal(got[k][i], _orc_chain(t, st, cursor=5)), (shape, k, i)


def test_batch_many_small_textures_bit1_restart(encoder):
    """32 textures of 512x512 (16 tiles each: restarts in the middle of scan parts) with partly transparent alpha, DXT1
    (the 1-bit alpha channel, whose transfer 'map' is a sum and needs the restart flag) -- against the per-texture path."""
    texs = [synth.synth_rgba(512, 512, seed=100 + i) for i in range(32)]
    sets = [Settings(O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE), Settings(O.DXT3, O.AVG, -1, O.LOOP, O.DITHER_SIMPLE)]
    got = _run_batch(encoder, texs, sets)
    for k, st in enumerate(sets):
        for i in (0, 1, 13, 31):
            assert np.array_equal(got[k][i], encoder.compress_mipchain(texs[i], st)), (k, i)
    want0 = _orc_chain(texs[31], sets[0])
    assert np.array_equal(got[0][31], want0)


def test_config4_full_size_chain_all_metrics(encoder):
    """BASELINE config 4 at its real texture size: 2048x2048, full chain (12 levels, 349 527 blocks), DXT3, all 8 metrics,
    REFINE=ALWAYS.  Two textures in the batch; every chain must equal the single-chain path byte for byte, the levels from
    256x256 down must equal the oracle, and level 0 is spot-checked against the oracle (first / last block rows)."""
    from test_oracle import orc_mip_reduce
    texs = [synth.synth_rgba(2048, 2048, seed=200), synth.synth_noise(2048, 2048, seed=201)]
    sets = [Settings(O.DXT3, cd, -1, O.ALWAYS, O.DITHER_SIMPLE) for cd in range(8)]
    got = _run_batch(encoder, texs, sets)
    chain_bytes = s2tc_b200.lib().s2tc_b200_mipchain_bytes(O.DXT3, 2048, 2048)
    assert chain_bytes == 349527 * 16
    # oracle: small levels of texture 0 (reduce on the CPU down to 256x256 first)
    level = texs[0]
    off = 0
    while level.shape[0] > 256:
        off += (level.shape[0] // 4) ** 2 * 16
        level = orc_mip_reduce(level)
    for cd in range(8):
        st = sets[cd]
        for i, t in enumerate(texs):
            single = encoder.compress_mipchain(t, st)
            assert hashlib.sha256(got[cd][i].tobytes()).digest() == hashlib.sha256(single.tobytes()).digest(), (cd, i)
        want_tail = _orc_chain(level, st)
        assert np.array_equal(got[cd][0][off:], want_tail), cd
        for r0, r1 in ((0, 2), (510, 512)):
            want = O.orc_rows(texs[0], O.DXT3, cd, -1, O.ALWAYS, O.DITHER_SIMPLE, (r0, r1))
            assert np.array_equal(got[cd][0][r0 * 512 * 16:r1 * 512 * 16], want), (cd, r0)
