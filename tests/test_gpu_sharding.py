"""GPU tier: the block-row sharding primitive (s2tc_b200_encode_rows_device + dither summaries) on one GPU:
an image encoded as several independent row shards must equal the whole-image encode byte for byte, with the
DITHER_SIMPLE carry and the rand() cursor crossing the cuts."""
import numpy as np
import pytest
import torch

import _oracle as O
from s2tc_b200 import Settings, synth
from s2tc_b200.sharding import fold_carry, shard_block_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dxt,cd,nr,rf", [(O.DXT1, O.WAVG, -1, O.ALWAYS), (O.DXT5, O.SRGB_MIXED, 0, O.LOOP), (O.DXT1, O.WAVG, 9, O.LOOP),
                                          (O.DXT3, O.YUV, 3, O.NEVER), (O.DXT5, O.WAVG, 5, O.ALWAYS)])
@pytest.mark.parametrize("dither", [O.DITHER_NONE, O.DITHER_SIMPLE])
def test_sharded_rows_equal_whole_image(encoder, dxt, cd, nr, rf, dither):
    width, height, world = 200, 150, 3          # ragged: 38 block rows, the last one 2 texels high
    img = synth.synth_noise(width, height, seed=41)
    st = Settings(dxt, cd, nr, rf, dither)
    bs = O.block_bytes(dxt)
    abits = {0: 1, 1: 4, 2: 8}[dxt]
    bw, bh = (width + 3) // 4, (height + 3) // 4
    want = O.orc_compress(img, dxt, cd, nr, rf, dither, cursor=21)
    d_img = torch.from_numpy(img).cuda()
    ranges = [shard_block_rows(bh, world, r) for r in range(world)]
    shards = [d_img[4 * a:min(4 * b, height)].contiguous() for a, b in ranges]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        summaries = [encoder.dither_summary_device(s, width, height, 4, abits, a, b, stream=stream.cuda_stream)
                     for s, (a, b) in zip(shards, ranges)] if dither == O.DITHER_SIMPLE else None
        outs = []
        for r in (2, 0, 1):                       # any order: shards are independent once their carry is known
            a, b = ranges[r]
            carry = fold_carry(summaries, r, 4, abits) if summaries else None
            if summaries:                         # the summary of THIS shard last, so that its maps are the ones reused
                encoder.dither_summary_device(shards[r], width, height, 4, abits, a, b, stream=stream.cuda_stream)
            d_out = torch.zeros((b - a) * bw * bs, dtype=torch.uint8, device="cuda")
            encoder.encode_rows_device(shards[r], width, height, 4, a, b, d_out, st, cursor0=21, carry=carry,
                                       stream=stream.cuda_stream)
            outs.append((r, d_out))
        stream.synchronize()
    got = np.concatenate([o.cpu().numpy() for _, o in sorted(outs, key=lambda x: x[0])])
    assert np.array_equal(got, want)


def test_sharded_rows_async_device_side_carry(encoder):
    """The no-host-sync protocol bench.py uses at N > 1: summaries written to device memory, gathered (here: placed
    side by side by hand), folded on the device, and the carry handed to the encode as a device pointer."""
    import ctypes as C
    import s2tc_b200
    from s2tc_b200.api import _addr, _check, lib
    width, height, world = 200, 150, 4
    img = synth.synth_noise(width, height, seed=43)
    dxt, cd, nr, rf = O.DXT3, O.WAVG, 5, O.LOOP
    st = Settings(dxt, cd, nr, rf, O.DITHER_SIMPLE)
    s = st.c()
    bw, bh = (width + 3) // 4, (height + 3) // 4
    want = O.orc_compress(img, dxt, cd, nr, rf, O.DITHER_SIMPLE, cursor=3)
    d_img = torch.from_numpy(img).cuda()
    ranges = [shard_block_rows(bh, world, r) for r in range(world)]
    shards = [d_img[4 * a:min(4 * b, height)].contiguous() for a, b in ranges]
    maps_all = torch.zeros(16 * world, dtype=torch.int64, device="cuda")
    carry = torch.zeros(4, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    outs = []
    with torch.cuda.stream(stream):
        for r, (a, b) in enumerate(ranges):   # "all-gather": every shard's summary lands in its slot
            _check(lib().s2tc_b200_dither_summary_async(encoder._ctx, 4, 4, width, height, _addr(shards[r]), a, b,
                                                        maps_all[16 * r:].data_ptr(), stream.cuda_stream))
        for r, (a, b) in enumerate(ranges):
            _check(lib().s2tc_b200_fold_carry_async(encoder._ctx, _addr(maps_all), r, 4, 4, _addr(carry), stream.cuda_stream))
            d_out = torch.zeros((b - a) * bw * 16, dtype=torch.uint8, device="cuda")
            _check(lib().s2tc_b200_encode_rows_async(encoder._ctx, C.byref(s), 4, width, height, _addr(shards[r]), a, b,
                                                     _addr(d_out), 3, _addr(carry), stream.cuda_stream))
            outs.append(d_out)
        stream.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in outs])
    assert np.array_equal(got, want)
