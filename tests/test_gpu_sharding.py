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


@pytest.mark.parametrize("dxt,cd,nr,rf,dither", [(O.DXT1, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE), (O.DXT5, O.SRGB_MIXED, 0, O.LOOP, O.DITHER_SIMPLE),
                                                 (O.DXT1, O.WAVG, 9, O.LOOP, O.DITHER_SIMPLE), (O.DXT3, O.YUV, 3, O.NEVER, O.DITHER_NONE)])
@pytest.mark.parametrize("world,nwave,height,weights", [(3, 4, 150, None), (2, 8, 58, None), (1, 3, 150, None), (2, 6, 150, [1, 2, 4, 4, 3, 2])])
def test_striped_host_shards_equal_whole_image(dxt, cd, nr, rf, dither, world, nwave, height, weights):
    """s2tc_b200_compress_host_striped: `world` shards (threads, one context each, all on this GPU) encode one image in
    nwave * world stripes from host memory to host memory, exchanging one 128-byte summary per shard and wave; the
    assembled output must be the oracle's whole-image bytes (carry and rand() cursor cross every stripe cut; height 58
    gives 15 block rows for 16 stripes, so one stripe is empty; weights: waves of unequal size)."""
    import threading
    import s2tc_b200
    width = 200
    img = synth.synth_noise(width, height, seed=47)
    st = Settings(dxt, cd, nr, rf, dither)
    bs = O.block_bytes(dxt)
    bw, bh = (width + 3) // 4, (height + 3) // 4
    want = O.orc_compress(img, dxt, cd, nr, rf, dither, cursor=11)
    out = np.zeros(bw * bh * bs, np.uint8)
    maps_mine = [torch.zeros(16 * nwave, dtype=torch.int64, device="cuda") for _ in range(world)]
    maps_all = [torch.zeros(16 * nwave * world, dtype=torch.int64, device="cuda") for _ in range(world)]
    barrier = threading.Barrier(world)
    errors = []

    def run(rank):
        enc = s2tc_b200.Encoder(0)
        try:
            rows = [s2tc_b200.Encoder.stripe_rows(height, world, nwave, w, rank, weights) for w in range(nwave)]
            srcs = [np.ascontiguousarray(img[4 * a:min(4 * b, height)]) if b > a else None for a, b in rows]
            dsts = [out[a * bw * bs:b * bw * bs] if b > a else None for a, b in rows]

            def all_gather(w):           # stands in for ncclAllGather: everything is on one device here
                torch.cuda.synchronize()
                barrier.wait()
                for r in range(world):
                    maps_all[rank][16 * (world * w + r):16 * (world * w + r + 1)] = maps_mine[r][16 * w:16 * w + 16]
                torch.cuda.synchronize()

            enc.compress_striped(srcs, width, height, dsts, st, rank, world, nwave, maps_mine[rank], maps_all[rank], all_gather, cursor0=11,
                                 weights=weights)
        except Exception as e:           # pragma: no cover
            errors.append(e)
            barrier.abort()
        finally:
            enc.close()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert np.array_equal(out, want)


class _QueueDist:
    """send / recv between threads of one process (stands in for torch.distributed point-to-point on one GPU)"""

    def __init__(self, world):
        import queue
        self.q = {(a, b): queue.Queue() for a in range(world) for b in range(world)}
        self.rank = None

    def view(self, rank):
        v = _QueueDist.__new__(_QueueDist)
        v.q, v.rank = self.q, rank
        return v

    def send(self, t, dst):
        torch.cuda.synchronize()
        self.q[(self.rank, dst)].put(t.clone())

    def recv(self, t, src):
        t.copy_(self.q[(src, self.rank)].get(timeout=60))
        torch.cuda.synchronize()


@pytest.mark.parametrize("dxt,comps", [(O.DXT1, 4), (O.DXT3, 4), (O.DXT5, 4), (O.DXT1, 3)])
@pytest.mark.parametrize("world,width,height", [(3, 200, 150), (2, 64, 77), (4, 36, 128), (1, 50, 40)])
def test_floyd_steinberg_row_shards_equal_whole_image(dxt, comps, world, width, height):
    """DITHER_FLOYDSTEINBERG across row shards (s2tc_b200_floyd_rows_device chained by sharding.floyd_steinberg_sharded):
    every shard on its own context, error rows and the alpha seed passed between them; the encoded shards must be the
    oracle's whole-image bytes.  Odd and even heights (the alpha seed differs), shard cuts inside 32-row bands."""
    import threading
    import s2tc_b200
    from s2tc_b200.sharding import floyd_steinberg_sharded
    img = synth.synth_noise(width, height, seed=53, comps=comps)
    cd, nr, rf = O.WAVG, -1, O.ALWAYS
    st = Settings(dxt, cd, nr, rf, O.DITHER_FS)
    bs = O.block_bytes(dxt)
    abits = {0: 1, 1: 4, 2: 8}[dxt]
    bw, bh = (width + 3) // 4, (height + 3) // 4
    want = O.orc_compress(img, dxt, cd, nr, rf, O.DITHER_FS)
    d_img = torch.from_numpy(img).cuda()
    outs = [None] * world
    qd = _QueueDist(world)
    errors = []

    def run(rank):
        enc = s2tc_b200.Encoder(0)
        try:
            a, b = shard_block_rows(bh, world, rank)
            rows = d_img[4 * a:min(4 * b, height)].contiguous()
            reduced = torch.zeros(rows.shape[0] * width, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            def new_ints(n):             # torch fills on ITS stream; the encoder runs on its own: finish the fill first
                t = torch.zeros(n, dtype=torch.int32, device="cuda")
                torch.cuda.synchronize()
                return t

            keep = floyd_steinberg_sharded(enc, qd.view(rank), rows, width, height, comps, abits, a, b, rank, world, reduced, new_ints)
            enc.sync()                   # the exchange buffers were allocated on torch's stream: keep them until the encoder is done
            del keep
            d_out = torch.zeros((b - a) * bw * bs, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            enc.encode_reduced_rows_device(reduced, width, height, a, b, d_out, st)
            enc.sync()
            outs[rank] = d_out.cpu().numpy()
        except Exception as e:           # pragma: no cover
            errors.append(e)
        finally:
            enc.close()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert np.array_equal(np.concatenate(outs), want)
