"""CPU tier, world_size 2 over gloo: the host-side logic of a block-row sharded encode (what bench.py and a
multi-GPU caller run) -- row partition, rand() cursor offsets, and the DITHER_SIMPLE carry exchange (an
all-gather of 128-byte transfer functions).  The device kernels are replaced by their CPU twins
(tests/hostsim = the same source compiled for the host, and the oracle); everything else -- including the
pure-host s2tc_b200_carry_apply of the real library -- is the code the GPU path uses."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _hostsim as H
    import _oracle as O
    from s2tc_b200 import synth
    from s2tc_b200.sharding import fold_carry, shard_block_rows, summary_from_i64, summary_to_i64

    width, height = 64, 72
    img = synth.synth_noise(width, height, seed=31)          # every rank can regenerate the texture
    bh = (height + 3) // 4
    row0, row1 = shard_block_rows(bh, world, rank)
    mine = img[4 * row0:4 * row1]

    # 1. DITHER_SIMPLE carry: summarise own texels, all-gather, fold the ranks before us
    for dxt, abits in ((O.DXT1, 1), (O.DXT3, 4), (O.DXT5, 8)):
        summary = torch.tensor(summary_to_i64(H.dither_summary(mine, 4, abits)), dtype=torch.int64)
        gathered = [torch.empty_like(summary) for _ in range(world)]
        dist.all_gather(gathered, summary)
        carry = fold_carry([summary_from_i64(g.tolist()) for g in gathered], rank, 4, abits)
        reduced, _ = H.prepass_range(mine, 4, abits, carry)
        want = O.orc_prepass(img, abits, O.DITHER_SIMPLE)[4 * row0:4 * row1]
        assert np.array_equal(reduced.reshape(want.shape), want), ("carry", rank, dxt)

    # 2. rand() cursor: a shard starts at cursor0 + row0 * blocks_per_row * draws_per_block
    full = O.orc_compress(img, O.DXT5, O.WAVG, 5, O.LOOP, O.DITHER_NONE, cursor=11)
    bw = (width + 3) // 4
    dpb = 5 * 4
    part = O.orc_compress(mine, O.DXT5, O.WAVG, 5, O.LOOP, O.DITHER_NONE, cursor=11 + row0 * bw * dpb)
    assert np.array_equal(part, full[row0 * bw * 16:row1 * bw * 16]), ("cursor", rank)

    # 3. the gather of the output slices is a plain concatenation in rank order
    out = torch.from_numpy(part.copy())
    sizes = [(shard_block_rows(bh, world, r)[1] - shard_block_rows(bh, world, r)[0]) * bw * 16 for r in range(world)]
    bufs = [torch.empty(s, dtype=torch.uint8) for s in sizes]
    dist.all_gather(bufs, out) if len(set(sizes)) == 1 else None
    if len(set(sizes)) == 1:
        assert np.array_equal(torch.cat(bufs).numpy(), full)
    # 4. striped shards (s2tc_b200_compress_host_striped): nwave * world stripes, stripe w * world + rank on this rank; per
    #    wave one all-gather of the stripes' summaries, the carry entering a stripe = carry entering the wave folded over
    #    the stripes of the ranks before us, the carry entering the next wave = folded over all of them
    import s2tc_b200
    nwave = 3
    for dxt, abits in ((O.DXT1, 1), (O.DXT5, 8)):
        want = O.orc_prepass(img, abits, O.DITHER_SIMPLE)
        wave_carry = [0, 0, 0, 0]
        for w in range(nwave):
            a, b = s2tc_b200.Encoder.stripe_rows(height, world, nwave, w, rank)
            wa, wb = bh * w // nwave, bh * (w + 1) // nwave      # the wave's rows, cut evenly among the ranks
            assert (a, b) == (wa + (wb - wa) * rank // world, wa + (wb - wa) * (rank + 1) // world)
            stripe = img[4 * a:4 * b]
            summary = torch.tensor(summary_to_i64(H.dither_summary(stripe, 4, abits)), dtype=torch.int64)
            gathered = [torch.empty_like(summary) for _ in range(world)]
            dist.all_gather(gathered, summary)
            sums = [summary_from_i64(g.tolist()) for g in gathered]
            carry = fold_carry(sums, rank, 4, abits, wave_carry)
            wave_carry = fold_carry(sums, world, 4, abits, wave_carry)
            reduced, _ = H.prepass_range(stripe, 4, abits, carry)
            assert np.array_equal(reduced.reshape(want[4 * a:4 * b].shape), want[4 * a:4 * b]), ("stripe", rank, w, dxt)

    # 5. the Floyd-Steinberg chain (sharding.floyd_steinberg_sharded) over real point-to-point messages: a stand-in encoder
    #    whose "passes" are a recurrence over rows with the same data flow (error row in, error row out, the alpha seed
    #    leaving the image's last rows through the first `width` ints of the colour pass's output)
    from s2tc_b200.sharding import floyd_steinberg_sharded

    class FakeEnc:
        def floyd_rows_device(self, src_rows, width, height, comps, alphabits, row0, row1, phase, err_in, err_out, reduced, stream=None):
            n = 3 * width if phase == 0 else width
            e = torch.zeros(n, dtype=torch.int32) if err_in is None else err_in.clone()
            for y in range(src_rows.shape[0]):
                e = (e * 3 + int(src_rows[y].sum()) + phase) % 1000003
                reduced[y, phase] = int(e.sum())
            last = row1 == (height + 3) // 4
            if phase == 0 and last:
                err_out[:width] = (e[:width] * 7 + 1) % 1000003   # the alpha seed
            else:
                err_out[:n] = e

    def fs_chain(ranges):
        """what the chain must compute, sequentially"""
        red = torch.zeros((height, 2), dtype=torch.int64)
        fake, e_in, outs = FakeEnc(), None, None
        for a, b in ranges:
            outs = torch.zeros(3 * width, dtype=torch.int32)
            fake.floyd_rows_device(torch.from_numpy(img[4 * a:4 * b].astype(np.int64)), width, height, 4, 1, a, b, 0, e_in, outs, red[4 * a:4 * b])
            e_in = outs
        e_in = outs[:width].clone()
        for a, b in ranges:
            outs = torch.zeros(width, dtype=torch.int32)
            fake.floyd_rows_device(torch.from_numpy(img[4 * a:4 * b].astype(np.int64)), width, height, 4, 1, a, b, 1, e_in, outs, red[4 * a:4 * b])
            e_in = outs
        return red

    ranges = [shard_block_rows(bh, world, r) for r in range(world)]
    red = torch.zeros((4 * (row1 - row0), 2), dtype=torch.int64)
    floyd_steinberg_sharded(FakeEnc(), dist, torch.from_numpy(mine.astype(np.int64)), width, height, 4, 1, row0, row1, rank, world, red,
                            lambda n: torch.zeros(n, dtype=torch.int32))
    assert torch.equal(red, fs_chain(ranges)[4 * row0:4 * row1]), ("floyd chain", rank)
    ret[rank] = True
    dist.destroy_process_group()


def test_two_rank_sharded_encode_logic():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
