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
    ret[rank] = True
    dist.destroy_process_group()


def test_two_rank_sharded_encode_logic():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
