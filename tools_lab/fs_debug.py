"""Debug helper: Floyd-Steinberg row shards run sequentially on ONE context (no threads), reduced texels compared with the
oracle's pre-pass row by row."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _oracle as O
import s2tc_b200
from s2tc_b200 import synth
from s2tc_b200.sharding import shard_block_rows

def run(world, width, height, comps, abits, seed=53, reps=5):
    img = synth.synth_noise(width, height, seed=seed, comps=comps)
    want = O.orc_prepass(img, abits, 2).reshape(height, width, 4)
    bh = (height + 3) // 4
    enc = s2tc_b200.Encoder(0)
    d_img = torch.from_numpy(img).cuda()
    for rep in range(reps):
        reds, err = [], None
        ranges = [shard_block_rows(bh, world, r) for r in range(world)]
        outs = []
        for r, (a, b) in enumerate(ranges):
            rows = d_img[4 * a:min(4 * b, height)].contiguous()
            red = torch.zeros(rows.shape[0] * width, dtype=torch.int32, device="cuda")
            eo = torch.zeros(3 * width, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            enc.floyd_rows_device(rows, width, height, comps, abits, a, b, 0, err, eo, red)
            enc.sync()
            err = eo
            reds.append(red)
            outs.append(eo)
        if comps == 4 and abits != 8:
            err = outs[-1][:width].clone()
            for r, (a, b) in enumerate(ranges):
                rows = d_img[4 * a:min(4 * b, height)].contiguous()
                eo = torch.zeros(width, dtype=torch.int32, device="cuda")
                torch.cuda.synchronize()
                enc.floyd_rows_device(rows, width, height, comps, abits, a, b, 1, err, eo, reds[r])
                enc.sync()
                err = eo
        got = np.concatenate([x.cpu().numpy().view(np.uint8).reshape(-1, width, 4) for x in reds])
        bad = np.argwhere(got != want)
        print(f"world {world} {width}x{height} comps {comps} abits {abits} rep {rep}: {len(bad)} bad bytes",
              ("first " + str(bad[:6].tolist()) + " got " + str([int(got[tuple(i)]) for i in bad[:6]]) + " want " + str([int(want[tuple(i)]) for i in bad[:6]])) if len(bad) else "")
    enc.close()

run(2, 64, 77, 3, 1)
run(2, 64, 77, 4, 1)
run(2, 64, 76, 3, 1)
run(1, 64, 77, 3, 1)
run(3, 64, 77, 3, 1)
run(2, 66, 77, 3, 1)
