#!/usr/bin/env python
"""bench_extra.py -- the one BASELINE.json workload that is not an encode: the s2tc_from_s3tc transcode of config 5
(reference s2tc_from_s3tc.cpp:254-263) on 16.7 M DXT5 blocks resident in HBM: pure streaming, 32 B of traffic per block.
Every encode workload, the batch of config 4 included, is a `bench.py --workload ...`.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import _oracle as O
    import s2tc_b200
    from s2tc_b200 import synth
    from bench import read_peaks

    torch.cuda.set_device(0)
    enc = s2tc_b200.Encoder(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    peak_gbs, _, peak_src = read_peaks()
    nblocks = 1 << 24
    blocks = torch.from_numpy(synth.synth_s3tc_blocks(1 << 20, 2, seed=9)).cuda().repeat(16, 1).contiguous()
    ref_in = synth.synth_s3tc_blocks(4096, 2, seed=9)
    assert np.array_equal(enc.transcode(ref_in, 2), O.orc_transcode(ref_in, 2)), "transcode differs from the oracle"
    for _ in range(3):
        enc.transcode_device(blocks, 2, nblocks, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        enc.transcode_device(blocks, 2, nblocks, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = nblocks * 32 / (ms * 1e-3) / 1e9
    print(json.dumps({"metric": "transcode_mblocks_per_s", "value": nblocks / ms / 1e3, "unit": "Mblocks/s", "n_gpus": 1, "steps": 10,
                      "ms_per_step": ms, "higher_is_better": True, "dtype": "u64", "data": "synthetic",
                      "config": {"workload": "s2tc_from_s3tc: 16.7 M DXT5 blocks (256 MiB, larger than L2), in place"},
                      "roofline": {"bound": "hbm", "kernel": "transcode_kernel", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s",
                                   "frac": gbs / peak_gbs, "traffic": None, "peak_source": peak_src,
                                   "algorithmic_bytes_per_block": 32}}), flush=True)


if __name__ == "__main__":
    main()
