#!/usr/bin/env python
"""bench_extra.py -- the secondary BASELINE.json workloads that do not fit bench.py's one-texture step:

  config4    DXT3, full mip chains of a batch of 2048x2048 textures, all 8 ColorDistModes, REFINE=ALWAYS
             (32 textures per GPU = BASELINE's 256 over 8 GPUs); every chain runs on the device
             (s2tc_b200_compress_mipchain_device: encode level, halve, repeat down to 1x1)
  transcode  s2tc_from_s3tc on 16.7 M DXT5 blocks resident in HBM (pure streaming: 32 B of traffic per block)

Prints one JSON line per workload (same vocabulary as bench.py).  Single GPU; a multi-GPU run shards the batch
by textures, there is no cross-texture state except the rand() cursor (closed form, unused here: nrandom = -1).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="all", choices=["all", "config4", "transcode"])
    ap.add_argument("--textures", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--lanes", type=int, default=8, help="chains in flight (streams / encoder contexts)")
    args = ap.parse_args()
    import torch
    import _oracle as O
    import s2tc_b200
    from s2tc_b200 import Settings, synth
    from bench import read_peaks

    torch.cuda.set_device(0)
    enc = s2tc_b200.Encoder(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    peak_gbs, peak_src = read_peaks()

    if args.workload in ("all", "config4"):
        size, ntex = 2048, args.textures
        variants = [torch.from_numpy(synth.synth_rgba(size, size, seed=100 + i)).cuda() for i in range(4)]
        chain_bytes = s2tc_b200.lib().s2tc_b200_mipchain_bytes(s2tc_b200.DXT3, size, size)
        blocks_per_chain = chain_bytes // 16
        # parity gate: one chain against the oracle, level by level
        st = Settings(s2tc_b200.DXT3, s2tc_b200.WAVG, -1, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_SIMPLE)
        small = synth.synth_rgba(64, 64, seed=3)
        from test_oracle import orc_mip_reduce
        want, level = [], small
        while True:
            want.append(O.orc_compress(level, O.DXT3, O.WAVG, -1, O.ALWAYS, O.DITHER_SIMPLE))
            if level.shape[0] == 1:
                break
            level = orc_mip_reduce(level)
        assert np.array_equal(enc.compress_mipchain(small, st), np.concatenate(want)), "mip chain differs from the oracle"

        # A chain is ~70 small DEPENDENT launches (12 levels x pre-pass / encode / halve), i.e. latency-bound on its
        # own: textures are independent, so `lanes` chains run concurrently, each on its own stream with its own
        # encoder context (workspaces), and each (lane, mode) chain is captured once into a CUDA graph and replayed.
        class Lane:
            def __init__(self):
                self.enc = s2tc_b200.Encoder(0)
                self.stream = torch.cuda.Stream()
                self.staging = torch.empty_like(variants[0])
                self.work = torch.empty_like(variants[0])
                self.scratch = torch.empty(size * size, dtype=torch.uint8, device="cuda")
                self.dst = torch.empty(chain_bytes, dtype=torch.uint8, device="cuda")
                self.graphs = []

            def chain(self, cd):
                s = Settings(s2tc_b200.DXT3, cd, -1, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_SIMPLE)
                self.work.copy_(self.staging, non_blocking=True)   # the chain overwrites its source
                self.enc.compress_mipchain_device(self.work, self.scratch, self.dst, size, size, s, stream=self.stream.cuda_stream)

        lanes = [Lane() for _ in range(max(1, args.lanes))]
        for ln in lanes:
            with torch.cuda.stream(ln.stream):
                for cd in range(8):
                    ln.chain(cd)                         # warm-up: workspaces are allocated outside the capture
            ln.stream.synchronize()
            if not args.no_graphs:
                for cd in range(8):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=ln.stream):
                        ln.chain(cd)
                    ln.graphs.append(g)
        graphs = not args.no_graphs

        def step():
            for cd in range(8):
                for t in range(ntex):
                    ln = lanes[t % len(lanes)]
                    with torch.cuda.stream(ln.stream):
                        ln.staging.copy_(variants[t % len(variants)], non_blocking=True)
                        if graphs:
                            ln.graphs[cd].replay()
                        else:
                            ln.chain(cd)

        def sync_lanes():
            for ln in lanes:
                stream.wait_stream(ln.stream)

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for ln in lanes:
            ln.stream.wait_stream(stream)
        for _ in range(args.steps):
            step()
        sync_lanes()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        blocks = blocks_per_chain * ntex * 8
        print(json.dumps({"metric": "encode_mblocks_per_s", "value": blocks / ms / 1e3, "unit": "Mblocks/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "dtype": "int32", "data": "synthetic",
                          "config": {"workload": f"config4: DXT3, {ntex} textures 2048x2048 with full mip chains (12 levels, {blocks_per_chain} blocks each), "
                                                 "all 8 S2TC_COLORDIST_MODEs, S2TC_RANDOM_COLORS=-1, S2TC_REFINE_COLORS=ALWAYS, S2TC_DITHER_MODE=SIMPLE",
                                     "note": f"{len(lanes)} chains in flight on {len(lanes)} streams" + (", each chain replayed as a CUDA graph" if graphs else "")},
                          "gpu_launches": 75 * 8 * ntex * args.steps}), flush=True)

    if args.workload in ("all", "transcode"):
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
