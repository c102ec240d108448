"""bench.py --workload config4: BASELINE.json's "DXT3 full mip chain batch of 256 2048x2048 textures, all 8 ColorDistModes,
REFINE=ALWAYS, batch sharded over 8 B200" -- 32 textures per GPU (weak scaling: the batch grows with the GPUs), whole
textures per rank, no exchange.  One step = the full chains of this rank's textures under all 8 metrics through
s2tc_b200_compress_mipchain_batch_device (every mip level of all textures per launch, pre-pass shared by the metrics).
The reference does this with one s2tc_compress process per texture and metric (s2tc_compress.c:722-733).
"""
import os
import time

import numpy as np


def run(args, wl, world, rank, local, dist, ClockSampler, read_peaks, workload_config, cpu_reference_run):
    import torch
    import _oracle as O
    import s2tc_b200
    from s2tc_b200 import Settings, synth
    from test_oracle import orc_mip_reduce

    dxt_n, _, nrandom, refine_n, width, height, gen = wl
    dither = {"NONE": 0, "SIMPLE": 1, "FLOYDSTEINBERG": 2}[args.dither]
    refine = {"NEVER": 0, "ALWAYS": 1, "LOOP": 2}[refine_n]
    dxt = {"DXT1": 0, "DXT3": 1, "DXT5": 2}[dxt_n]
    sets = [Settings(dxt, cd, nrandom, refine, dither) for cd in range(8)]
    ntex = args.textures
    bs = s2tc_b200.block_bytes(dxt)
    enc = s2tc_b200.Encoder(local)
    chain = s2tc_b200.lib().s2tc_b200_mipchain_bytes(dxt, width, height)
    blocks_chain = chain // bs
    blocks_step = blocks_chain * ntex * len(sets)          # per GPU
    set_stride = (chain * ntex + 15) & ~15
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    # this rank's textures: four generated variants repeated (host generation time, not GPU work, limits more)
    variants = [getattr(synth, gen)(width, height, seed=1000 + 16 * rank + i) for i in range(4)]
    h_src = torch.empty((ntex, height, width, 4), dtype=torch.uint8).pin_memory()
    for i in range(ntex):
        h_src[i] = torch.from_numpy(variants[i % 4])
    d_src = h_src.cuda()
    scratch = torch.empty(ntex * (width * height + width * height // 4) + 4096, dtype=torch.uint8, device="cuda")
    d_dst = torch.empty(set_stride * len(sets), dtype=torch.uint8, device="cuda")
    h_dst = torch.empty(set_stride * len(sets), dtype=torch.uint8).pin_memory()

    def step_device():
        enc.compress_mipchain_batch_device(d_src, scratch, d_dst, width, height, ntex, sets, stream=stream.cuda_stream)

    # end to end: sub-batches of 8 textures; upload of the next, kernels of this and download of the previous overlap
    SUB = 8 if ntex % 8 == 0 else ntex
    nsub = ntex // SUB
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    sub_stride = (chain * SUB + 15) & ~15
    d_in = [torch.empty((SUB, height, width, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_out = [torch.empty(sub_stride * len(sets), dtype=torch.uint8, device="cuda") for _ in range(2)]
    h_out = torch.empty((nsub, sub_stride * len(sets)), dtype=torch.uint8).pin_memory()
    sub_scratch = torch.empty(SUB * (width * height + width * height // 4) + 4096, dtype=torch.uint8, device="cuda")

    def step_e2e():
        ev_up = [torch.cuda.Event() for _ in range(nsub)]
        ev_done = [torch.cuda.Event() for _ in range(nsub)]
        ev_down = [torch.cuda.Event() for _ in range(nsub)]
        for i in range(nsub):
            b = i & 1
            with torch.cuda.stream(up):
                if i >= 2:
                    up.wait_event(ev_done[i - 2])        # the kernels that read this input buffer are done
                d_in[b].copy_(h_src[i * SUB:(i + 1) * SUB], non_blocking=True)
                ev_up[i].record(up)
            stream.wait_event(ev_up[i])
            if i >= 2:
                stream.wait_event(ev_down[i - 2])        # the download that reads this output buffer is done
            enc.compress_mipchain_batch_device(d_in[b], sub_scratch, d_out[b], width, height, SUB, sets, stream=stream.cuda_stream)
            ev_done[i].record(stream)
            with torch.cuda.stream(down):
                down.wait_event(ev_done[i])
                h_out[i].copy_(d_out[b], non_blocking=True)
                ev_down[i].record(down)
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate (every rank): small batch against the oracle; full size: the levels from 128x128 down of
    # texture 0 under every metric against the oracle, and the e2e path against the device path
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    ok, checked = True, 0
    if not args.no_check:
        def orc_chain(img, st):
            out, level = [], img
            while True:
                out.append(O.orc_compress(level, st.dxt, st.cd, st.nrandom, st.refine, st.dither))
                if level.shape[0] == 1 and level.shape[1] == 1:
                    return np.concatenate(out)
                level = orc_mip_reduce(level)
        level, off = variants[0], 0
        while level.shape[0] > 128 or level.shape[1] > 128:
            off += ((level.shape[1] + 3) // 4) * ((level.shape[0] + 3) // 4) * bs
            level = orc_mip_reduce(level)
        got = d_dst.cpu().numpy()
        for k, st in enumerate(sets):
            want = orc_chain(level, st)
            ok = ok and np.array_equal(got[k * set_stride + off:k * set_stride + chain], want)
            checked += want.size // bs
        step_e2e()
        for i in range(nsub):      # sub-batch i, setting k == textures [i*SUB, (i+1)*SUB) of setting k in the one-shot result
            for k in range(len(sets)):
                a = h_out[i, k * sub_stride:k * sub_stride + chain * SUB].numpy()
                b = got[k * set_stride + i * SUB * chain:k * set_stride + (i + 1) * SUB * chain]
                ok = ok and np.array_equal(a, b)
        if dist is not None:
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t.item())
        if not ok:
            raise SystemExit("bench.py: GPU output differs from the oracle; refusing to report a number")

    # ---- timed region: device-resident
    launches0 = enc.launch_count()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    launches = enc.launch_count() - launches0
    value = blocks_step * world / (ms_step * 1e-3) / 1e6

    enc.profile(True)
    enc.profile_read(reset=True)
    for _ in range(args.steps):
        step_device()
    fam = enc.profile_read(reset=True)
    enc.profile(False)

    e2e_steps = 0 if args.kernel_only else args.steps
    for _ in range(1 if e2e_steps else 0):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 if e2e_steps else float("nan")], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item()) / args.steps
    copy_only = None
    if e2e_steps:      # the copies alone, all ranks at once: the floor of an end-to-end step on this host
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(up):
                d_src.copy_(h_src, non_blocking=True)
            with torch.cuda.stream(down):
                h_dst.copy_(d_dst, non_blocking=True)
        barrier()
        t_c = torch.tensor([(time.perf_counter() - t0) * 1e3 / 2], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
        copy_only = {"ms_per_step": float(t_c.item()),
                     "note": "pinned H2D of the textures and D2H of the chains of every rank at once, no kernels"}
    if rank != 0:
        return None

    peak_gbs, _, peak_src = read_peaks()
    kernel_ms = {k: v[0] / args.steps for k, v in fam.items() if v[1]}
    dom = max(fam, key=lambda k: fam[k][0])
    alg = (64 + bs) * blocks_step          # every (texture, level, metric) is one tx_compress_dxtn call: texels in, blocks out
    achieved = alg / (ms_step * 1e-3) / 1e9
    # the fast-encode launches alone: 7 of the 8 metrics, 64 B + 16 B per block each
    fast_ms = fam["fast"][0] / args.steps
    fast_blocks = blocks_chain * ntex * 7
    roofline = {"bound": "hbm", "kernel": "fast_encode_kernel (7 metrics) + search16/finish (NORMALMAP) + shared pre-pass",
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_block": 64 + bs,
                "note": "whole step: algorithmic bytes of all calls over the step time; the 565 pre-pass and the mip reduction are "
                        "shared by the 8 metrics, so the step moves fewer bytes than 8 separate runs would",
                "dominant_family": dom, "kernel_ms_per_step": kernel_ms,
                "fast_encode_only": {"achieved": (64 + bs) * fast_blocks / (fast_ms * 1e-3) / 1e9 if fast_ms > 0 else None,
                                     "frac": (64 + bs) * fast_blocks / (fast_ms * 1e-3) / 1e9 / peak_gbs if fast_ms > 0 else None}}
    cpu = None
    if world == 1 and not args.kernel_only:
        threads = os.cpu_count() or 1
        bw = (width + 3) // 4
        rows = min((height + 3) // 4, max(4, int(args.cpu_blocks // bw)))
        kind, tp, tb, _ = cpu_reference_run(variants[0], dxt, 5, nrandom, refine, dither, rows, threads)
        cpu = {"value": rows * bw / (tp + tb) / 1e6, "unit": "Mblocks/s", "cores": threads, "kind": kind,
               "sample": f"level 0 of one texture, S2TC_COLORDIST_MODE=WAVG, first {rows} block rows ({rows * bw} blocks): pre-pass "
                         f"{tp:.2f} s on 1 thread + blocks {tb:.2f} s on {threads} threads"}
    return {
        "metric": "encode_mblocks_per_s", "value": value, "unit": "Mblocks/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": workload_config(args, wl, world),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": blocks_step * world / (ms_e2e * 1e-3) / 1e6, "unit": "Mblocks/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": ntex * width * height * 4, "d2h_bytes_per_step": chain * ntex * len(sets),
                "path": f"pinned host textures -> sub-batches of {SUB} (upload / batch encode / download overlapped) -> pinned host chains",
                "copy_only": copy_only},
        "gpu_launches": launches, "launches_per_step": launches / args.steps, "clocks": clocks,
        "blocks_per_step_per_gpu": blocks_step, "checked_blocks_vs_oracle": checked,
    }
