#!/usr/bin/env python
"""bench.py -- S2TC encode throughput on B200 (metric of BASELINE.json: encode Mblocks/s + roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|defaults|config5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
    python bench.py --impl reference ...        the reference's own CPU encoder on the host cores

One "step" = one pass of the hot path (565 pre-pass -> [random candidates] -> pair search -> refinement
and packing, or the fused fast kernel) over one texture resident in HBM.  At N > 1 every rank owns a
contiguous range of block rows of one tall texture (weak scaling: 8192-row shard per GPU); the only
exchange is the DITHER_SIMPLE carry (a 128-byte transfer function per rank).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the
reference-facing host call with host<->device copies in the timed region, `roofline` is for the
dominant kernel, `cpu_baseline` is the reference CPU encoder timed on this host.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# name -> (dxt, cd, nrandom, refine, width, height, generator)
WORKLOADS = {
    "config2": ("DXT5", "SRGB_MIXED", 0, "LOOP", 8192, 8192, "synth_rgba"),
    "config3": ("DXT1", "WAVG", 64, "LOOP", 16384, 16384, "synth_rgba"),
    "defaults": ("DXT1", "WAVG", -1, "ALWAYS", 8192, 8192, "synth_rgba"),
    "config5": ("DXT5", "NORMALMAP", -1, "NEVER", 4096, 4096, "synth_normal"),
}
DXT = {"DXT1": 0, "DXT3": 1, "DXT5": 2}
CD = {n: i for i, n in enumerate(["RGB", "YUV", "SRGB", "SRGB_MIXED", "AVG", "WAVG", "W0AVG", "NORMALMAP"])}
REFINE = {"NEVER": 0, "ALWAYS": 1, "LOOP": 2}
DITHER = {"NONE": 0, "SIMPLE": 1, "FLOYDSTEINBERG": 2}
GL = {0: 0x83F1, 1: 0x83F2, 2: 0x83F3}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(float(r[0]) for r in rows), "sm_max_mhz": float(rows[0][1]),
                "reasons": reasons, "samples": len(rows)}


def make_image(gen, width, height, seed):
    from s2tc_b200 import synth
    return getattr(synth, gen)(width, height, seed=seed)


def cpu_reference_run(img, dxt, cd, nrandom, refine, dither, rows, threads):
    """The reference's CPU encoder (oracle/_ref, else the oracle port) over the first `rows` block rows."""
    import ctypes as C
    import _oracle as O
    h, w, comps = img.shape
    kind = "reference" if O.ref_available() else "port"
    handle = O.ref_handle(True) if kind == "reference" else None
    out = np.zeros(rows * ((w + 3) // 4) * O.block_bytes(dxt) + 64, np.uint8)
    t = (C.c_double * 2)()
    sub = np.ascontiguousarray(img[:rows * 4])
    rc = O.lib().refh_encode_mt(handle, comps, w, sub.shape[0], sub.ctypes.data_as(C.POINTER(C.c_ubyte)), GL[dxt], dither, cd,
                                nrandom, refine, 0, out.ctypes.data_as(C.POINTER(C.c_ubyte)), 0, 0, rows, threads, t)
    assert rc == 0, rc
    return kind, t[0], t[1], out


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dxt_n, cd_n, nrandom, refine_n, width, height, gen = wl
    dxt, cd, refine, dither = DXT[dxt_n], CD[cd_n], REFINE[refine_n], DITHER[args.dither]
    threads = os.cpu_count() or 1
    bw = (width + 3) // 4
    # bounded sample: block rows sized for a few seconds per step on this host
    rows = args.cpu_rows or max(4, min((height + 3) // 4, int(args.cpu_blocks // bw)))
    img = make_image(gen, width, rows * 4, 1234)
    times = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        kind, tp, tb, _ = cpu_reference_run(img, dxt, cd, nrandom, refine, dither, rows, threads)
        if i >= args.warmup:
            times.append(tp + tb)
    blocks = rows * bw
    ms = 1e3 * sum(times) / len(times)
    value = blocks / (ms * 1e-3) / 1e6
    sample = f"first {rows} of {(height + 3) // 4} block rows ({blocks} blocks) of the {width}x{height} texture per step"
    line = {
        "impl": "reference", "metric": "encode_mblocks_per_s", "value": value, "unit": "Mblocks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, wl),
        "cpu_baseline": {"value": value, "unit": "Mblocks/s", "cores": threads, "kind": kind, "sample": sample,
                         "note": "565 pre-pass single-threaded as upstream, block rows over all host threads"},
        "e2e": {"value": value, "unit": "Mblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, wl):
    dxt_n, cd_n, nrandom, refine_n, width, height, gen = wl
    return {"workload": f"{args.workload}: {dxt_n} {width}x{height} {gen}, S2TC_COLORDIST_MODE={cd_n}, "
                        f"S2TC_RANDOM_COLORS={nrandom}, S2TC_REFINE_COLORS={refine_n}, S2TC_DITHER_MODE={args.dither}",
            "per_gpu_texture": f"{width}x{height} RGBA8", "sharding": "block rows of one tall texture, one shard per GPU",
            "l2": (f"inputs {width * height * 4 >> 20} MiB per GPU exceed the 126 MiB L2; no flush needed"
                   if width * height * 4 > (126 << 20) else
                   f"inputs {width * height * 4 >> 20} MiB per GPU FIT the 126 MiB L2 (a --size/--workload choice for "
                   f"debugging or profiling, not a bench configuration)")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=list(WORKLOADS))
    ap.add_argument("--dither", default="SIMPLE", choices=list(DITHER))
    ap.add_argument("--size", type=int, default=0, help="override texture width=height (debug)")
    ap.add_argument("--cpu-blocks", type=float, default=0, help="blocks per CPU sample step (0 = auto)")
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and CPU legs (for runs under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = list(WORKLOADS[args.workload])
    if args.size:
        wl[4] = wl[5] = args.size
    wl = tuple(wl)
    if not args.cpu_blocks:
        # ~10-30 s of single-core work spread over the host threads (reference speeds from BASELINE.md)
        per_core = {"config2": 0.16e6, "config3": 0.03e6, "defaults": 1.9e6, "config5": 0.2e6}[args.workload]
        args.cpu_blocks = per_core * 16

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    # stdout carries exactly one JSON line: anything libraries print while initialising (NCCL's version banner,
    # for one) is sent to stderr instead
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import s2tc_b200
    from s2tc_b200 import Settings

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the encoder has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dxt_n, cd_n, nrandom, refine_n, width, height, gen = wl
    st = Settings(DXT[dxt_n], CD[cd_n], nrandom, REFINE[refine_n], DITHER[args.dither])
    bs = s2tc_b200.block_bytes(st.dxt)
    abits = {0: 1, 1: 4, 2: 8}[st.dxt]
    bw, bh = (width + 3) // 4, (height + 3) // 4
    blocks = bw * bh
    total_h = height * world
    row0, row1 = rank * bh, (rank + 1) * bh
    dpb = s2tc_b200.draws_per_block(st.dxt, nrandom)

    enc = s2tc_b200.Encoder(local)
    img = make_image(gen, width, height, 1234 + rank)
    h_src = torch.from_numpy(img).pin_memory()
    h_dst = torch.empty(blocks * bs, dtype=torch.uint8).pin_memory()
    d_src = h_src.cuda(non_blocking=False)
    d_dst = torch.empty(blocks * bs, dtype=torch.uint8, device="cuda")
    # a non-default stream: its handle is what the C ABI launches on, and torch events recorded on it
    # bracket exactly those launches (the legacy default stream's handle is 0 = "use the context's own")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    
    def incoming_carry():
        """DITHER_SIMPLE across shards: all-gather the 128-byte transfer functions, fold the lower ranks."""
        if world == 1 or st.dither != 1:
            return None
        maps = enc.dither_summary_device(d_src, width, total_h, 4, abits, row0, row1, stream=stream.cuda_stream)
        mine = torch.tensor([m - (1 << 64) if m >= (1 << 63) else m for m in maps], dtype=torch.int64, device="cuda")
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        carry = [0, 0, 0, 0]
        for r in range(rank):
            m = [int(x) & ((1 << 64) - 1) for x in gathered[r].tolist()]
            carry = enc.carry_apply(m, 4, abits, carry)
        return carry

    maps_mine = torch.zeros(16, dtype=torch.int64, device="cuda")
    maps_all = torch.zeros(16 * world, dtype=torch.int64, device="cuda")
    carry_dev = torch.zeros(4, dtype=torch.int32, device="cuda")

    def step_device():
        if world == 1:
            enc.encode_rows_device(d_src, width, total_h, 4, row0, row1, d_dst, st, cursor0=0, carry=None,
                                   stream=stream.cuda_stream)
        else:   # summary -> all-gather (128 B per rank, NCCL) -> fold -> encode, all on the device, no host sync
            enc.sharded_encode_async(d_src, width, total_h, 4, row0, row1, d_dst, st, maps_mine,
                                     lambda: dist.all_gather_into_tensor(maps_all, maps_mine), maps_all, rank, carry_dev,
                                     cursor0=0, stream=stream.cuda_stream)

    def step_e2e():
        if world == 1:
            enc.compress(h_src, st, cursor=0, out=h_dst)   # the reference-facing host call, pinned buffers
        else:
            d_src.copy_(h_src, non_blocking=True)
            step_device()
            h_dst.copy_(d_dst, non_blocking=True)
            torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate on rank 0: the first block rows against the oracle ----------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    checked = None
    if not args.no_check and rank == 0:
        import _oracle as O
        rows = max(1, min(bh, 32768 // bw))
        if st.dither == 2:
            # Floyd-Steinberg: the alpha pass of the reference is seeded from the LAST image row, so a
            # cropped image is not a prefix of the full one; check a smaller whole image through the same code path
            small = np.ascontiguousarray(img[:256, :256])
            got_small = enc.compress(small, st)
            if not np.array_equal(got_small, O.orc_compress(small, st.dxt, st.cd, st.nrandom, st.refine, st.dither)):
                raise SystemExit("bench.py: GPU output differs from the oracle; refusing to report a number")
            rows = 0
        want = O.orc_compress(img[:rows * 4], st.dxt, st.cd, st.nrandom, st.refine, st.dither) if rows else np.zeros(0, np.uint8)
        got = d_dst[:rows * bw * bs].cpu().numpy()
        if not np.array_equal(got, want):
            raise SystemExit("bench.py: GPU output differs from the oracle; refusing to report a number")
        checked = rows * bw

    # ---- timed region: device-resident -------------------------------------------------------------
    enc.profile(True)
    enc.profile_read(reset=True)
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
    ms_total = ev0.elapsed_time(ev1)
    launches = enc.launch_count() - launches0
    fam = enc.profile_read(reset=True)
    enc.profile(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = blocks * world / (ms_step * 1e-3) / 1e6

    # ---- timed region: end to end through the host-facing call -------------------------------------
    e2e_steps = 0 if args.kernel_only else args.steps
    for _ in range(2 if e2e_steps else 0):
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
    e2e_value = blocks * world / (ms_e2e * 1e-3) / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peak_gbs, peak_src = read_peaks()
    dom = max(fam, key=lambda k: fam[k][0])
    dom_ms, dom_n = fam[dom]
    if dom == "prepass":
        dom_n //= 5   # maps, three scan launches and the replay are timed as one group
    search_group = 2 if (nrandom <= 0 and st.dxt == 2) else 1   # search16 runs DXT5 as a colour launch + an alpha launch
    if dom == "search":
        dom_n //= search_group
    dom_ms_launch = dom_ms / max(dom_n, 1)
    per_launch_blocks = blocks * args.steps / max(dom_n, 1)
    # algorithmic bytes per block the kernel must move: 64 B of texels in, the kernel's result out
    alg_bytes = {"fast": 64 + bs, "search": 64 + 8, "finish": 64 + 8 + bs, "prepass": 64 + 64, "candidates": 64 + nrandom * 2,
                 "transcode": 2 * bs}[dom]
    achieved = alg_bytes * per_launch_blocks / (dom_ms_launch * 1e-3) / 1e9 if dom_ms_launch > 0 else 0.0
    # integer roofline of the pair search: one min + one add per (pair, texel), for colours and -- DXT5 --
    # once more for alpha.  Rows whose distances fit 16 bits (alpha; the AVG-family metrics) are scanned two texels per
    # instruction, so their share is rated against the packed peak: peak = ops / (ops32 / R_scalar + ops16 / R_packed).
    peak_scalar, peak_packed = enc.int_peaks_gops()
    n_pairs = (16 + max(nrandom, 0)) * (15 + max(nrandom, 0)) // 2
    colour_ops = n_pairs * 16 * 2
    colour16 = cd_n in ("AVG", "WAVG", "W0AVG")
    ops32 = 0 if colour16 else colour_ops
    ops16 = (colour_ops if colour16 else 0) + (colour_ops if st.dxt == 2 else 0)
    int_ops_block = ops32 + ops16
    search_n = fam["search"][1] // search_group
    search_ms = fam["search"][0] / max(search_n, 1)
    search_blocks = blocks * args.steps / max(search_n, 1)
    int32 = None
    if fam["search"][1] and peak_scalar and peak_packed:
        a = int_ops_block * search_blocks / (search_ms * 1e-3) / 1e9
        peak = int_ops_block / (ops32 / peak_scalar + ops16 / peak_packed)
        int32 = {"kernel": "pair_search_kernel" if nrandom > 0 else "search16_kernel", "achieved": a, "peak": peak,
                 "unit": "Gop/s (integer min+add)", "frac": a / peak,
                 "ops_per_block": int_ops_block, "ops_per_block_32bit": ops32, "ops_per_block_16bit_packed": ops16,
                 "peak_scalar": peak_scalar, "peak_packed16": peak_packed,
                 "peak_source": "measured in this run (s2tc_b200_int_peaks: VIMNMX + IMAD on two pipes; VIMNMX.U16x2 + IDP.2A)"}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f).get(args.workload, {}).get(dom)
        if rec and not args.size:
            per_block = rec["dram_bytes"] / rec["blocks_per_launch"]
            traffic = {"dram_bytes_per_launch": per_block * per_launch_blocks, "bytes_per_block": per_block,
                       "captured_blocks_per_launch": rec["blocks_per_launch"], "source": rec["source"]}
    except OSError:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_block": alg_bytes, "ms_per_launch": dom_ms_launch,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in fam.items() if v[1]},
                "int32": int32}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.kernel_only:
        threads = os.cpu_count() or 1
        rows = args.cpu_rows or max(4, min(bh, int(args.cpu_blocks // bw)))
        kind, tp, tb, out = cpu_reference_run(img, st.dxt, st.cd, nrandom, st.refine, st.dither, rows, threads)
        same = bool(np.array_equal(out[:rows * bw * bs], d_dst[:rows * bw * bs].cpu().numpy()))
        cpu = {"value": rows * bw / (tp + tb) / 1e6, "unit": "Mblocks/s", "cores": threads, "kind": kind,
               "sample": f"first {rows} of {bh} block rows ({rows * bw} blocks): pre-pass {tp:.2f} s on 1 thread + "
                         f"blocks {tb:.2f} s on {threads} threads",
               "matches_gpu_output": same}

    line = {
        "metric": "encode_mblocks_per_s", "value": value, "unit": "Mblocks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": workload_config(args, wl),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "Mblocks/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": width * height * 4,
                "d2h_bytes_per_step": blocks * bs,
                "path": "s2tc_b200_compress_host (what tx_compress_dxtn calls), pinned host buffers" if world == 1
                else "pinned H2D + s2tc_b200_encode_rows_device + D2H per rank"},
        "gpu_launches": launches, "clocks": clocks,
        "checked_blocks_vs_oracle": checked,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
