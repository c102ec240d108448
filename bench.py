#!/usr/bin/env python
"""bench.py -- S2TC encode throughput on B200 (metric of BASELINE.json: encode Mblocks/s + roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config3|config2|defaults|config5|config4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
    python bench.py --impl reference ...        the reference's own CPU encoder on the host cores

Default workload = BASELINE.json's north-star configuration: ONE 16384x16384 RGBA texture, DXT1, WAVG,
S2TC_RANDOM_COLORS=64, S2TC_REFINE_COLORS=LOOP (config 3).  One "step" = one pass of the hot path (565 pre-pass ->
random candidates + pair search -> refinement and packing, or the fused fast kernel) over that texture, resident in
HBM.  At N > 1 the texture is STRONG-scaled: rank r owns the contiguous block rows [R r / N, R (r+1) / N) (SURVEY.md
8e); the only exchange is the DITHER_SIMPLE carry (128-byte transfer functions, one all-gather); every rank writes its
own slice of the output.  config4 (a batch of mip-mapped textures) shards whole textures instead.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput of the whole job, `e2e` goes through the
host-facing call with host<->device copies in the timed region, `roofline` is for the dominant kernel (integer pipes
for the search modes, HBM for the quick modes), `cpu_baseline` is the reference CPU encoder timed on this host.
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
    "config3": ("DXT1", "WAVG", 64, "LOOP", 16384, 16384, "synth_rgba"),
    "config2": ("DXT5", "SRGB_MIXED", 0, "LOOP", 8192, 8192, "synth_rgba"),
    "defaults": ("DXT1", "WAVG", -1, "ALWAYS", 8192, 8192, "synth_rgba"),
    "config5": ("DXT5", "NORMALMAP", -1, "NEVER", 4096, 4096, "synth_normal"),
    "config4": ("DXT3", "ALL8", -1, "ALWAYS", 2048, 2048, "synth_rgba"),   # 256 textures x full mip chain x 8 metrics
}
DXT = {"DXT1": 0, "DXT3": 1, "DXT5": 2}
CD_NAMES = ["RGB", "YUV", "SRGB", "SRGB_MIXED", "AVG", "WAVG", "W0AVG", "NORMALMAP"]
CD = {n: i for i, n in enumerate(CD_NAMES)}
REFINE = {"NEVER": 0, "ALWAYS": 1, "LOOP": 2}
DITHER = {"NONE": 0, "SIMPLE": 1, "FLOYDSTEINBERG": 2}
GL = {0: 0x83F1, 1: 0x83F2, 2: 0x83F3}
# integer operations of one distance evaluation
C_CD = {"AVG": 8, "WAVG": 8, "W0AVG": 8, "RGB": 18, "YUV": 18, "SRGB": 35, "SRGB_MIXED": 14, "NORMALMAP": 12}


def read_peaks():
    """(HBM GB/s, SM clock MHz, source)"""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


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
    if args.workload == "config4":
        cd_n = "WAVG"   # the CPU sample runs one of the eight metrics; the mip levels are separate calls of the same path
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
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.workload == "config4" else "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, wl, 1),
        "cpu_baseline": {"value": value, "unit": "Mblocks/s", "cores": threads, "kind": kind, "sample": sample,
                         "note": "565 pre-pass single-threaded as upstream, block rows over all host threads"},
        "e2e": {"value": value, "unit": "Mblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, wl, world):
    dxt_n, cd_n, nrandom, refine_n, width, height, gen = wl
    if args.workload == "config4":
        return {"workload": f"config4: batch of {args.textures} x {width}x{height} {gen} textures per GPU with full mip chains, "
                            f"{dxt_n}, all 8 S2TC_COLORDIST_MODEs, S2TC_RANDOM_COLORS={nrandom}, S2TC_REFINE_COLORS={refine_n}, "
                            f"S2TC_DITHER_MODE={args.dither}",
                "sharding": "whole textures per GPU (batch split), no exchange",
                "l2": f"{args.textures} textures x {width * height * 4 >> 20} MiB per GPU exceed the 126 MiB L2; no flush needed"}
    per_gpu = width * height * 4 // max(world, 1)
    return {"workload": f"{args.workload}: {dxt_n} {width}x{height} {gen}, S2TC_COLORDIST_MODE={cd_n}, "
                        f"S2TC_RANDOM_COLORS={nrandom}, S2TC_REFINE_COLORS={refine_n}, S2TC_DITHER_MODE={args.dither}",
            "texture": f"{width}x{height} RGBA8", "sharding": f"contiguous block rows of the one texture, {world} shard(s)" + (" (end to end: striped, see e2e.path)" if world > 1 else ""),
            "l2": (f"inputs {per_gpu >> 20} MiB per GPU exceed the 126 MiB L2; no flush needed"
                   if per_gpu > (126 << 20) else
                   f"inputs {per_gpu >> 20} MiB per GPU FIT the 126 MiB L2: a {max(per_gpu, 192 << 20) >> 20} MiB buffer is "
                   f"rewritten between timed steps to flush it")}


def gathered_counts(red, dxt):
    """Histogram of n = colours the reference gathers per block (s2tc_algorithm.cpp:940-959) from the reduced texels."""
    h, w = red.shape[:2]
    bh, bw = (h + 3) // 4, (w + 3) // 4
    valid = np.zeros((bh * 4, bw * 4), bool)
    valid[:h, :w] = True if dxt != 0 else (red[..., 3] != 0)
    n = valid.reshape(bh, 4, bw, 4).sum((1, 3)).ravel()
    n = np.maximum(n, 1)   # empty block: one black candidate (:952-959)
    return np.bincount(n, minlength=17)


def search_ops(hist, nrandom, dxt, cd_n):
    """Algorithmic integer operations per block of the pair search: one min + one add per (pair, texel)
    for the colours and again for DXT5 alpha (the two fixed points folded into the rows), plus the distance evaluations
    of the matrix fill.  Returned as (16-bit-packable ops, 32-bit ops), means over the blocks of `hist`."""
    tot = hist.sum()
    ops16 = ops32 = 0.0
    for n in range(1, 17):
        if not hist[n]:
            continue
        m = n + max(nrandom, 0)
        if nrandom <= 0 and n == 1:
            m = 2
        p = m * (m - 1) // 2
        pair = p * n * 2
        dist = (n * (n - 1) // 2 + (m - n) * n) * C_CD[cd_n]
        f = hist[n] / tot
        if cd_n in ("AVG", "WAVG", "W0AVG"):
            ops16 += f * pair
        else:
            ops32 += f * pair
        ops32 += f * dist
        if dxt == 2:
            ops16 += f * pair
            ops32 += f * (n * (n - 1) // 2 + (m - n) * n) * 3
    return ops16, ops32


def main():
    if os.environ.get("S2TC_BENCH_WATCHDOG"):   # debugging aid: dump every thread's Python stack to stderr if the run takes too long
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["S2TC_BENCH_WATCHDOG"]), repeat=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=list(WORKLOADS))
    ap.add_argument("--dither", default="SIMPLE", choices=list(DITHER))
    ap.add_argument("--size", type=int, default=0, help="override texture width=height (debug)")
    ap.add_argument("--textures", type=int, default=32, help="config4: textures per GPU")
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
        per_core = {"config2": 0.16e6, "config3": 0.03e6, "defaults": 1.9e6, "config5": 0.2e6, "config4": 1.9e6}[args.workload]
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the encoder has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.workload == "config4":
        import bench_batch
        line = bench_batch.run(args, wl, world, rank, local, dist, ClockSampler, read_peaks, workload_config, cpu_reference_run)
    else:
        line = run_texture(args, wl, world, rank, local, dist)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads (and therefore the pages its pinned buffers are first touched on) to the CPUs of the NUMA
    node its GPU hangs off, so that eight ranks do not all stage their uploads through one socket."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_texture(args, wl, world, rank, local, dist):
    import torch
    import s2tc_b200
    from s2tc_b200 import Settings

    dxt_n, cd_n, nrandom, refine_n, width, height, gen = wl
    st = Settings(DXT[dxt_n], CD[cd_n], nrandom, REFINE[refine_n], DITHER[args.dither])
    fs_chain = st.dither == 2 and world > 1   # Floyd-Steinberg shards run as a chain (sharding.floyd_steinberg_sharded): nothing scales
    bs = s2tc_b200.block_bytes(st.dxt)
    abits = {0: 1, 1: 4, 2: 8}[st.dxt]
    bw, bh = (width + 3) // 4, (height + 3) // 4
    total_blocks = bw * bh
    row0, row1 = (bh * rank) // world, (bh * (rank + 1)) // world
    my_rows = row1 - row0
    my_blocks = my_rows * bw
    y0, y1 = row0 * 4, min(row1 * 4, height)

    enc = s2tc_b200.Encoder(local)
    img = make_image(gen, width, height, 1234)        # every rank regenerates the texture and keeps its rows
    mine = np.ascontiguousarray(img[y0:y1])
    h_src = torch.from_numpy(mine).pin_memory()
    h_dst = torch.empty(max(my_blocks * bs, 1), dtype=torch.uint8).pin_memory()
    # N > 1, end to end: the texture is cut into NWAVE * world stripes of block rows, stripe w * world + rank belongs to this
    # rank (s2tc_b200_compress_host_striped): wave w is encoded while wave w + 1 is uploaded.  Contiguous shards cannot
    # overlap anything -- no shard can start before every shard above it has been uploaded and summarised.
    WAVES = [1, 2, 4, 4, 3, 2]   # relative wave sizes: a small first wave starts the kernels early, a small last one ends the tail
    NWAVE = len(WAVES)
    stripes = [s2tc_b200.Encoder.stripe_rows(height, world, NWAVE, w, rank, WAVES) for w in range(NWAVE)] if world > 1 else []
    if world > 1:
        s_tex = [np.ascontiguousarray(img[4 * a:min(4 * b, height)]) for a, b in stripes]
        hs_src = torch.from_numpy(np.concatenate([t.reshape(-1) for t in s_tex])).pin_memory()
        hs_dst = torch.empty(max(sum((b - a) * bw * bs for a, b in stripes), 1), dtype=torch.uint8).pin_memory()
        src_off = np.cumsum([0] + [t.size for t in s_tex])
        dst_off = np.cumsum([0] + [(b - a) * bw * bs for a, b in stripes])
        src_stripes = [hs_src[src_off[w]:src_off[w + 1]] if stripes[w][1] > stripes[w][0] else None for w in range(NWAVE)]
        dst_stripes = [hs_dst[dst_off[w]:dst_off[w + 1]] if stripes[w][1] > stripes[w][0] else None for w in range(NWAVE)]
        del s_tex
    d_src = h_src.cuda(non_blocking=False)
    d_dst = torch.empty(max(my_blocks * bs, 1), dtype=torch.uint8, device="cuda")
    # a non-default stream: its handle is what the C ABI launches on, and torch events recorded on it
    # bracket exactly those launches (the legacy default stream's handle is 0 = "use the context's own")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    maps_mine = torch.zeros(16 * NWAVE, dtype=torch.int64, device="cuda")
    maps_all = torch.zeros(16 * NWAVE * world, dtype=torch.int64, device="cuda")
    maps1_all = torch.zeros(16 * world, dtype=torch.int64, device="cuda")
    carry_dev = torch.zeros(4, dtype=torch.int32, device="cuda")
    per_gpu_bytes = (y1 - y0) * width * 4
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda") if per_gpu_bytes <= (126 << 20) else None

    def fs_chain_rows(d_rows, w_, h_, a, b, d_out):
        """DITHER_FLOYDSTEINBERG over the ranks' shards: error rows travel down the ranks by NCCL send / recv (colour pass,
        then the alpha seed from the last rank to rank 0, then the alpha pass), then every rank encodes its reduced texels"""
        from s2tc_b200.sharding import floyd_steinberg_sharded
        reduced = torch.empty((min(4 * b, h_) - 4 * a) * w_, dtype=torch.int32, device="cuda")
        keep = floyd_steinberg_sharded(enc, dist, d_rows, w_, h_, 4, abits, a, b, rank, world, reduced,
                                       lambda n: torch.zeros(n, dtype=torch.int32, device="cuda"), stream=stream.cuda_stream)
        enc.encode_reduced_rows_device(reduced, w_, h_, a, b, d_out, st, cursor0=0, stream=stream.cuda_stream)
        return keep, reduced

    def step_device():
        if flush is not None:
            flush.add_(1)      # rewrite a buffer larger than the L2 between steps
        if fs_chain:
            fs_chain_rows(d_src, width, height, row0, row1, d_dst)
        elif world == 1:
            enc.encode_rows_device(d_src, width, height, 4, 0, bh, d_dst, st, cursor0=0, carry=None, stream=stream.cuda_stream)
        else:   # summary -> all-gather (128 B per rank, NCCL) -> fold -> encode, all on the device, no host sync
            enc.sharded_encode_async(d_src, width, height, 4, row0, row1, d_dst, st, maps_mine[:16],
                                     lambda: dist.all_gather_into_tensor(maps1_all, maps_mine[:16]), maps1_all, rank, carry_dev,
                                     cursor0=0, stream=stream.cuda_stream)

    def step_e2e():
        if fs_chain:           # no pipeline here: the chain is the critical path
            d_src.copy_(h_src, non_blocking=True)
            fs_chain_rows(d_src, width, height, row0, row1, d_dst)
            h_dst.copy_(d_dst, non_blocking=True)
            stream.synchronize()
        elif world == 1:
            enc.compress(h_src, st, cursor=0, out=h_dst)   # the reference-facing host call, pinned buffers
        else:               # striped shards: one 128-byte all-gather per wave, uploads / kernels / downloads overlap
            enc.compress_striped(src_stripes, width, height, dst_stripes, st, rank, world, NWAVE, maps_mine, maps_all,
                                 lambda w: dist.all_gather_into_tensor(maps_all[16 * world * w:16 * world * (w + 1)],
                                                                       maps_mine[16 * w:16 * (w + 1)]),
                                 cursor0=0, stream=stream.cuda_stream, weights=WAVES)

    def row_checksum(buf, rows):
        """sum over block rows of (row number + 1) * crc32(row bytes): equal for two partitions of the same image"""
        import zlib
        tot, o = 0, 0
        for a, b in rows:
            for r in range(a, b):
                tot += (r + 1) * zlib.crc32(buf[o:o + bw * bs].tobytes())
                o += bw * bs
        return tot % (1 << 62)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_ok(ok):
        if dist is None:
            return ok
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # ---- correctness gate on EVERY rank: the first and last block rows of its shard against the oracle ---------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    checked = None
    if not args.no_check:
        import _oracle as O
        ok = True
        got_dev = d_dst[:my_blocks * bs].cpu().numpy()
        if st.dither == 2:
            # Floyd-Steinberg: the alpha pass of the reference is seeded from the LAST image row, so a
            # cropped image is not a prefix of the full one; check a smaller whole image through the same code path
            small = np.ascontiguousarray(img[:64 * world + 4, :256])
            want = O.orc_compress(small, st.dxt, st.cd, st.nrandom, st.refine, st.dither)
            if fs_chain:       # the same chain on the small image; rank 0 compares everybody's rows
                sbh = (small.shape[0] + 3) // 4
                a, b = (sbh * rank) // world, (sbh * (rank + 1)) // world
                d_small = torch.from_numpy(np.ascontiguousarray(small[4 * a:4 * b])).cuda()
                d_sout = torch.zeros((sbh // world + 1) * 64 * bs, dtype=torch.uint8, device="cuda")
                fs_chain_rows(d_small, 256, small.shape[0], a, b, d_sout)
                stream.synchronize()
                outs = [torch.zeros_like(d_sout) for _ in range(world)]
                dist.all_gather(outs, d_sout)
                got_small = np.concatenate([outs[r][:((sbh * (r + 1)) // world - (sbh * r) // world) * 64 * bs].cpu().numpy()
                                            for r in range(world)])
                ok = np.array_equal(got_small, want)
            else:
                ok = np.array_equal(enc.compress(small, st), want)
            checked = want.size // bs
        else:
            nrows = max(1, min(my_rows, (16384 if nrandom > 0 else 32768) // bw // 2))
            spans = [(row0, row0 + nrows)] if my_rows <= 2 * nrows else [(row0, row0 + nrows), (row1 - nrows, row1)]
            checked = 0
            for a, b in spans:
                want = O.orc_rows(img, st.dxt, st.cd, st.nrandom, st.refine, st.dither, (a, b), cursor=0)
                ok = ok and np.array_equal(got_dev[(a - row0) * bw * bs:(b - row0) * bw * bs], want)
                checked += (b - a) * bw
        step_e2e()     # and the host path must give the same bytes as the device path
        torch.cuda.synchronize()
        if world == 1 or fs_chain:
            ok = ok and np.array_equal(h_dst[:my_blocks * bs].numpy(), got_dev)
        else:          # the ranks own different rows in the two paths: compare checksums over all block rows of the image
            sums = torch.tensor([row_checksum(got_dev, [(row0, row1)]), row_checksum(hs_dst.numpy(), stripes)], dtype=torch.int64,
                                device="cuda")
            gathered = [torch.zeros_like(sums) for _ in range(world)]
            dist.all_gather(gathered, sums)
            tot = [sum(int(g[k].item()) for g in gathered) % (1 << 62) for k in (0, 1)]
            ok = ok and tot[0] == tot[1]
        if not all_ok(ok):
            raise SystemExit("bench.py: GPU output differs from the oracle; refusing to report a number")
        if dist is not None:
            t = torch.tensor([checked], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            checked = int(t.item())
    del img

    # ---- timed region: device-resident -------------------------------------------------------------
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
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if flush is not None:      # the flush kernel is inside the events: time it alone and take it out
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            flush.add_(1)
        f1.record(stream)
        torch.cuda.synchronize()
        ms_step -= f0.elapsed_time(f1) / args.steps
    value = total_blocks / (ms_step * 1e-3) / 1e6

    # ---- a second, untimed pass with per-family events: which kernel family dominates, and how long its launches take ----
    enc.profile(True)
    enc.profile_read(reset=True)
    for _ in range(args.steps):
        step_device()
    fam = enc.profile_read(reset=True)
    enc.profile(False)

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
    e2e_value = total_blocks / (ms_e2e * 1e-3) / 1e6

    # what the copies alone cost on this box: every rank moves its step's bytes (pinned host <-> device, both directions
    # at once, no kernels) at the same time; an end-to-end step cannot be shorter than this
    copy_only = None
    if e2e_steps:
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(s_up):
                d_src.copy_(h_src, non_blocking=True)
            with torch.cuda.stream(s_down):
                h_dst.copy_(d_dst, non_blocking=True)
        barrier()
        t_c = torch.tensor([(time.perf_counter() - t0) * 1e3 / 3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
        ms_c = float(t_c.item())
        copy_only = {"ms_per_step": ms_c, "h2d_gbs_all_gpus": width * height * 4 / (ms_c * 1e-3) / 1e9,
                     "note": "pinned H2D of every rank's texels and D2H of its blocks, all ranks at once, no kernels: the floor of "
                             "an end-to-end step on this host"}

    # the entry point a drop-in user calls, with the memory they pass: tx_compress_dxtn on malloc'd (pageable) buffers
    pageable = None
    if world == 1 and e2e_steps:
        os.environ.update({"S2TC_DITHER_MODE": args.dither, "S2TC_COLORDIST_MODE": cd_n, "S2TC_RANDOM_COLORS": str(nrandom),
                           "S2TC_REFINE_COLORS": refine_n})
        p_dst = np.zeros(my_blocks * bs, np.uint8)
        reps = max(1, min(3, args.steps))
        s2tc_b200.tx_compress_dxtn(4, width, height, mine, GL[st.dxt], p_dst, 0)
        t0 = time.perf_counter()
        for _ in range(reps):
            s2tc_b200.lib().s2tc_b200_rand_cursor_set(0)
            s2tc_b200.tx_compress_dxtn(4, width, height, mine, GL[st.dxt], p_dst, 0)
        ms_p = (time.perf_counter() - t0) * 1e3 / reps
        pageable = {"value": total_blocks / (ms_p * 1e-3) / 1e6, "ms_per_step": ms_p, "steps": reps,
                    "path": "tx_compress_dxtn itself, S2TC_* from the environment, numpy (malloc'd, pageable) src and dest",
                    "matches_pinned_path": bool(np.array_equal(p_dst, h_dst[:my_blocks * bs].numpy()))}

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peak_gbs, sm_mhz, peak_src = read_peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    dom = max(fam, key=lambda k: fam[k][0])
    dom_ms, dom_n = fam[dom]
    launch_group = {"prepass": 5 if world == 1 else 8, "candidates": 2}.get(dom, 1)   # launches timed as one group
    if dom == "search" and nrandom <= 0 and st.dxt == 2:
        launch_group = 2           # search16 runs DXT5 as a colour launch + an alpha launch
    groups = max(dom_n // launch_group, 1)
    dom_ms_launch = dom_ms / groups
    per_launch_blocks = my_blocks * args.steps / groups
    # algorithmic bytes per block the kernel must move: 64 B of texels in, the kernel's result out
    alg_bytes = {"fast": 64 + bs, "search": 64 + 8, "finish": 64 + 8 + bs, "prepass": 64 + 64, "candidates": 0.5,
                 "transcode": 2 * bs}[dom]
    hbm_achieved = alg_bytes * per_launch_blocks / (dom_ms_launch * 1e-3) / 1e9 if dom_ms_launch > 0 else 0.0
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
    hbm_view = {"achieved": hbm_achieved, "peak": peak_gbs, "unit": "GB/s", "frac": hbm_achieved / peak_gbs,
                "algorithmic_bytes_per_block": alg_bytes, "peak_source": peak_src}
    kernel_ms = {k: v[0] / args.steps for k, v in fam.items() if v[1]}
    search_mode = nrandom >= 0 or cd_n == "NORMALMAP"
    if search_mode and fam["search"][1]:
        # Integer roofline.  Operations: what the reference's algorithm performs on THIS texture (the number of
        # gathered colours per block comes from the reduced texels of rank 0's shard).  Peak: the issue ceiling of the SMs
        # -- one warp instruction per scheduler and clock; a 32-bit min or add is one operation per lane, a packed 16-bit one
        # is two -- at the SM clock of MEASURED_PEAKS.json; the rates the kernels' own instruction mixes sustain in a
        # micro-benchmark of this run are reported beside it.
        # Per-block figure = SURVEY.md 8(d)'s: every block has n = 16 colours, m = 16 + nrandom candidates.  What the reference
        # would really execute on THIS texture (fewer colours in partly transparent DXT1 blocks) is reported beside it.
        nom16, nom32 = search_ops(np.bincount([16], minlength=17), nrandom, st.dxt, cd_n)
        red = enc.rgb565_image(mine, abits, st.dither) if st.dxt == 0 else None
        hist = gathered_counts(red if red is not None else mine, st.dxt)
        act16, act32 = search_ops(hist, nrandom, st.dxt, cd_n)
        r32 = sms * 128 * sm_mhz * 1e6 / 1e9           # Gop/s, scalar
        r16 = 2 * r32                                   # two 16-bit lanes per register
        ops = nom16 + nom32
        peak = ops / (nom32 / r32 + nom16 / r16)
        s_launch_group = 2 if (nrandom <= 0 and st.dxt == 2) else 1
        s_groups = max(fam["search"][1] // s_launch_group, 1)
        s_ms = fam["search"][0] / s_groups
        s_blocks = my_blocks * args.steps / s_groups
        achieved = ops * s_blocks / (s_ms * 1e-3) / 1e9
        act = act16 + act32
        act_peak = act / (act32 / r32 + act16 / r16)
        act_achieved = act * s_blocks / (s_ms * 1e-3) / 1e9
        peak_scalar, peak_packed = enc.int_peaks_gops()
        roofline = {"bound": "int32", "kernel": "pair_search_kernel" if nrandom > 0 else "search16_kernel",
                    "achieved": achieved, "peak": peak, "unit": "Gop/s", "frac": achieved / peak,
                    "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                    "ms_per_launch": s_ms, "blocks_per_launch": s_blocks,
                    "ops_per_block": ops, "ops_per_block_16bit_packed": nom16, "ops_per_block_32bit": nom32,
                    "ops_definition": "SURVEY.md 8(d): 2 per (pair, texel) over P = m(m-1)/2 pairs and n = 16 texels for colours (+ the "
                                      "same for DXT5 alpha, fixed points folded) + D distance evaluations x C_cd of the matrix fill",
                    "on_this_texture": {"ops_per_block": act, "achieved": act_achieved, "peak": act_peak, "frac": act_achieved / act_peak,
                                        "note": "the reference's operation count with the n it really gathers per block (DXT1 skips "
                                                "transparent texels; single-colour blocks cost it the same pairs but n columns)"},
                    "note": "algorithmic work per second, not pipe occupancy: the pruned scan never evaluates pairs whose lower bound "
                            "exceeds the best sum, and answers single-colour blocks at once" if nrandom > 0 else None,
                    "peak_source": f"issue ceiling: {sms} SMs x 128 lanes x {sm_mhz:.0f} MHz (x2 for packed 16-bit operands)",
                    "measured_mix_rates": {"scalar_gops": peak_scalar, "packed16_gops": peak_packed,
                                           "how": "s2tc_b200_int_peaks in this run: VIMNMX + IMAD; VIMNMX.U16x2 + IDP.2A"},
                    "kernel_ms_per_step": kernel_ms, "hbm": hbm_view}
    else:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": hbm_achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": hbm_achieved / peak_gbs, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                    "traffic_detail": traffic, "peak_source": peak_src, "algorithmic_bytes_per_block": alg_bytes,
                    "ms_per_launch": dom_ms_launch, "kernel_ms_per_step": kernel_ms,
                    "whole_step": {"achieved": (64 + bs) * my_blocks / (ms_step * 1e-3) / 1e9,
                                   "frac": (64 + bs) * my_blocks / (ms_step * 1e-3) / 1e9 / peak_gbs,
                                   "note": "algorithmic bytes of the whole step (texels in, blocks out) over the step time"}}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.kernel_only:
        threads = os.cpu_count() or 1
        rows = args.cpu_rows or max(4, min(bh, int(args.cpu_blocks // bw)))
        kind, tp, tb, out = cpu_reference_run(mine, st.dxt, st.cd, nrandom, st.refine, st.dither, rows, threads)
        same = bool(np.array_equal(out[:rows * bw * bs], d_dst[:rows * bw * bs].cpu().numpy()))
        cpu = {"value": rows * bw / (tp + tb) / 1e6, "unit": "Mblocks/s", "cores": threads, "kind": kind,
               "sample": f"first {rows} of {bh} block rows ({rows * bw} blocks): pre-pass {tp:.2f} s on 1 thread + "
                         f"blocks {tb:.2f} s on {threads} threads",
               "matches_gpu_output": same}

    return {
        "metric": "encode_mblocks_per_s", "value": value, "unit": "Mblocks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": workload_config(args, wl, world),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "Mblocks/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": width * height * 4,
                "d2h_bytes_per_step": total_blocks * bs,
                "path": "s2tc_b200_compress_host (what tx_compress_dxtn calls), pinned host buffers" if world == 1
                else "upload, Floyd-Steinberg chain over the ranks (s2tc_b200_floyd_rows_device, NCCL send / recv of the error rows), "
                     "encode, download: the chain is the critical path, nothing overlaps" if fs_chain
                else f"s2tc_b200_compress_host_striped per rank: {NWAVE} waves (sizes {WAVES}) x {world} stripes of block rows, stripe w * world + rank "
                     "on rank `rank`; wave w is encoded while wave w + 1 is uploaded; one 128-byte all-gather of DITHER_SIMPLE "
                     "summaries per wave (NCCL); pinned host buffers",
                "copy_only": copy_only, "pageable": pageable},
        "gpu_launches": launches, "clocks": clocks,
        "checked_blocks_vs_oracle": checked,
    }


if __name__ == "__main__":
    main()
