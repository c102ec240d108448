"""Workload for compute-sanitizer (memcheck / racecheck): every kernel family on small ragged images, checked against the oracle."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import s2tc_b200, _oracle as O
from s2tc_b200 import Settings, synth
enc = s2tc_b200.Encoder(0)
bad = 0
imgs = [synth.synth_rgba(64, 36, seed=3), synth.synth_noise(37, 21, seed=5), synth.synth_noise(16, 16, seed=6, comps=3)]
for img in imgs:
    for dxt in (s2tc_b200.DXT1, s2tc_b200.DXT3, s2tc_b200.DXT5):
        for cd in (s2tc_b200.WAVG, s2tc_b200.SRGB_MIXED, s2tc_b200.SRGB, s2tc_b200.NORMALMAP):
            for nr in (-1, 0, 3, 64):
                for dither in (s2tc_b200.DITHER_SIMPLE, s2tc_b200.DITHER_NONE):
                    s = Settings(dxt, cd, nr, s2tc_b200.REFINE_LOOP, dither)
                    got = enc.compress(img, s)
                    want = O.orc_compress(img, s.dxt, s.cd, s.nrandom, s.refine, s.dither)
                    if not np.array_equal(got, want):
                        bad += 1
                        print("MISMATCH", img.shape, dxt, cd, nr, dither)
big = synth.synth_rgba(1024, 512, seed=9)   # several dither tiles
for dxt in (s2tc_b200.DXT1, s2tc_b200.DXT3, s2tc_b200.DXT5):
    s = Settings(dxt, s2tc_b200.WAVG, 0, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_SIMPLE)
    if not np.array_equal(enc.compress(big, s), O.orc_compress(big, s.dxt, s.cd, s.nrandom, s.refine, s.dither)):
        bad += 1; print("MISMATCH big", dxt)
# round 2: Floyd-Steinberg (whole image and as a chain of row shards), striped host shards, staged pageable memory
import torch
from s2tc_b200.sharding import shard_block_rows
fs_img = synth.synth_noise(70, 77, seed=11)
for dxt, abits in ((s2tc_b200.DXT1, 1), (s2tc_b200.DXT3, 4), (s2tc_b200.DXT5, 8)):
    s = Settings(dxt, s2tc_b200.WAVG, -1, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_FLOYDSTEINBERG)
    want = O.orc_compress(fs_img, s.dxt, s.cd, s.nrandom, s.refine, s.dither)
    if not np.array_equal(enc.compress(fs_img, s), want):
        bad += 1; print("MISMATCH floyd", dxt)
    h, w = fs_img.shape[:2]
    bh, bw = (h + 3) // 4, (w + 3) // 4
    d_img = torch.from_numpy(fs_img).cuda()
    ranges = [shard_block_rows(bh, 3, r) for r in range(3)]
    reds, err = [], None
    for a, b in ranges:
        rows = d_img[4 * a:min(4 * b, h)].contiguous()
        red = torch.zeros(rows.shape[0] * w, dtype=torch.int32, device="cuda")
        eo = torch.zeros(3 * w, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        enc.floyd_rows_device(rows, w, h, 4, abits, a, b, 0, err, eo, red)
        enc.sync()
        err = eo
        reds.append((rows, red))
    if abits != 8:
        err = err[:w].clone()
        for (a, b), (rows, red) in zip(ranges, reds):
            eo = torch.zeros(w, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            enc.floyd_rows_device(rows, w, h, 4, abits, a, b, 1, err, eo, red)
            enc.sync()
            err = eo
    outs = []
    for (a, b), (rows, red) in zip(ranges, reds):
        d_out = torch.zeros((b - a) * bw * O.block_bytes(dxt), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        enc.encode_reduced_rows_device(red, w, h, a, b, d_out, s)
        enc.sync()
        outs.append(d_out.cpu().numpy())
    if not np.array_equal(np.concatenate(outs), want):
        bad += 1; print("MISMATCH floyd shards", dxt)
st_img = synth.synth_noise(200, 150, seed=12)
for nr in (-1, 9):
    s = Settings(s2tc_b200.DXT1, s2tc_b200.WAVG, nr, s2tc_b200.REFINE_LOOP, s2tc_b200.DITHER_SIMPLE)
    h, w = st_img.shape[:2]
    bw, bh = (w + 3) // 4, (h + 3) // 4
    out = np.zeros(bw * bh * 8, np.uint8)
    rows = [s2tc_b200.Encoder.stripe_rows(h, 1, 3, k, 0, [1, 2, 1]) for k in range(3)]
    enc.compress_striped([np.ascontiguousarray(st_img[4 * a:min(4 * b, h)]) for a, b in rows], w, h,
                         [out[a * bw * 8:b * bw * 8] for a, b in rows], s, 0, 1, 3, weights=[1, 2, 1])
    if not np.array_equal(out, O.orc_compress(st_img, s.dxt, s.cd, s.nrandom, s.refine, s.dither)):
        bad += 1; print("MISMATCH striped", nr)
pg = synth.synth_rgba(2048, 1200, seed=13)    # 9.4 MiB pageable: staged in two chunks
s = Settings(s2tc_b200.DXT1, s2tc_b200.WAVG, -1, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_SIMPLE)
if not np.array_equal(enc.compress(pg, s)[:2 * 512 * 8], O.orc_rows(pg, s.dxt, s.cd, s.nrandom, s.refine, s.dither, (0, 2))):
    bad += 1; print("MISMATCH staged")
print("sanitizer workload done, mismatches:", bad)
