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
            for nr in (-1, 0, 3, 40, 64):
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
print("sanitizer workload done, mismatches:", bad)
