"""Floyd-Steinberg pre-pass timing by image height (bands): python tools_lab/fs_timing.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import s2tc_b200
from s2tc_b200 import Settings, synth
enc = s2tc_b200.Encoder(0)
st = Settings(s2tc_b200.DXT5, s2tc_b200.WAVG, -1, s2tc_b200.REFINE_ALWAYS, s2tc_b200.DITHER_FLOYDSTEINBERG)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for w in (256, 8192,):
    for h in (32, 64, 96, 128, 2048, 8192):
        img = torch.from_numpy(synth.synth_noise(w, h, seed=1)).cuda()
        out = torch.empty((w // 4) * (h // 4) * 16, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            enc.encode_rows_device(img, w, h, 4, 0, h // 4, out, st, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        enc.profile(True); enc.profile_read(reset=True)
        for _ in range(3):
            enc.encode_rows_device(img, w, h, 4, 0, h // 4, out, st, stream=stream.cuda_stream)
        fam = enc.profile_read(reset=True); enc.profile(False)
        ms = fam["prepass"][0] / 3
        steps = w + 63 + (h // 32 - 1) * 71
        print(f"w={w} h={h} bands={h//32}: prepass {ms:.3f} ms; model steps {steps}: {ms*1e6/steps:.0f} ns/step", flush=True)
