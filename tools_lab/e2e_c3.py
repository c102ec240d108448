"""tools_lab/e2e_c3.py W H [reps] -- the config-3 host-to-host call (s2tc_b200_compress_host, pinned buffers) against the
device-resident encode of the same texture: how much of the end-to-end step is head and tail of the slab pipeline.
Run with S2TC_B200_SLAB_MB=n and S2TC_B200_TRACE=1 to see the per-slab timeline."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import s2tc_b200
from s2tc_b200 import Settings, synth

w, h = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
st = Settings(s2tc_b200.DXT1, s2tc_b200.WAVG, 64, s2tc_b200.REFINE_LOOP, s2tc_b200.DITHER_SIMPLE)
enc = s2tc_b200.Encoder(0)
img = synth.synth_rgba(w, h, 1234)
h_src = torch.from_numpy(img).pin_memory()
nb = ((w + 3) // 4) * ((h + 3) // 4)
h_dst = torch.empty(nb * 8, dtype=torch.uint8).pin_memory()
d_src = h_src.cuda()
d_dst = torch.empty(nb * 8, dtype=torch.uint8, device="cuda")
for _ in range(2):
    enc.compress(h_src, st, cursor=0, out=h_dst)
t0 = time.perf_counter()
for _ in range(reps):
    enc.compress(h_src, st, cursor=0, out=h_dst)
ms_e2e = (time.perf_counter() - t0) * 1e3 / reps
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for _ in range(2):
    enc.encode_rows_device(d_src, w, h, 4, 0, (h + 3) // 4, d_dst, st, cursor0=0, carry=None, stream=stream.cuda_stream)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(reps):
    enc.encode_rows_device(d_src, w, h, 4, 0, (h + 3) // 4, d_dst, st, cursor0=0, carry=None, stream=stream.cuda_stream)
b.record(stream)
torch.cuda.synchronize()
ms_dev = a.elapsed_time(b) / reps
same = bool(np.array_equal(h_dst.numpy(), d_dst.cpu().numpy()))
print(f"e2e {ms_e2e:.2f} ms  device {ms_dev:.2f} ms  tail {ms_e2e - ms_dev:.2f} ms  same_bytes {same}")
