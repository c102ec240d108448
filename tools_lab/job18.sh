#!/bin/bash
set -u
OUT=gpurun_out/job18; mkdir -p $OUT
for mb in 4 16 64; do
  S2TC_B200_SLAB_MB=$mb python bench.py --workload config3 --steps 3 --no-check --cpu-rows 4 > $OUT/s$mb.json 2> $OUT/s$mb.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/job18/s$mb.json").read().strip().splitlines()[-1])
print("slab_mb(x4)=$mb", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],1))
PY
done
