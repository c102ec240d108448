#!/bin/bash
set -u
OUT=gpurun_out/job14; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for mb in 8 12 16 24 32 64; do
  S2TC_B200_SLAB_MB=$mb python bench.py --steps 10 --no-check --cpu-rows 4 > $OUT/s$mb.json 2> $OUT/s$mb.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/job14/s$mb.json").read().strip().splitlines()[-1])
print("slab_mb=$mb", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],1))
PY
done
S2TC_B200_SLAB_MB=16 python bench.py --steps 10 --no-check --cpu-rows 4 --workload defaults 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('defaults e2e', round(d['e2e']['ms_per_step'],3))"
