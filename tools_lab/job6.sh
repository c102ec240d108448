#!/bin/bash
set -u
OUT=gpurun_out/job6; mkdir -p $OUT
for v in a b c d e; do
  S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_$v python bench.py --steps 10 --kernel-only > $OUT/$v.json 2> $OUT/$v.err
done
python - <<'PY'
import json
for s in "abcde":
    try:
        d=json.loads(open(f"gpurun_out/job6/{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job6/{s}.err").read()[-800:])
PY
