#!/bin/bash
set -u
OUT=gpurun_out/job2; mkdir -p $OUT
for split in 0 1; do
  S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_x S2TC_B200_ENCODE16_SPLIT=$split python bench.py --steps 10 --kernel-only > $OUT/split$split.json 2> $OUT/split$split.err
done
python - <<'PY'
import json
for s in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/job2/split{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job2/split{s}.err").read()[-800:])
PY
