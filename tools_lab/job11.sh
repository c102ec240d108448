#!/bin/bash
set -u
OUT=gpurun_out/job11; mkdir -p $OUT
for v in a b; do
  S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_$v python bench.py --steps 10 --kernel-only > $OUT/$v.json 2> $OUT/$v.err
done
S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_a ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --kernel-only --no-check > /dev/null 2>&1
grep -E "search16|finish" $OUT/launches.csv | tail -6 | awk -F'","' '{print $5, $NF}' | cut -c1-160
python - <<'PY'
import json
for s in ["a","b"]:
    try:
        d=json.loads(open(f"gpurun_out/job11/{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job11/{s}.err").read()[-800:])
PY
