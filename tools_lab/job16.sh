#!/bin/bash
set -u
OUT=gpurun_out/job16; mkdir -p $OUT
for v in a b; do
  L=$PWD/s2tc_b200/lib_$v; [ $v = a ] && L=$PWD/s2tc_b200/lib
  S2TC_B200_LIBDIR=$L python bench.py --steps 10 --kernel-only > $OUT/$v.json 2> $OUT/$v.err
done
S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_b ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --kernel-only --no-check > /dev/null 2>&1
grep -E "search16|finish" $OUT/launches.csv | tail -3 | awk -F'","' '{print substr($5,1,50), $NF}'
python - <<'PY'
import json
for s in ["a","b"]:
    try:
        d=json.loads(open(f"gpurun_out/job16/{s}.json").read().strip().splitlines()[-1])
        print(s, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job16/{s}.err").read()[-800:])
PY
