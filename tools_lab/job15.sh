#!/bin/bash
set -u
OUT=gpurun_out/job15; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_sharding.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
for v in a b c; do
  L=$PWD/s2tc_b200/lib_$v; [ $v = a ] && L=$PWD/s2tc_b200/lib
  S2TC_B200_LIBDIR=$L python bench.py --steps 10 --kernel-only --workload defaults > $OUT/$v.json 2> $OUT/$v.err
done
python bench.py --steps 3 --kernel-only --workload config3 > $OUT/c3.json 2> $OUT/c3.err
ncu --clock-control none --metrics gpu__time_duration.sum -c 40 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --kernel-only --no-check --workload defaults > /dev/null 2>&1
grep -E "dither|fast" $OUT/launches.csv | tail -4 | awk -F'","' '{print $5, $NF}' | cut -c1-120
python - <<'PY'
import json
for s in ["a","b","c","c3"]:
    try:
        d=json.loads(open(f"gpurun_out/job15/{s}.json").read().strip().splitlines()[-1])
        print(s, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job15/{s}.err").read()[-800:])
PY
