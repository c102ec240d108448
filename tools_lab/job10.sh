#!/bin/bash
set -u
OUT=gpurun_out/job10; mkdir -p $OUT
for v in a b c; do
  L=$PWD/s2tc_b200/lib_$v; [ $v = a ] && L=$PWD/s2tc_b200/lib
  S2TC_B200_LIBDIR=$L python bench.py --steps 10 --kernel-only > $OUT/$v.json 2> $OUT/$v.err
  S2TC_B200_LIBDIR=$L python bench.py --steps 3 --kernel-only --workload config3 --size 8192 > $OUT/${v}3.json 2> $OUT/${v}3.err
done
python - <<'PY'
import json
for s in ["a","b","c","a3","b3","c3"]:
    try:
        d=json.loads(open(f"gpurun_out/job10/{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job10/{s}.err").read()[-800:])
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
