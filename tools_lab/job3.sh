#!/bin/bash
set -u
OUT=gpurun_out/job4; mkdir -p $OUT
S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_x S2TC_B200_ENCODE16_SPLIT=0 python bench.py --steps 10 --kernel-only > $OUT/x0.json 2> $OUT/x0.err
for v in x y z w; do
  S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_$v S2TC_B200_ENCODE16_SPLIT=1 python bench.py --steps 10 --kernel-only > $OUT/$v.json 2> $OUT/$v.err
done
export S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_x S2TC_B200_ENCODE16_SPLIT=1
ncu --clock-control none --set full --import-source on -k regex:encode16 -s 3 -c 1 -f -o /tmp/enc16 python bench.py --steps 1 --kernel-only --no-check > /dev/null 2>&1
ncu -i /tmp/enc16.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/enc16_sass.csv.gz
python profiles/ncu_summary.py /tmp/enc16.ncu-rep > $OUT/enc16.ncu.txt 2>&1
ncu --clock-control none --set full --import-source on -k regex:finish_kernel -s 3 -c 1 -f -o /tmp/fin python bench.py --steps 1 --kernel-only --no-check > /dev/null 2>&1
ncu -i /tmp/fin.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/fin_sass.csv.gz
python profiles/ncu_summary.py /tmp/fin.ncu-rep > $OUT/fin.ncu.txt 2>&1
python - <<'PY'
import json
for s in ["x0","x","y","z","w"]:
    try:
        d=json.loads(open(f"gpurun_out/job4/{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job4/{s}.err").read()[-800:])
PY
