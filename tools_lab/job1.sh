#!/bin/bash
# A/B: fused encode16 vs split (search kernel + finish kernel); source-level ncu capture of the fused kernel
set -u
OUT=gpurun_out/job1; mkdir -p $OUT
for split in 0 1; do
  S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_x S2TC_B200_ENCODE16_SPLIT=$split python bench.py --steps 10 --kernel-only > $OUT/split$split.json 2> $OUT/split$split.err
done
ncu --clock-control none --set full --import-source on -k regex:encode16 -s 3 -c 1 -f -o /tmp/enc16 python bench.py --steps 1 --kernel-only --no-check > /dev/null 2>&1
ncu -i /tmp/enc16.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/enc16_sass.csv.gz
ncu -i /tmp/enc16.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $OUT/enc16_cuda_sass.csv.gz
python profiles/ncu_summary.py /tmp/enc16.ncu-rep > $OUT/enc16.ncu.txt 2>&1
ls -la $OUT
python - <<'PY'
import json
for s in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/job1/split{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job1/split{s}.err").read()[-500:])
PY
