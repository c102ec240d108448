#!/bin/bash
set -u
OUT=gpurun_out/job8; mkdir -p $OUT
ncu --clock-control none --set full --import-source on -k regex:pair_search -s 3 -c 1 -f -o /tmp/ps python bench.py --steps 1 --kernel-only --no-check --workload config3 --size 4096 > /dev/null 2>&1
ncu -i /tmp/ps.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/ps_sass.csv.gz
python profiles/ncu_summary.py /tmp/ps.ncu-rep > $OUT/ps.ncu.txt 2>&1
ncu -i /tmp/ps.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for i,k in enumerate(h):
    if 'pipe' in k and 'inst_executed' in k: print(k, r[i])
" > $OUT/ps_pipes.txt
