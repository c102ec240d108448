#!/bin/bash
# ncu capture of the config-3 kernels on a 4096x4096 slab: tools_lab/cap_c3.sh TAG [kernel regex]
TAG=${1:-x}
RE=${2:-pair_search}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
ncu --clock-control none --set full --import-source on -k regex:"$RE" -s 1 -c 1 -f -o /tmp/prof_c3 python bench.py --workload config3 --size 4096 --steps 1 --kernel-only --no-check > $OUT/bench.log 2>&1
python profiles/ncu_summary.py /tmp/prof_c3.ncu-rep > $OUT/c3.ncu.txt 2>&1
python profiles/ncu_source.py /tmp/prof_c3.ncu-rep 48 > $OUT/c3.source.txt 2>&1
ncu -i /tmp/prof_c3.ncu-rep --page source --csv --print-source sass > $OUT/c3.sass.csv 2>/dev/null
ls -la $OUT
