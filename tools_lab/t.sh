#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2
python bench.py --steps 10 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config2', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['checked_blocks_vs_oracle'])"
ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file /tmp/l.csv python bench.py --steps 2 --kernel-only --no-check > /dev/null 2>&1
grep -E "search16|finish" /tmp/l.csv | tail -3 | awk -F'","' '{print substr($5,1,50), $NF}'
