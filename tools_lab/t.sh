#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_sharding.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
for wl in defaults config2 config3; do
python bench.py --steps 5 --kernel-only --workload $wl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['checked_blocks_vs_oracle'], d['gpu_launches'])"
done
ncu --clock-control none --metrics gpu__time_duration.sum -c 40 --csv --log-file /tmp/l.csv python bench.py --steps 2 --kernel-only --no-check --workload defaults > /dev/null 2>&1
grep -E "dither|fast|scan_" /tmp/l.csv | tail -6 | awk -F'","' '{print substr($5,1,40), $NF}'
