#!/bin/bash
set -u
OUT=gpurun_out/job20; mkdir -p $OUT
S2TC_B200_TRACE=1 python bench.py --workload config3 --steps 1 --no-check --cpu-rows 4 > $OUT/t3.json 2> $OUT/t3.err
grep "trace" $OUT/t3.err | tail -18
S2TC_B200_TRACE=1 python bench.py --steps 1 --no-check --cpu-rows 4 > $OUT/t2.json 2> $OUT/t2.err
grep "trace" $OUT/t2.err | tail -17
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q 2>&1 | tail -2
python bench.py --workload config3 --steps 3 --cpu-rows 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1), d['gpu_launches'], d['checked_blocks_vs_oracle'])"
