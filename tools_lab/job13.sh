#!/bin/bash
set -u
OUT=gpurun_out/job13; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -3
python bench.py --workload config3 --steps 3 --kernel-only > $OUT/c3.json 2> $OUT/c3.err
python bench.py --steps 5 --kernel-only > $OUT/c2.json 2> $OUT/c2.err
python - <<'PY'
import json
for s in ["c2","c3"]:
    try:
        d=json.loads(open(f"gpurun_out/job13/{s}.json").read().strip().splitlines()[-1])
        print(s, round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["roofline"]["int32"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job13/{s}.err").read()[-800:])
PY
