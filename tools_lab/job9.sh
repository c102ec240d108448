#!/bin/bash
set -u
OUT=gpurun_out/job9; mkdir -p $OUT
python bench.py --steps 3 --kernel-only --workload config3 > $OUT/c3.json 2> $OUT/c3.err
python bench.py --steps 10 --kernel-only > $OUT/c2.json 2> $OUT/c2.err
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
tail -5 $OUT/pytest.log
python - <<'PY'
import json
for s in ["c2","c3"]:
    try:
        d=json.loads(open(f"gpurun_out/job9/{s}.json").read().strip().splitlines()[-1])
        print(s, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d.get("checked_blocks_vs_oracle"))
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job9/{s}.err").read()[-800:])
PY
