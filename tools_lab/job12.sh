#!/bin/bash
set -u
OUT=gpurun_out/job12; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
tail -3 $OUT/pytest.log
python bench.py > $OUT/bench_config2.json 2> $OUT/bench_config2.err
python bench.py --workload config3 --steps 3 > $OUT/bench_config3.json 2> $OUT/bench_config3.err
python bench.py --workload config5 --steps 5 > $OUT/bench_config5.json 2> $OUT/bench_config5.err
python bench.py --workload defaults --steps 5 > $OUT/bench_defaults.json 2> $OUT/bench_defaults.err
python - <<'PY'
import json
for s in ["config2","config3","config5","defaults"]:
    try:
        d=json.loads(open(f"gpurun_out/job12/bench_{s}.json").read().strip().splitlines()[-1])
        print(s, round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "int32", d["roofline"]["int32"] and round(d["roofline"]["int32"]["frac"],3), d["gpu_launches"], d["clocks"])
    except Exception as e: print(s, "ERR", e, open(f"gpurun_out/job12/bench_{s}.err").read()[-800:])
PY
