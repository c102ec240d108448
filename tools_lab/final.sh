#!/bin/bash
# tools_lab/final.sh TAG -- the end-of-step routine on the GPU box: GPU tests, the four bench workloads (full legs),
# the reference arm, smoke(), then profiles/capture.sh TAG.  Everything lands under gpurun_out/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out/final_$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
python bench.py > $OUT/bench_config2.json 2> $OUT/bench_config2.err
python bench.py --workload config3 --steps 3 > $OUT/bench_config3.json 2> $OUT/bench_config3.err
python bench.py --workload config5 --steps 5 > $OUT/bench_config5.json 2> $OUT/bench_config5.err
python bench.py --workload defaults --steps 5 > $OUT/bench_defaults.json 2> $OUT/bench_defaults.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python - <<PY
import json
for s in ["config2","config3","config5","defaults","reference"]:
    try:
        d=json.loads(open("$OUT/bench_%s.json" % s).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(s, round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "e2e", d["e2e"].get("ms_per_step") and round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "int", r.get("int32") and round(r["int32"]["frac"],3), "hbm", r.get("frac") and round(r["frac"],4), d.get("gpu_launches"))
    except Exception as e: print(s, "ERR", e, open("$OUT/bench_%s.err" % s).read()[-600:])
PY
bash profiles/capture.sh $TAG > $OUT/capture.log 2>&1; tail -3 $OUT/capture.log
