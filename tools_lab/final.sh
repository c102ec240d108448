#!/bin/bash
# tools_lab/final.sh TAG -- the end-of-step routine on the GPU box: GPU tests, smoke(), the bench workloads (full legs),
# the reference arm, then profiles/capture.sh TAG.  Everything lands under gpurun_out/.  Every step has its own timeout.
set -u
TAG=${1:-rXX}
OUT=gpurun_out/final_$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 300 python bench.py > $OUT/bench_config3.json 2> $OUT/bench_config3.err
timeout 200 python bench.py --workload config2 --steps 5 > $OUT/bench_config2.json 2> $OUT/bench_config2.err
timeout 200 python bench.py --workload config5 --steps 5 > $OUT/bench_config5.json 2> $OUT/bench_config5.err
timeout 200 python bench.py --workload defaults --steps 5 > $OUT/bench_defaults.json 2> $OUT/bench_defaults.err
timeout 300 python bench.py --workload config4 --steps 5 > $OUT/bench_config4.json 2> $OUT/bench_config4.err
timeout 200 python bench.py --workload defaults --dither FLOYDSTEINBERG --steps 3 > $OUT/bench_defaults_fs.json 2> $OUT/bench_defaults_fs.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python - <<PY
import json
for s in ["config3","config2","config5","defaults","config4","defaults_fs","reference"]:
    try:
        d=json.loads(open("$OUT/bench_%s.json" % s).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        e=d.get("e2e") or {}
        print(s, round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "e2e", e.get("ms_per_step") and round(e["ms_per_step"],2), e.get("value") and round(e["value"],1), "pageable", (e.get("pageable") or {}).get("ms_per_step"), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],2), "bound", r.get("bound"), "frac", r.get("frac") and round(r["frac"],4), d.get("gpu_launches"))
    except Exception as ex: print(s, "ERR", ex, open("$OUT/bench_%s.err" % s).read()[-600:])
PY
timeout 900 bash profiles/capture.sh $TAG > $OUT/capture.log 2>&1; tail -3 $OUT/capture.log
