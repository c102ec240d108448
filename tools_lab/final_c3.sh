set -u
OUT=gpurun_out/final_r02c; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
timeout 300 python bench.py > $OUT/bench_config3.json 2> $OUT/bench_config3.err
timeout 200 python bench.py --workload config5 --steps 5 > $OUT/bench_config5.json 2> $OUT/bench_config5.err
NCU="ncu --clock-control none"
timeout 240 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_config3.csv python bench.py --steps 2 --kernel-only --no-check > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"pair_search|rand_windows|finish_kernel" -s 3 -c 3 -f -o /tmp/prof_c3 python bench.py --size 4096 --steps 1 --kernel-only --no-check > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/prof_c3.ncu-rep > $OUT/config3_search_windows_finish.ncu.txt 2>&1
python profiles/ncu_source.py /tmp/prof_c3.ncu-rep 32 > $OUT/config3_search_windows_finish.source.txt 2>&1
python - <<PY
import json
for s in ["config3","config5"]:
    d=json.loads(open("$OUT/bench_%s.json" % s).read().strip().splitlines()[-1])
    r=d["roofline"]; e=d["e2e"]
    print(s, round(d["ms_per_step"],3), round(d["value"],1), r["kernel_ms_per_step"], "e2e", round(e["ms_per_step"],2), round(e["value"],1), "pageable", (e.get("pageable") or {}).get("ms_per_step"), "frac", round(r["frac"],4))
PY
grep -E "smsp__inst_executed.sum |issue_active.avg.pct_of_peak_sustained_active|pipe_alu.sum.pct|pipe_fma.sum.pct|gpu__time_duration" $OUT/config3_search_windows_finish.ncu.txt | sed -n 6,10p
