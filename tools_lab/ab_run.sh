#!/bin/bash
# tools_lab/ab_run.sh OUTTAG VARIANT...  -- on the GPU box: for every A/B build (tools_lab/ab.sh) run the nrandom > 0
# parity tests and a kernel-only config-3 bench line (8192^2 slab unless SIZE is set); results under gpurun_out/ab_OUTTAG/.
TAG=$1; shift
OUT=gpurun_out/ab_$TAG
mkdir -p $OUT
SIZE=${SIZE:-8192}
WORKLOAD=${WORKLOAD:-config3}
TESTS=${TESTS:-"tests/test_gpu_parity.py tests/test_gpu_golden.py"}
for v in "$@"; do
	export S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_$v
	[ "$v" = main ] && export S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib
	if [ -n "$TESTS" ]; then
		timeout 900 python -m pytest $TESTS -m gpu -x -q > $OUT/test_$v.log 2>&1
		echo "$v tests rc=$? $(tail -1 $OUT/test_$v.log)"
	fi
	for w in $WORKLOAD; do
		timeout 600 python bench.py --workload $w --size $SIZE --steps ${STEPS:-5} --warmup 3 --kernel-only > $OUT/bench_${w}_$v.json 2> $OUT/bench_${w}_$v.err
		echo "$v $w rc=$? $(python - $OUT/bench_${w}_$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms_per_step", d.get("ms_per_step"), "value", d.get("value"), "fam", (d.get("roofline") or {}).get("kernel_ms_per_step"), "checked", d.get("checked_blocks_vs_oracle"))
except Exception as e:
    print("no json", e)
PY
)"
	done
done
unset S2TC_B200_LIBDIR
