#!/bin/bash
# profiles/capture.sh TAG -- run on the GPU box (gpurun -- 'bash profiles/capture.sh r01f'):
# launch lists (gpu__time_duration per launch) and `ncu --set full` captures of the main kernels for the three bench
# workloads; the .ncu-rep files are summarised on the box (profiles/ncu_summary.py) because only 64 MiB come back.
set -u
TAG=${1:-rXX}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
NCU="ncu --clock-control none"
launches() { # name, bench args
	$NCU --metrics gpu__time_duration.sum -c 150 --csv --log-file $OUT/launches_$1.csv python bench.py ${@:2} --steps 2 --kernel-only --no-check > /dev/null 2>&1
}
full() { # name, kernel regex, skip, count, bench args
	$NCU --set full --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/prof_$1 python bench.py ${@:5} --steps 1 --kernel-only --no-check > /dev/null 2>&1
	python profiles/ncu_summary.py /tmp/prof_$1.ncu-rep > $OUT/$1.ncu.txt 2>&1
	python profiles/ncu_source.py /tmp/prof_$1.ncu-rep 32 > $OUT/$1.source.txt 2>&1
	rm -f /tmp/prof_$1.ncu-rep
}
launches config2
full config2_search16_finish "search16|finish_kernel" 9 3
launches config3_4096 --workload config3 --size 4096
full config3_search_cand_finish "pair_search|random_cand|finish_kernel" 3 3 --workload config3 --size 4096
launches defaults --workload defaults
full defaults_fast_dither "fast_encode|dither_|scan_" 6 6 --workload defaults
launches config5 --workload config5
full config5_search16_finish "search16|finish_kernel" 9 3 --workload config5
ls -la $OUT
