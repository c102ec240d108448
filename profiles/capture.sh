#!/bin/bash
# profiles/capture.sh TAG -- run on the GPU box (gpurun -- 'bash profiles/capture.sh r02a'):
# launch lists (gpu__time_duration per launch) and `ncu --set full` captures of the main kernels for the bench
# workloads; the .ncu-rep files are summarised on the box (profiles/ncu_summary.py, ncu_source.py) because only 64 MiB
# come back.  bench.py's default workload is config 3 (the north-star configuration).
set -u
TAG=${1:-rXX}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
NCU="ncu --clock-control none"
launches() { # name, bench args
	timeout 240 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_$1.csv python bench.py ${@:2} --steps 2 --kernel-only --no-check > /dev/null 2>&1
}
full() { # name, kernel regex, skip, count, bench args
	timeout 300 $NCU --set full --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/prof_$1 python bench.py ${@:5} --steps 1 --kernel-only --no-check > /dev/null 2>&1
	python profiles/ncu_summary.py /tmp/prof_$1.ncu-rep > $OUT/$1.ncu.txt 2>&1
	python profiles/ncu_source.py /tmp/prof_$1.ncu-rep 32 > $OUT/$1.source.txt 2>&1
	rm -f /tmp/prof_$1.ncu-rep
}
# config 3 (default): the same command as the bench line for the launch list; a 4096x4096 slab for the full captures
launches config3
full config3_search_windows_finish "pair_search|rand_windows|finish_kernel" 3 3 --size 4096
launches config2 --workload config2
full config2_search16_finish "search16|finish_kernel" 9 3 --workload config2
launches defaults --workload defaults
full defaults_fast_dither "fast_encode|dither_|scan_" 6 6 --workload defaults
launches config4 --workload config4 --textures 8
full config4_fast "fast_encode" 40 2 --workload config4 --textures 8
launches config5 --workload config5
ls -la $OUT
