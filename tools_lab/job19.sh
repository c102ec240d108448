#!/bin/bash
set -u
OUT=gpurun_out/job20; mkdir -p $OUT
S2TC_B200_TRACE=1 python bench.py --workload config3 --steps 1 --no-check --cpu-rows 4 > $OUT/t3.json 2> $OUT/t3.err
grep "trace" $OUT/t3.err | tail -18
S2TC_B200_TRACE=1 python bench.py --steps 1 --no-check --cpu-rows 4 > $OUT/t2.json 2> $OUT/t2.err
grep "trace" $OUT/t2.err | tail -17
