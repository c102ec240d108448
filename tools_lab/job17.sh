#!/bin/bash
set -u
OUT=gpurun_out/job17; mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools_lab/san.py > $OUT/$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|mismatches|Invalid|Race|hazard" $OUT/$tool.log | head -12
done
