#!/bin/bash
# tools_lab/sanitize.sh -- compute-sanitizer memcheck + racecheck over tools_lab/san.py on the GPU box
OUT=gpurun_out/sanitize; mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools_lab/san.py > $OUT/$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|mismatches|Invalid|hazard" $OUT/$tool.log | head -8
done
