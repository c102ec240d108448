#!/bin/bash
# tools_lab/ab.sh NAME "EXTRA nvcc flags" -- A/B build of the library under another name:
#   s2tc_b200/csrc/build_NAME/ (objects, gpurun-ignored) and s2tc_b200/lib_NAME/libs2tc_b200.so (travels with gpurun).
# Run a workload against it with S2TC_B200_LIBDIR=$PWD/s2tc_b200/lib_NAME python bench.py --kernel-only ...
# Useful flags: -DS2TC_SEARCH16_ONLY_CD=3 (compile search16 for one metric only: 30 s instead of 2 min),
# -DS2TC_SEARCH16_CHAINS=n, -DS2TC_SEARCH16_MINBLOCKS=n, -DS2TC_FINISH_MINBLOCKS=n, -DS2TC_APPLY_MINBLOCKS=n, -DS2TC_MAPS_MINBLOCKS=n, -DS2TC_SEARCH16_ONE_LAUNCH.
set -eu
NAME=$1; EXTRA=${2:-}
cd "$(dirname "$0")/../s2tc_b200/csrc"
make -j8 BUILD=build_$NAME LIBDIR=../lib_$NAME BINDIR=../bin_$NAME EXTRA="$EXTRA"
rm -f ../lib_$NAME/libtxc_dxtn.so   # same file as libs2tc_b200.so; halves the snapshot gpurun pushes
grep -h -A2 "Compiling entry function" build_$NAME/*.ptxas.log | grep -B2 "spill stores" | grep -v "^--" | grep -v " 0 bytes spill stores" | head -20 || true
