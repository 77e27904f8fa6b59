#!/bin/bash
TAG=${1:-r2g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
ISLE_HEAD8_TRACE=$OUT/trace.txt timeout 600 python tools/spmm_check.py --config c2 --only-head-serial --no-ref --reps 1 ${SPMM_ARGS:---density-ppm 6000 --head-max 8192} > $OUT/run.log 2>&1
echo "rc=$?"; tail -3 $OUT/run.log; python tools/head8_trace.py $OUT/trace.txt > $OUT/trace_summary.txt; tail -50 $OUT/trace_summary.txt
