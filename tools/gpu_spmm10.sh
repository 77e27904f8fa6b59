#!/bin/bash
TAG=${1:-r2t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python tools/spmm_check.py --config c2 --density-ppm 12000 --i8 1 --only-head-serial --reps 1 ${SPMM_ARGS} > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | tail -12
