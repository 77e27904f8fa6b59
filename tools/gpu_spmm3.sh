#!/bin/bash
TAG=${1:-r2i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python tools/spmm_check.py --config c2 --no-ref ${SPMM_ARGS:---density-ppm 12000 8000 6000 --i8 1 --head-max 8192} > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product" | tail -40
