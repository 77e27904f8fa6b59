#!/bin/bash
TAG=${1:-r2hint}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for opts in "spmm_head8_stages=4" "spmm_head8_stages=2"; do
for ppm in 12000 7000; do
echo "== ppm=$ppm $opts"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm $ppm --i8 1 --head-max 8192 --opt $opts > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product" | tail -2 | cut -c1-250
done; done
