#!/bin/bash
TAG=${1:-r2o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for ppm in 9000 7000 6000 5000 4000; do
for opts in "spmm_head8_stages=3" ; do
echo "== ppm=$ppm $opts"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm $ppm --i8 1 --head-max 8192 --opt $opts > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product" | tail -2
done; done
