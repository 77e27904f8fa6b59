#!/bin/bash
TAG=${1:-r2l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for carve in -1 50 100; do
echo "== spmm_tail_carveout=$carve"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm 12000 --i8 1 --head-max 8192 --opt spmm_tail_carveout=$carve > $OUT/spmm_check_$carve.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check_$carve.log | grep "per product" | tail -2
done
