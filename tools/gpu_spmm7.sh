#!/bin/bash
TAG=${1:-r2m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for opts in "spmm_tail_pipe=0" "spmm_tail_pipe=2" "spmm_tail_pipe=3" "spmm_tail_pipe=2 spmm_tail_carveout=100" "spmm_tail_pipe=2 spmm_head8_slab=2 spmm_head8_stages=4"; do
echo "== $opts"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm 12000 --i8 1 --head-max 8192 --opt $opts > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product" | tail -2
done
