#!/bin/bash
TAG=${1:-r2dummy}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for opts in "spmm_dummy_head=60 spmm_dummy_smem_kb=57" "spmm_dummy_head=100 spmm_dummy_smem_kb=57" "spmm_dummy_head=100 spmm_dummy_smem_kb=24" "spmm_dummy_head=100 spmm_dummy_smem_kb=1"; do
echo "== $opts"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm 12000 --i8 1 --opt $opts > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product" | grep fork | cut -c1-120
done
