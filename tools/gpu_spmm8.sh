#!/bin/bash
TAG=${1:-r2n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spsptr or engines" > $OUT/pytest_op.log 2>&1; echo "op tests rc=$?"; tail -3 $OUT/pytest_op.log
for opts in "spmm_head8_stages=4" "spmm_head8_stages=3" "spmm_head8_stages=2" "spmm_head8_slab=4 spmm_head8_stages=6" "spmm_head8_stages=4 spmm_head2_waves=2"; do
echo "== $opts"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm ${PPM:-12000} --i8 1 --head-max 8192 --opt $opts > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product" | tail -2
done
