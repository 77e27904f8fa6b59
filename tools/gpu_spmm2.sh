#!/bin/bash
TAG=${1:-r2h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spsptr or engines" > $OUT/pytest_op.log 2>&1; echo "op tests rc=$?" | tee -a $OUT/pytest_op.log
grep -v "^OpenBLAS" $OUT/pytest_op.log | tail -8
timeout 600 python tools/spmm_check.py --config c2 --no-ref ${SPMM_ARGS:---density-ppm 12000 6000 --i8 1 --head-max 8192} > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | grep "per product\|H=" | tail -40
ISLE_HEAD8_TRACE=$OUT/trace.txt timeout 600 python tools/spmm_check.py --config c2 --only-head-serial --no-ref --reps 1 --density-ppm 6000 --head-max 8192 > $OUT/run.log 2>&1
python tools/head8_trace.py $OUT/trace.txt > $OUT/trace_summary.txt; tail -24 $OUT/trace_summary.txt
