#!/bin/bash
# operator engines: parity tests of the operator, then per-kernel timings on c2
TAG=${1:-r2d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spsptr or engines or block_ks_tiny or block_ks_c1" > $OUT/pytest_op.log 2>&1; echo "op tests rc=$?" | tee -a $OUT/pytest_op.log
grep -v "^OpenBLAS" $OUT/pytest_op.log | tail -25
timeout 600 python tools/spmm_check.py --config c2 ${SPMM_ARGS:---density-ppm 12000 6000 4000 --i8 1 0 --head-max 8192} > $OUT/spmm_check.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check.log | tail -40
