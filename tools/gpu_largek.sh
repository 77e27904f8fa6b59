#!/bin/bash
TAG=${1:-r2b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_largek.py -m gpu -q -rP --durations=5 > $OUT/pytest_largek.log 2>&1; echo "largek rc=$?" | tee -a $OUT/pytest_largek.log
grep -v "^E   \|^$" $OUT/pytest_largek.log | tail -60
