#!/bin/bash
TAG=${1:-r2m3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_integration_gpu.py -m gpu -q > $OUT/pytest_integration.log 2>&1; echo "pytest rc=$?"
grep -v "^OpenBLAS" $OUT/pytest_integration.log | tail -30
