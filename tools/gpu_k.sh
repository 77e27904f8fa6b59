#!/bin/bash
# run a pytest -k selection on the GPU
TAG=${1:-r2}; K=${2:-ingest}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "$K" > $OUT/pytest_k.log 2>&1; echo "pytest rc=$?"
grep -v "^OpenBLAS" $OUT/pytest_k.log | tail -${3:-30}
