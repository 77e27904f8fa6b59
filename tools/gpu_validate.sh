#!/bin/bash
# parity tests + both bench arms (no profiler)
TAG=${1:-rX}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
