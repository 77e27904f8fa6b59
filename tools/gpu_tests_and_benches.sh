#!/bin/bash
# GPU visit: parity tests (no -x), c2 bench, c3-shard bench
TAG=${1:-rX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 ${BENCH_FLAGS} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 420 python bench.py --config c3s --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_c3s.json 2> $OUT/bench_c3s.err; echo "c3s rc=$?"
cat $OUT/bench_c3s.json; tail -5 $OUT/bench_c3s.err
