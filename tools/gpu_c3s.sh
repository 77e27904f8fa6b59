#!/bin/bash
# one c3-shard (k = 2000) bench line
TAG=${1:-rX}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python bench.py --config c3s --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_c3s.json 2> $OUT/bench_c3s.err; echo "c3s rc=$?"
cat $OUT/bench_c3s.json | cut -c1-600; tail -5 $OUT/bench_c3s.err
