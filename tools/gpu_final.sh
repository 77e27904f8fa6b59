#!/bin/bash
# Round-end style visit: parity tests, both bench arms, ncu launch list of one device-resident step.
TAG=${1:-rX}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
ISLE_BENCH_SKIP_E2E=1 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
