#!/bin/bash
# ncu launch list of one device-resident step of the c2 bench (the end-to-end leg is skipped)
TAG=${1:-rX}; OUT=gpurun_out/$TAG; mkdir -p $OUT
ISLE_BENCH_SKIP_E2E=1 timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 1250 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
wc -l $OUT/launches.csv; tail -2 $OUT/ncu_list.log | cut -c1-300
