#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu full capture of the SpMM kernels.
# usage: tools/gpu_round.sh <tag> [quick]
TAG=${1:-rX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
if [ "$2" != "quick" ]; then
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
cat $OUT/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 5000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spmm_' -s 40 -c 6 -f -o $OUT/spmm_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
