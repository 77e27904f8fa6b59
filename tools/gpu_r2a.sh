#!/bin/bash
# round 2, visit a: new large-k parity tests, then the whole GPU suite, then both bench arms
TAG=${1:-r2a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/smi.txt; nproc >> $OUT/smi.txt
timeout 900 python -m pytest tests/test_gpu_largek.py -m gpu -q --durations=10 > $OUT/pytest_largek.log 2>&1; echo "largek rc=$?" | tee -a $OUT/pytest_largek.log
tail -40 $OUT/pytest_largek.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_largek.py > $OUT/pytest_rest.log 2>&1; echo "rest rc=$?" | tee -a $OUT/pytest_rest.log
tail -5 $OUT/pytest_rest.log
timeout 400 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cut -c1-1500 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
