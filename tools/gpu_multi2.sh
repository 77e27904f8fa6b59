#!/bin/bash
TAG=${1:-r2y}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
for cfg in c1 c3m; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py $cfg > $OUT/multi_check_$cfg.log 2>&1; echo "multi_check $cfg rc=$?"; grep -v "^OpenBLAS\|^W0\|^\*\*\*" $OUT/multi_check_$cfg.log | tail -6
done
ISLE_KS_ROW_SHARD=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tests/multi_gpu_check.py c1 > $OUT/multi_check_c1_repl.log 2>&1; echo "multi_check replicated rc=$?"; tail -2 $OUT/multi_check_c1_repl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; cut -c1-700 $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
