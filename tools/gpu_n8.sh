#!/bin/bash
# the north-star run: PubMed-shaped c3 (8 x 1.025 M docs, 141k vocab, k = 2000) document-sharded over 8 B200s
TAG=${1:-r2n8}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/smi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tests/multi_gpu_check.py c3m > $OUT/multi_check_c3m.log 2>&1; echo "multi_check c3m rc=$?"; grep "multi_gpu_check\|FAIL" $OUT/multi_check_c3m.log | tail -4
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 $TR --master-port 29534 bench.py --config c3s --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err; echo "c3 sharded rc=$?"
grep -h "NVLS\|nranks\|comm 0x" $OUT/bench_c3_n$N.err | grep "rank 0\|NVLS" | head -8 > $OUT/nccl_c3_n$N.txt
cut -c1-900 $OUT/bench_c3_n$N.json; grep -v "NCCL INFO\|OMP_NUM\|^\*\*\*\|^W0" $OUT/bench_c3_n$N.err | tail -3
ISLE_BENCH_SKIP_E2E=1 timeout 900 $TR --master-port 29535 bench.py --config c3s --gpus $N --steps 1 --warmup 1 --no-cpu-baseline --opt ks_row_shard=0 > $OUT/bench_c3_n${N}_replicated.json 2> $OUT/bench_c3_n${N}_replicated.err; echo "c3 replicated rc=$?"
cut -c1-300 $OUT/bench_c3_n${N}_replicated.json
timeout 600 $TR --master-port 29536 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_n$N.json 2> $OUT/bench_c2_n$N.err; echo "c2 weak rc=$?"
cut -c1-300 $OUT/bench_c2_n$N.json
