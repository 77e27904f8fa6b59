#!/bin/bash
# peer-to-peer collectives (coll.cu): self-test vs NCCL + sharded-vs-single parity, then the weak-scaling bench with P2P on / off
TAG=${1:-r2p2p}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tests/multi_gpu_check.py c1 > $OUT/multi_check_c1.log 2>&1; echo "multi_check c1 rc=$?"; grep "multi_gpu_check\|FAIL\|rror\|latency" $OUT/multi_check_c1.log | tail -5
timeout 600 $TR --master-port 29542 tests/multi_gpu_check.py c3m > $OUT/multi_check_c3m.log 2>&1; echo "multi_check c3m rc=$?"; grep "multi_gpu_check\|FAIL\|rror" $OUT/multi_check_c3m.log | tail -5
timeout 600 python -m pytest tests/test_integration_gpu.py -q -k "multi_gpu_context or two_gpus" > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; grep -v "^OpenBLAS" $OUT/pytest_multi.log | tail -4
for P in 1 0; do
  timeout 400 $TR --master-port 2955$P bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --opt p2p=$P > $OUT/bench_c2_n${N}_p2p$P.json 2> $OUT/bench_c2_n${N}_p2p$P.err; echo "bench p2p=$P rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_c2_n${N}_p2p$P.json").read().strip().splitlines()[-1])
    s=d["stage_ms_per_step"]
    print("p2p=$P", {k:round(d[k],2) for k in ("value","ms_per_step")}, "allreduce", round(s.get("allreduce",0),2), "allgather", round(s.get("allgather",0),2), "ks_op", round(s["ks_op"],2), "pp", round(s["pp_round"],2), "lloyd", round(s["lloyd_iter"],2))
except Exception as e: print("no line", e)
PY
  grep -v "NCCL INFO\|OMP_NUM\|^\*\*\*\|^W0" $OUT/bench_c2_n${N}_p2p$P.err | tail -3
done
