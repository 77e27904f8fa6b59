#!/bin/bash
# spread of the sharded-vs-single Lloyd assignment mismatch at c3m (k = 320) with the peer-memory collectives and with NCCL
TAG=${1:-r2noise}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for P in 1 0 1 0; do
  i=$((i+1))
  ISLE_P2P=$P timeout 300 $TR --master-port 2971$i tests/multi_gpu_check.py c3m > $OUT/c3m_p2p${P}_$i.log 2>&1; echo "p2p=$P rc=$?"; grep "multi_gpu_check\|FAIL" $OUT/c3m_p2p${P}_$i.log | tail -3
done
