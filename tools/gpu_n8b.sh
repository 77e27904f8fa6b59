#!/bin/bash
# 8 x B200 with the peer-memory collectives: sharded-vs-single parity at k = 320, the north-star c3 step, the c2 weak-scaling step
TAG=${1:-r2n8b}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29633 tests/multi_gpu_check.py c3m > $OUT/multi_check_c3m.log 2>&1; echo "multi_check c3m rc=$?"; grep "multi_gpu_check\|FAIL\|latency" $OUT/multi_check_c3m.log | tail -4
timeout 500 $TR --master-port 29634 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_n$N.json 2> $OUT/bench_c2_n$N.err; echo "c2 weak rc=$?"
timeout 700 $TR --master-port 29635 bench.py --config c3s --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err; echo "c3 sharded rc=$?"
python - <<PY
import json
for f in ("bench_c2_n$N.json", "bench_c3_n$N.json"):
    try:
        d=json.loads(open("$OUT/"+f).read().strip().splitlines()[-1]); s=d["stage_ms_per_step"]
        print(f, {k:round(d[k],2) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"] or 0), "p2p/step", d["run"].get("p2p_collectives_per_step"),
              {k:round(v,1) for k,v in s.items() if v>0.5})
    except Exception as e: print(f, "no line", e)
PY
grep -v "NCCL INFO\|OMP_NUM\|^\*\*\*\|^W0" $OUT/bench_c3_n$N.err | tail -3
