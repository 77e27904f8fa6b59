#!/bin/bash
# whole GPU suite + bench (+ optional reference arm)
TAG=${1:-r2p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -v "^OpenBLAS" $OUT/pytest_gpu.log | tail -12
timeout 400 python bench.py --steps 3 --warmup 3 ${BENCH_FLAGS} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
print(d["stage_ms_per_step"])
print(d.get("cpu_baseline"))
PY
tail -3 $OUT/bench.err
