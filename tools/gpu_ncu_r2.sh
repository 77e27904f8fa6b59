#!/bin/bash
# round-2 evidence: block-KS tests with the Jacobi eig, a bench line, the launch list of one c2 step, ncu --set full of the SpMM kernels
TAG=${1:-r2ncu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "block_ks or end_to_end or c2_eigensolver" > $OUT/pytest_ks.log 2>&1; echo "pytest rc=$?"; grep -v "^OpenBLAS" $OUT/pytest_ks.log | tail -4
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"])
print({k:round(v,2) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
ISLE_BENCH_SKIP_E2E=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 1400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"; wc -l $OUT/launches.csv
ISLE_BENCH_SKIP_E2E=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spmm_head_i8|spmm_gather_bfp' -s 40 -c 4 -o $OUT/spmm \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la $OUT | tail -5
