#!/bin/bash
# round-end evidence on 1 x B200: smoke, the whole GPU suite, the default bench line (with the CPU baseline), the ncu launch list of one step
TAG=${1:-r2final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v OpenBLAS | tail -2
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^OpenBLAS" $OUT/pytest_gpu.log | tail -4
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","steps","warmup")}, d["e2e"]["value"], d["roofline"]["frac"], d["tensor"]["frac"])
print({k:round(v,2) for k,v in d["stage_ms_per_step"].items() if v>0.25})
print(d.get("cpu_baseline",{}).get("value"), d["clocks"])
PY
tail -2 $OUT/bench.err
ISLE_BENCH_SKIP_E2E=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 4200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"; wc -l $OUT/launches.csv
