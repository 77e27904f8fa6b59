#!/bin/bash
TAG=${1:-r2e2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "background_download or isletrain_cli_with" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^OpenBLAS" $OUT/pytest.log | tail -5
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d["roofline"]["frac"])
PY
tail -3 $OUT/bench.err
