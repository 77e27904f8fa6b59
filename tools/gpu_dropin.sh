#!/bin/bash
TAG=${1:-r2drop}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python tools/dropin_timing.py --config c2 --docs ${DOCS:-100000} --ngpus ${NG:-1} > $OUT/dropin.json 2> $OUT/dropin.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/dropin.json"))
print(d["workload"], d["host_cores"])
for k,v in d.items():
    if isinstance(v,dict): print(k, {kk: (round(vv,3) if not isinstance(vv,dict) else None) for kk,vv in v.items()})
PY
tail -3 $OUT/dropin.err
