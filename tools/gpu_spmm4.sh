#!/bin/bash
TAG=${1:-r2j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for pipe in 0 1; do
echo "== spmm_tail_pipe=$pipe"
timeout 600 python tools/spmm_check.py --config c2 --no-ref --density-ppm 12000 --i8 1 --head-max 8192 --opt spmm_tail_pipe=$pipe > $OUT/spmm_check_$pipe.log 2>&1; echo "spmm_check rc=$?"
grep -v "^OpenBLAS" $OUT/spmm_check_$pipe.log | grep "per product" | tail -40
done
# launch list of one operator application (serial mode), device time per kernel
ISLE_BENCH_SKIP_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|colmax|pack|ysplit|unpack' -s 60 -c 24 --csv --log-file $OUT/launches.csv python tools/spmm_check.py --config c2 --no-ref --only-head-serial --reps 3 --density-ppm 12000 > $OUT/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]: print(f"{float(r[iv])/1000:9.1f} us  {r[ik][:90]}")
PY
