#!/bin/bash
TAG=${1:-r2k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
# head trace while the tail runs beside it (fork mode)
ISLE_HEAD8_TRACE=$OUT/trace_fork.txt timeout 600 python tools/spmm_check.py --config c2 --no-ref --reps 1 --density-ppm 12000 --opt spmm_tail_pipe=0 > $OUT/run.log 2>&1
python tools/head8_trace.py $OUT/trace_fork.txt > $OUT/trace_fork_summary.txt; grep "steady\|launch" $OUT/trace_fork_summary.txt | tail -12
# launch list of one operator application (serial mode), device time per kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|colmax|pack|ysplit|unpack' -s 14 -c 14 --csv --log-file $OUT/launches.csv python tools/spmm_check.py --config c2 --no-ref --only-head-serial --reps 2 --density-ppm 12000 > $OUT/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
ik=rows[h].index("Kernel Name"); iv=rows[h].index("Metric Value")
for r in rows[h+1:]: print(f"{float(r[iv])/1000:9.1f} us  {r[ik][:90]}")
PY
