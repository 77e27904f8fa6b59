#!/bin/bash
# one ncu --set full capture of the int8 head kernel (both passes) on c2, serial mode
TAG=${1:-r2f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERN:-spmm_head_i8} -s ${SKIP:-4} -c ${COUNT:-2} -o $OUT/head8 \
  python tools/spmm_check.py --config c2 --only-head-serial --no-ref --reps 2 ${SPMM_ARGS:---density-ppm 6000 --head-max 8192} > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -5 $OUT/ncu.log; ls -la $OUT
