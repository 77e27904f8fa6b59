#!/bin/bash
# ncu --set full captures of the tensor-core panel kernels (c3 shard, k = 2000) and of the stage-F / k-means kernels (c2)
TAG=${1:-rX}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'panel_tc_kernel' -s 2000 -c 2 -f -o $OUT/panel_tc_full \
    python bench.py --config c3s --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_panel.log 2>&1; echo "ncu panel rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'assign_full|count_members|pp_skinny|dist_tc_kernel|project_kernel' -c 6 -f -o $OUT/kmeans_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_kmeans.log 2>&1; echo "ncu kmeans rc=$?"
ls -la $OUT
