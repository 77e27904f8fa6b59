#!/bin/bash
TAG=${1:-rX}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python tools/panel_check.py --big > $OUT/panel_check.log 2>&1; echo "panel_check rc=$?"; cat $OUT/panel_check.log | tail -30
