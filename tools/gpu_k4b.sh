#!/bin/bash
TAG=${1:-k4b}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/k4_probe.py > $OUT/${TAG}_k4.txt 2>&1; cat $OUT/${TAG}_k4.txt
CMD="python tools/k4_probe.py 262144 1"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sample_collide_cull_kernel|condensed_cols_kernel" -s 0 -c 2 -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
