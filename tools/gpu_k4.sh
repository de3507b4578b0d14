#!/bin/bash
TAG=${1:-k4}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/latency_probe.py > $OUT/${TAG}_latency.json 2> $OUT/${TAG}_latency.err; cat $OUT/${TAG}_latency.json; tail -3 $OUT/${TAG}_latency.err
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_collision.py -q -x 2>&1 | tail -4
timeout 300 python tools/k4_probe.py > $OUT/${TAG}_k4.txt 2>&1; cat $OUT/${TAG}_k4.txt
CMD="python tools/k4_probe.py 262144 1"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sample_collide_cull_kernel" -s 1 -c 1 -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
