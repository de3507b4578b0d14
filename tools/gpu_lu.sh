#!/bin/bash
# pivoted banded solver: all GPU tests, its throughput probe, the other-configuration timings
TAG=${1:-lu}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/lu_probe.py > $OUT/${TAG}_lu_probe.log 2>&1
echo "lu_probe rc=$?"; tail -7 $OUT/${TAG}_lu_probe.log
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 600 python tools/other_configs.py > $OUT/${TAG}_other.log 2>&1
echo "other rc=$?"; cat $OUT/${TAG}_other.log | tail -12
