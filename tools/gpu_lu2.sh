#!/bin/bash
TAG=${1:-lu2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/lu_probe.py > $OUT/${TAG}_lu_probe.log 2>&1
echo "lu_probe rc=$?"; tail -7 $OUT/${TAG}_lu_probe.log | cut -c1-250 | head -5
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
bash tools/gpu_lu_ncu.sh ${TAG} | head -2
timeout 600 python tools/other_configs.py > $OUT/${TAG}_other.log 2>&1
echo "other rc=$?"; tail -8 $OUT/${TAG}_other.log
