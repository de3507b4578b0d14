#!/bin/bash
TAG=${1:-lu}
OUT=gpurun_out
mkdir -p $OUT
for v in 0 1; do
  MST_LU_VARIANT=$v timeout 300 python tools/lu_probe.py 2 > $OUT/${TAG}_v$v.log 2>&1
  echo "variant $v rc=$?"; head -2 $OUT/${TAG}_v$v.log | cut -c1-120
done
timeout 1200 python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_edge_cases.py tests/test_gpu_onepass.py tests/test_gpu_fullsize.py -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
bash tools/gpu_lu_ncu.sh ${TAG} | head -2
