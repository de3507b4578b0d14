#!/bin/bash
# The pivoted banded solver on a GPU box: A/B of its pivot search (MST_LU_VARIANT), the throughput probe on five
# shapes, the GPU test tier, the other-configuration timings, then one ncu capture (tools/gpu_lu_ncu.sh).
# usage: tools/gpu_lu.sh [tag]   (outputs under gpurun_out/<tag>_*)
TAG=${1:-lu}
OUT=gpurun_out
mkdir -p $OUT
for v in 0 1; do
  MST_LU_VARIANT=$v timeout 300 python tools/lu_probe.py 2 > $OUT/${TAG}_v$v.log 2>&1
  echo "variant $v rc=$?"; head -2 $OUT/${TAG}_v$v.log | cut -c1-120
done
timeout 600 python tools/lu_probe.py > $OUT/${TAG}_lu_probe.log 2>&1
echo "lu_probe rc=$?"; tail -1 $OUT/${TAG}_lu_probe.log | cut -c1-300
timeout 120 python tools/lu_small_shapes.py > $OUT/${TAG}_small.log 2>&1; echo "small shapes rc=$?"
timeout 120 python tools/lu_list_probe.py > $OUT/${TAG}_list.log 2>&1; echo "list mode rc=$?"; cat $OUT/${TAG}_list.log
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python tools/other_configs.py > $OUT/${TAG}_other.log 2>&1
echo "other rc=$?"; tail -8 $OUT/${TAG}_other.log
bash tools/gpu_lu_ncu.sh ${TAG} | head -2
