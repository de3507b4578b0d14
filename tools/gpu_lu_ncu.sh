#!/bin/bash
# ncu capture of the pivoted banded solver on config 3's shape (65 536 x 20 pieces)
TAG=${1:-luncu}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/lu_one.py <<PY
import numpy as np, torch, sys
sys.path.insert(0, ".")
import drone_path_planning_python_b200 as mst
rng = np.random.default_rng(5)
B, n, K = 65536, 20, 3
T = np.clip(rng.uniform(0.5, 2, (B, n)) * np.exp(rng.normal(size=(B, n))), 0.05, 5.0)
t = torch.as_tensor(np.concatenate([np.zeros((B, 1)), np.cumsum(T, 1)], 1), device="cuda")
wp = torch.as_tensor(np.cumsum(rng.normal(0, 0.3, (B, n + 1, K)), 1), device="cuda")
for _ in range(3):
    mst.solve_batch(wp, t, solver="banded_lu")
torch.cuda.synchronize()
PY
timeout 300 python /tmp/lu_one.py > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:banded_lu_kernel" -s 1 -c 1 -o $OUT/${TAG}_prof python /tmp/lu_one.py > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"

