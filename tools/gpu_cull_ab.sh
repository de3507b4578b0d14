#!/bin/bash
# A/B of the culling sampler's tile scheduling (device tickets vs fixed stride) on the benchmark shape and with yaw
OUT=gpurun_out; mkdir -p $OUT
for m in fixed tickets; do
  if [ $m = fixed ]; then export MST_CULL_FIXED_STRIDE=1; else unset MST_CULL_FIXED_STRIDE; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/abt_$m.json 2> $OUT/abt_$m.err
  python - <<PY
import json
d=json.loads(open("$OUT/abt_$m.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels_ms"]
print("$m: step %.3f solver %.3f sampler %.3f" % (d["ms_per_step"], [v for n,v in k.items() if n.startswith("condensed")][0], k["sample_collide_cull_kernel"]))
PY
  timeout 300 python tools/k4_probe.py 2>&1 | sed -n 2p
done
unset MST_CULL_FIXED_STRIDE
timeout 900 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_fullsize.py tests/test_gpu_host_pipeline.py tests/test_gpu_edge_cases.py -q -x 2>&1 | tail -2
