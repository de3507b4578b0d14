#!/bin/bash
# A/B of the condensed solver's grid size (CTAs per SM) on the benchmark shape: stage timings of bench.py
OUT=gpurun_out; mkdir -p $OUT
for c in 24 32 48 64 128 100000; do
  MST_COLS_CTAS_PER_SM=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/ab_$c.json 2> $OUT/ab_$c.err
  python - <<PY
import json
d=json.loads(open("$OUT/ab_$c.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels_ms"]
print("ctas/SM $c: step %.3f solver %.3f sampler %.3f" % (d["ms_per_step"], [v for n,v in k.items() if n.startswith("condensed")][0], k["sample_collide_cull_kernel"]))
PY
done
