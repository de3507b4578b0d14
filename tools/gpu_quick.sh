#!/bin/bash
# quick GPU pass: selected tests, a bench line, one ncu capture of the pipeline kernel
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_collision.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py tests/test_gpu_host_pipeline.py -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.1f M/s  %.3f ms/step" % (d["value"]/1e6, d["ms_per_step"]))
print(json.dumps(d["roofline"]["kernels_ms"], indent=1))
print("e2e", d["e2e"]["value"]/1e6, d["e2e"]["pol_matrix_f32_wire"]["value"]/1e6)
PY
CMD="python bench.py --traj 262144 --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum --clock-control none --import-source on -k "regex:onepass_kernel|sample_collide_cull_kernel|condensed_cols_kernel" -s 6 -c 6 -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu rc=$?"
