#!/bin/bash
# One GPU-box pass: GPU test tier, benchmark line, probes, then the ncu captures of the same bench command.
# usage: tools/gpu_check.sh [tag]   (outputs under gpurun_out/<tag>_*)
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/${TAG}_clocks.csv 2>/dev/null &
SMI=$!
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 600 $OUT/${TAG}_bench_reference.json
timeout 120 python tools/latency_probe.py > $OUT/${TAG}_latency.json 2> $OUT/${TAG}_latency.err; cat $OUT/${TAG}_latency.json
timeout 120 python tools/fp64_peak.py > $OUT/${TAG}_fp64.json 2>&1; cat $OUT/${TAG}_fp64.json
kill $SMI
CMD="python bench.py --traj 262144 --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
timeout 300 $CMD > $OUT/${TAG}_ncu_plain2.log 2>&1 && \
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum --clock-control none --import-source on -k "regex:onepass_kernel|sample_collide_cull_kernel|condensed_cols_kernel" -s 6 -c 6 -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu rc=$?"
ls -la $OUT | tail -20
