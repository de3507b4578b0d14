#!/bin/bash
# multi-GPU pass on one box: gather correctness (N ranks), then the benchmark at N ranks
N=${1:-2}; TAG=${2:-m}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > $OUT/${TAG}_n${N}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/${TAG}_n${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_n${N}_bench.json 2> $OUT/${TAG}_n${N}_bench.err
echo "bench rc=$?"; tail -5 $OUT/${TAG}_n${N}_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_n${N}_bench.json") if l.startswith("{")][-1])
    print("N=%d value %.1f M/s  %.3f ms/step" % (d["n_gpus"], d["value"]/1e6, d["ms_per_step"]))
    print(json.dumps(d.get("gather_modes"), indent=1)); print(d.get("strong_scaling")); print(d["checks"])
except Exception as e: print("no line", e)
PY
