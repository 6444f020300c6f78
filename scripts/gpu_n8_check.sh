#!/bin/bash
# 8-GPU check of a round (gpurun --gpus 8): multi-rank parity tests at world 8, then the strong-scaling C3 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "8" 2>&1 | tail -15 > gpurun_out/pytest_multirank_N8.log
cat gpurun_out/pytest_multirank_N8.log
for coupling in halo pipelined; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --coupling $coupling 2> gpurun_out/bench_N8_C3_$coupling.err | grep '^{' > gpurun_out/bench_N8_C3_$coupling.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_N8_C3_$coupling.json").read().strip().splitlines()[-1])
    print("$coupling", "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "LB", d["config"]["lower_bound_after_last_step"], d.get("bound_check"), d["clocks"])
except Exception as e:
    print("$coupling FAILED", e)
PY
  tail -3 gpurun_out/bench_N8_C3_$coupling.err
done
