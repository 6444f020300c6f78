#!/bin/bash
# 4-GPU check (gpurun --gpus 4): strong-scaling C3 bench lines, both couplings
mkdir -p gpurun_out
for coupling in halo pipelined; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 --coupling $coupling 2> gpurun_out/bench_N4_C3_$coupling.err | grep '^{' > gpurun_out/bench_N4_C3_$coupling.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_N4_C3_$coupling.json").read().strip().splitlines()[-1])
    print("$coupling", "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "LB", d["config"]["lower_bound_after_last_step"], d.get("bound_check"), d["clocks"])
except Exception as e:
    print("$coupling FAILED", e)
PY
  tail -3 gpurun_out/bench_N4_C3_$coupling.err
done
