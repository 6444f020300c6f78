#!/bin/bash
# multi-GPU evidence of a round (run with gpurun --gpus N): multi-rank parity tests, then strong-scaling bench lines on C3
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_multirank_N$N.log
cat gpurun_out/pytest_multirank_N$N.log
for coupling in halo pipelined; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --coupling $coupling > gpurun_out/bench_N${N}_C3_$coupling.json 2> gpurun_out/bench_N${N}_C3_$coupling.err
  tail -c 3000 gpurun_out/bench_N${N}_C3_$coupling.json; tail -5 gpurun_out/bench_N${N}_C3_$coupling.err
done
