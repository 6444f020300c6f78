#!/bin/bash
# strong-scaling C3 bench lines on N GPUs (gpurun --gpus N): per-pass halo and exact pipelined coupling
N=${1:-2}
mkdir -p gpurun_out
for coupling in halo pipelined; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --coupling $coupling > gpurun_out/bench_N${N}_C3_$coupling.json 2> gpurun_out/bench_N${N}_C3_$coupling.err
  grep '^{' gpurun_out/bench_N${N}_C3_$coupling.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], d['config']['stripe_coupling'][:24], 'ms/iter', round(d['ms_per_step'],3), 'G upd/s', round(d['value']/1e9,2), 'bound', d['config']['lower_bound_after_last_step'], d.get('bound_check'))"
done
