N=${1:-8}
mkdir -p gpurun_out
echo "== N=$N default (C2, halo)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench_N${N}_C2_halo.json | cut -c1-400
echo "== N=$N C3 halo"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3 --workload C3 2>&1 | grep '^{' | tee gpurun_out/bench_N${N}_C3_halo.json | cut -c1-400
echo "== N=$N C3 pipelined"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 5 --warmup 3 --workload C3 --coupling pipelined 2>&1 | grep '^{' | tee gpurun_out/bench_N${N}_C3_pipelined.json | cut -c1-400
