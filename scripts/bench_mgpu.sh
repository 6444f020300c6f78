N=${1:-2}
mkdir -p gpurun_out
for c in halo pipelined; do
echo "== N=$N coupling=$c"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --coupling $c 2>&1 | grep -v "^W\|^\*\*\*\|peak pool\|objects of size" | tail -3 | tee gpurun_out/bench_N${N}_$c.json | cut -c1-1800
done
echo "== N=$N C3 halo"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload C3 2>&1 | grep -v "^W\|^\*\*\*\|peak pool\|objects of size" | tail -2 | tee gpurun_out/bench_N${N}_C3_halo.json | cut -c1-1200
echo "== N=$N stripe tests"
timeout 600 python -m pytest tests -m gpu -q -k stripe 2>&1 | grep -v "peak pool\|objects of size" | tail -3
