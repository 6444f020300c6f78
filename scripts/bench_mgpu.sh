N=${1:-2}
for c in halo pipelined; do
echo "== N=$N coupling=$c"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --coupling $c 2>&1 | grep -v "^W\|^\*\*\*\|peak pool\|objects of size" | tail -3
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
