timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --coupling halo 2>&1 | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'], d['e2e']['phases_ms'])"
