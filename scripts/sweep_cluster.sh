t() { timeout 100 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4))"; }
for c in 1 2 4 8; do t 1024 1024 16 trunclin f32 wavefront_cluster=$c; t 4096 4096 32 potts f32 wavefront_cluster=$c; done
t 2048 2048 32 potts f32
t 4096 4096 8 potts f32
t 2048 2048 64 potts f32
t 2048 2048 32 trunclin f32
