( timeout 900 python -m pytest tests -m gpu -x -q -k "lean or launch_shapes or stripe or full_size" 2>&1 | grep -v "peak pool\|objects of size" | tail -3 )
t() { timeout 100 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', d['lb'][-1])"; }
t 1024 1024 16 trunclin f32
t 4096 4096 32 potts f32
t 4096 4096 32 potts f32 wavefront_cluster=4
