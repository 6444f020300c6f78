t() { timeout 100 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', d['lb'][-1])"; }
for f in 1 3 5 7; do
t 1024 1024 16 trunclin f32 wavefront_lean=$f
t 4096 4096 32 potts f32 wavefront_lean=$f
done
