( timeout 300 python -m pytest tests -m gpu -x -q -k "lean or launch_shapes or stripe" 2>&1 | grep -v "peak pool\|objects of size" | tail -3 )
for t in "C2 wavefront_cluster=8" "C3 wavefront_cluster=8"; do timeout 120 python scripts/trace_steps.py $t 2>&1 | grep "consumer end-of-step\|step ns\|first 48"; done
t() { timeout 100 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', d['lb'][-1])"; }
t 1024 1024 16 trunclin f32 wavefront_cluster=8
t 1024 1024 16 trunclin f32 wavefront_cluster=1
t 4096 4096 32 potts f32 wavefront_cluster=8
t 4096 4096 32 potts f32 wavefront_cluster=1
