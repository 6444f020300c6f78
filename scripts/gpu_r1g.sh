mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "lean or potts_negative or negative or launch_shapes" 2>&1 | grep -v "peak pool\|objects of size" | tail -5 ) > gpurun_out/pytest_lean.log 2>&1
tail -5 gpurun_out/pytest_lean.log
t() { timeout 200 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', d['lb'][-1])"; }
{
t 64 4096 32 potts f32 wavefront_max_ctas=1,wavefront_lean=0
t 64 4096 32 potts f32 wavefront_max_ctas=1,wavefront_lean=1
t 64 1024 16 trunclin f32 wavefront_max_ctas=1,wavefront_lean=0
t 64 1024 16 trunclin f32 wavefront_max_ctas=1,wavefront_lean=1
t 1024 1024 16 trunclin f32 wavefront_lean=0
t 1024 1024 16 trunclin f32 wavefront_lean=1
t 4096 4096 32 potts f32 wavefront_lean=0
t 4096 4096 32 potts f32 wavefront_lean=1
} > gpurun_out/lean_times.txt 2>&1
cat gpurun_out/lean_times.txt
