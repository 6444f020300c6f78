mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q -k "lean or launch_shapes or stripe" 2>&1 | grep -v "peak pool\|objects of size" | tail -8 ) > gpurun_out/pytest_lean.log 2>&1
tail -4 gpurun_out/pytest_lean.log
t() { timeout 100 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', d['lb'][-1])"; }
{
t 1024 1024 16 trunclin f32 wavefront_cluster=1
t 1024 1024 16 trunclin f32 wavefront_cluster=8
t 1024 1024 16 trunclin f32 wavefront_cluster=8,wavefront_tile_cols=4
t 4096 4096 32 potts f32 wavefront_cluster=1
t 4096 4096 32 potts f32 wavefront_cluster=8
t 4096 4096 32 potts f32 wavefront_cluster=8,wavefront_ctas_per_sm=3
} > gpurun_out/cluster_times.txt 2>&1
cat gpurun_out/cluster_times.txt
for t in "C2 wavefront_cluster=8" "C3 wavefront_cluster=8"; do
  timeout 120 python scripts/trace_pass.py $t 2>&1 | tail -26
done > gpurun_out/traces_cluster.txt 2>&1
grep -A8 "lag between" gpurun_out/traces_cluster.txt
