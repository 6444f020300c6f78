mkdir -p gpurun_out
t() { timeout 200 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', round(d['lb_ms'],4))"; }
{
t 64 4096 32 potts f32 wavefront_max_ctas=1,wavefront_warps_per_cta=1
t 64 4096 32 potts f32 wavefront_warps_per_cta=1
t 64 4096 32 potts f32 wavefront_warps_per_cta=4
t 64 1024 16 trunclin f32 wavefront_max_ctas=1,wavefront_warps_per_cta=1
t 64 1024 16 trunclin f32 wavefront_warps_per_cta=1
t 64 1024 16 trunclin f32 wavefront_warps_per_cta=2
t 64 1024 16 trunclin f32 wavefront_warps_per_cta=4,wavefront_tile_cols=4
t 64 1024 16 trunclin f32 wavefront_warps_per_cta=1,wavefront_tile_cols=4
t 4096 4096 32 potts f32 wavefront_warps_per_cta=1,wavefront_ctas_per_sm=2
t 4096 4096 32 potts f32 wavefront_warps_per_cta=1,wavefront_ctas_per_sm=3
t 2048 4096 32 potts f32 wavefront_warps_per_cta=1
t 1024 4096 32 potts f32 wavefront_warps_per_cta=1
t 512 4096 32 potts f32 wavefront_warps_per_cta=1
} > gpurun_out/step_times.txt 2>&1
cat gpurun_out/step_times.txt
