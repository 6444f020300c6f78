t() { timeout 100 python scripts/time_grid.py "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4))"; }
t 1024 1024 16 trunclin f32 wavefront_tile_cols=4
t 1024 1024 16 trunclin f32 wavefront_tile_cols=8
t 4096 4096 32 potts f32 wavefront_tile_cols=8
t 4096 4096 32 potts f32 wavefront_ctas_per_sm=3
t 4096 4096 16 potts f32
t 4096 4096 16 potts f32 wavefront_tile_cols=4
