# separates the per-step time S of a row (8 x W grids) from the row-to-row latency (H x 64 grids) of the split kernel
for g in "8 8192" "8 16384" "64 8192" "1024 64" "2048 64" "1024 1024" "512 4096"; do
  timeout 120 python scripts/time_grid.py $g 16 trunclin f32 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4))"
done
for g in "8 8192" "1024 64" "512 4096"; do
  timeout 120 python scripts/time_grid.py $g 32 potts f32 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4))"
done
