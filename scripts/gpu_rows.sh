mkdir -p gpurun_out
for g in "1024 1024 16 trunclin f32" "1024 1024 16 trunclin f64" "4096 4096 32 potts f32" "4096 4096 32 potts f64 wavefront_rows=1" "4096 4096 32 potts f64 wavefront_rows=0" "4096 4096 8 potts f32" "4096 4096 8 potts f32 wavefront_rows=0" "2048 2048 16 potts f32" "4096 4096 64 potts f32" "1024 1024 32 trunclin f32"  "1024 1024 32 trunclin f32 wavefront_rows=0"; do
  timeout 120 python scripts/time_grid.py $g | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'lb', round(d['lb_ms'],4), 'GB/s', round(d['GBps_fwd']))"
done 2>&1 | tee gpurun_out/rows_times2.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
