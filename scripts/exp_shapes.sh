for g in "8 8192" "8 16384" "1024 64" "2048 64" "64 1024" "128 1024"; do
  for t in "wavefront_rows=1" "wavefront_rows=0"; do timeout 120 python scripts/time_grid.py $g 16 trunclin f32 $t | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4))"; done
done
