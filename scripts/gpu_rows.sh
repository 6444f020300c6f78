# rows / split variants: parity tests + pass times (one gpurun call)
mkdir -p gpurun_out
timeout 300 python scripts/dbg_split.py 2>&1 | tail -20
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rows_variant or split_variant" 2>&1 | tail -15 > gpurun_out/pytest_rows.log
cat gpurun_out/pytest_rows.log
for t in wavefront_split=1 wavefront_rows=1; do
for g in "1024 1024 16 trunclin" "8 8192 16 trunclin" "1024 64 16 trunclin" "4096 4096 32 potts"; do
  timeout 120 python scripts/time_grid.py $g f32 $t | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'GB/s', round(d['GBps_fwd']))"
done; done 2>&1 | tee gpurun_out/rows_times.txt
