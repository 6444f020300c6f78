# rows / split variants: parity tests + pass times (one gpurun call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_variant or rows_variant" 2>&1 | tail -5 > gpurun_out/pytest_rows.log
cat gpurun_out/pytest_rows.log
for t in wavefront_split=1 wavefront_rows=1; do
for g in "1024 1024 16 trunclin" "8 8192 16 trunclin" "1024 64 16 trunclin" "4096 4096 32 potts"; do
  timeout 120 python scripts/time_grid.py $g f32 $t | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], d['tune'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'GB/s', round(d['GBps_fwd']))"
done; done 2>&1 | tee gpurun_out/rows_times.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wavefront_split -s 3 -c 1 -f -o gpurun_out/prof_split_8x8192 python scripts/time_grid.py 8 8192 16 trunclin f32 wavefront_split=1 > gpurun_out/ncu_split.log 2>&1
