# rows variant: parity tests + pass times against the lean kernel (one gpurun call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rows_variant" 2>&1 | tail -15 > gpurun_out/pytest_rows.log
cat gpurun_out/pytest_rows.log
for t in "wavefront_rows=1" "wavefront_rows=0"; do
  timeout 120 python scripts/time_grid.py 1024 1024 16 trunclin f32 $t
  timeout 120 python scripts/time_grid.py 4096 4096 32 potts f32 $t
done 2>&1 | tee gpurun_out/rows_times.txt
