mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "dense" 2>&1 | grep -v "peak pool\|objects of size" | tail -30 ) > gpurun_out/pytest_dense.log 2>&1
tail -5 gpurun_out/pytest_dense.log
for args in "256 2048 64 dense f32" "256 2048 64 dense f32 wavefront_ctas_per_sm=1" "512 2048 64 dense f32" "256 1024 32 dense f32" "128 1024 64 dense f64"; do
  timeout 300 python scripts/time_grid.py $args 2>&1 | tail -1
done
sed -i 's/lp.set_tuning(wavefront_trace=1)/lp.set_tuning(wavefront_trace=1, wavefront_max_ctas=64)/' scripts/trace_c2.py
python scripts/trace_c2.py > gpurun_out/trace_c2_64.txt 2>&1; tail -22 gpurun_out/trace_c2_64.txt
