# one GPU call: parity tests, bench lines, ncu launch list + one full capture of the pass kernel (results -> gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "peak pool\|objects of size" | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for wl in C2 C3; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  tail -c 1500 gpurun_out/bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/bench_reference.json
cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavefront_pass -s 4 -c 1 -f -o gpurun_out/prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavefront_pass -s 4 -c 1 -f -o gpurun_out/prof_c3 python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out
