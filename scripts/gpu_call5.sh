mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "dense" 2>&1 | grep -v "peak pool\|objects of size" | tail -30 ) > gpurun_out/pytest_dense.log 2>&1
tail -5 gpurun_out/pytest_dense.log
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; tail -c 2500 gpurun_out/bench_C4.json; tail -3 gpurun_out/bench_C4.err
timeout 900 python bench.py --workload C5h --steps 5 --warmup 3 > gpurun_out/bench_C5h.json 2> gpurun_out/bench_C5h.err; tail -c 2500 gpurun_out/bench_C5h.json; tail -3 gpurun_out/bench_C5h.err
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err; tail -c 2500 gpurun_out/bench_C5.json; tail -3 gpurun_out/bench_C5.err
