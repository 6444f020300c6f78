# full GPU parity suite + smoke + the default bench line and the reference arm (results -> gpurun_out/)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "peak pool\|objects of size" | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "peak pool\|objects of size" | tail -4 ) > gpurun_out/smoke.log 2>&1
cat gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/bench_reference.json; cat gpurun_out/bench_reference.json
