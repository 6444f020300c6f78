mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "peak pool\|objects of size" | tail -4 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "peak pool\|objects of size" | tail -2 )
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 > gpurun_out/bench_C4.json 2>/dev/null; python -c "
import json
d = json.loads(open('gpurun_out/bench_C4.json').read().strip().splitlines()[-1]); print('C4 ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'])"
timeout 600 python bench.py > gpurun_out/bench_default.json 2>/dev/null; python -c "
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print('default ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'], d['clocks'], d['gpu_launches'])"
