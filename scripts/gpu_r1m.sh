( timeout 900 python -m pytest tests -m gpu -x -q -k "not wavefront and not stripe" 2>&1 | grep -v "peak pool\|objects of size" | tail -3 )
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 2>/dev/null | tee gpurun_out/bench_C4.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 ms/step', d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'launch ms', d['roofline']['mean_launch_ms'], 'e2e %.4g' % d['e2e']['value'])"
