python scripts/time_setup.py 2>&1 | tail -4
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "peak pool\|objects of size" | tail -3 )
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 ms/step', d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['e2e']['phases_ms'])"
timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 ms/step', d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['e2e']['phases_ms'])"
