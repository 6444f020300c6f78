for wl in C2 C3; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl ms/step', d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['time_to_bound'])"
done
