for g in 0 4 2 1; do
timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu-baseline --tune level_group=$g 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 level_group=$g ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'launch ms %.4f' % d['roofline']['mean_launch_ms'], 'LB', d['config'].get('lower_bound_after_last_step'))"
done
( timeout 600 python -m pytest tests -m gpu -x -q -k "general_graph or colouring or degenerate or edge_order" 2>&1 | grep -v "peak pool\|objects of size" | tail -2 )
