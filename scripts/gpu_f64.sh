mkdir -p gpurun_out
for wl in C2 C3; do
timeout 600 python bench.py --workload $wl --dtype f64 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_${wl}_f64.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl f64 ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], 'launch ms %.3f' % d['roofline']['mean_launch_ms'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'])"
done
