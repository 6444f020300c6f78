timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed"
python scripts/trace_c2.py | tail -6
for wl in C2 C3; do
echo "== $wl"; timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms/step',d['ms_per_step'],'pass ms', d['roofline']['mean_launch_ms'],'frac', d['roofline']['frac'],'upd/s %.3g'%d['value'],'e2e %.3g'%d['e2e']['value'])
"
done
