mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "peak pool\|objects of size" | tail -8 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
./scripts/repro/pingpong > gpurun_out/pingpong.txt 2>&1; cat gpurun_out/pingpong.txt
python scripts/trace_c2.py > gpurun_out/trace_c2.txt 2>&1; tail -22 gpurun_out/trace_c2.txt
for wl in C2 C3; do for tune in "" "wavefront_tile_cols=4"; do
echo "== $wl $tune"; timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --tune "$tune" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms/step',d['ms_per_step'],'pass ms', d['roofline']['mean_launch_ms'],'frac', d['roofline']['frac'],'upd/s %.3g'%d['value'],'e2e %.3g'%d['e2e']['value'])
"
done; done
