mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "wavefront or stripe or golden" 2>&1 | grep -v "peak pool\|objects of size" | tail -15 ) > gpurun_out/pytest_wave.log 2>&1
tail -4 gpurun_out/pytest_wave.log
b() { # name workload tune
  timeout 300 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu-baseline --tune "$3" > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step", round(d["ms_per_step"], 3), "pass ms", round(d["roofline"]["mean_launch_ms"], 3), "frac", round(d["roofline"]["frac"], 3), "LB", d["config"]["lower_bound_after_last_step"], d["clocks"])
except Exception as e:
    print("$1 FAILED", e)
PY
}
b c2_nw1 C2 wavefront_warps_per_cta=1
b c2_def C2 ""
b c2_k4 C2 wavefront_tile_cols=4
b c3_nw1 C3 wavefront_warps_per_cta=1
b c3_def C3 ""
b c3_nw2 C3 wavefront_warps_per_cta=2
for t in "C2 wavefront_warps_per_cta=1" "C2 wavefront_tile_cols=8" "C2 wavefront_tile_cols=4" "C3 wavefront_warps_per_cta=1" "C3 wavefront_tile_cols=4"; do
  timeout 120 python scripts/trace_pass.py $t 2>&1 | tail -22
done > gpurun_out/traces.txt 2>&1
