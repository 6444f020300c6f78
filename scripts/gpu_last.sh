mkdir -p gpurun_out
for wl in C2 C3; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$wl.json").read().strip().splitlines()[-1])
print("$wl", "ms/step", round(d["ms_per_step"], 3), "value", "%.4g" % d["value"], "launch", round(d["roofline"]["mean_launch_ms"],4), "frac", round(d["roofline"]["frac"], 3), "e2e", "%.4g" % d["e2e"]["value"], "cpu", d["cpu_baseline"] and "%.3g" % d["cpu_baseline"]["value"])
PY
done
