# one GPU call: bench lines, ncu launch list + full captures of the pass kernel (results -> gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
for wl in C2 C3 C4 C5h; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "ms/step", round(d["ms_per_step"], 3), "value", "%.3g" % d["value"], "frac", round(d["roofline"]["frac"], 3), "e2e", "%.3g" % d["e2e"]["value"], "cpu", d["cpu_baseline"] and "%.3g" % d["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("$wl FAILED", e)
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavefront_lean -s 4 -c 1 -f -o gpurun_out/prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavefront_lean -s 4 -c 1 -f -o gpurun_out/prof_c3 python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out | tail -20
for t in "C2" "C3"; do timeout 120 python scripts/trace_steps.py $t 2>&1 | grep -v "first 48\|increases\|first 24"; done > gpurun_out/trace_steps_final.txt 2>&1
