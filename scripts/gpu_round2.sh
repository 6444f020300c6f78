# round-2 evidence in one GPU call: ncu --set full captures of the split wavefront kernel on C2 and C3, launch list of the default
# bench command (results -> gpurun_out/)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_split -s 4 -c 1 -f -o gpurun_out/prof_r02_c2 python scripts/time_grid.py 1024 1024 16 trunclin f32 > gpurun_out/ncu_r02_c2.log 2>&1; tail -1 gpurun_out/ncu_r02_c2.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavefront_split -s 4 -c 2 -f -o gpurun_out/prof_r02_c3 python scripts/time_grid.py 4096 4096 32 potts f32 > gpurun_out/ncu_r02_c3.log 2>&1; tail -1 gpurun_out/ncu_r02_c3.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_r02_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_r02_launch.log 2>&1; tail -8 gpurun_out/launches_r02_default.csv | cut -c1-260
