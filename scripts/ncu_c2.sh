# one full ncu capture of the wavefront pass kernel on C2 + the launch list
ncu --set full --clock-control none --import-source on -k regex:wavefront_pass -s 4 -c 1 -o gpurun_out/prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log
