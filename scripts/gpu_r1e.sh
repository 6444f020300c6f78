mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_lean -s 4 -c 1 -f -o gpurun_out/prof_lband_c3 python scripts/time_grid.py 4 4096 32 potts f32 wavefront_lean=1 > gpurun_out/ncu_lband_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_lean -s 4 -c 1 -f -o gpurun_out/prof_lband_c2 python scripts/time_grid.py 8 1024 16 trunclin f32 wavefront_lean=1 > gpurun_out/ncu_lband_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
