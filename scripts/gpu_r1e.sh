mkdir -p gpurun_out
# single free-running band (no predecessor, no polling): where does a step's time go?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_pass -s 4 -c 1 -f -o gpurun_out/prof_band_c3 python scripts/time_grid.py 4 4096 32 potts f32 wavefront_warps_per_cta=1 > gpurun_out/ncu_band_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_pass -s 4 -c 1 -f -o gpurun_out/prof_band_c2 python scripts/time_grid.py 8 1024 16 trunclin f32 wavefront_warps_per_cta=1 > gpurun_out/ncu_band_c2.log 2>&1
tail -3 gpurun_out/ncu_band_c3.log gpurun_out/ncu_band_c2.log
ls -la gpurun_out/*.ncu-rep
