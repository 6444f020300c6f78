#!/bin/bash
# quick single-GPU check: the tests named on the command line (default: the whole GPU suite), then pass times of the headline grids
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -25 | tee gpurun_out/pytest_quick.log
for cfg in "1024 1024 16 trunclin" "4096 4096 32 potts" ; do python scripts/time_grid.py $cfg f32 2>&1 | tail -2; done | tee gpurun_out/quick_times.txt
