#!/bin/bash
# parity tests of the grid kernels, then the chain-shape experiment (scripts/exp_chain.sh) and the headline grids
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_chain.log
bash scripts/exp_chain.sh 2>&1 | tee gpurun_out/exp_chain.txt
python scripts/time_grid.py 4096 4096 32 potts f32 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['grid'], 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4))" | tee -a gpurun_out/exp_chain.txt
