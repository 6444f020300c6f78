mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stripes" 2>&1 | tail -15 > gpurun_out/pytest_stripes.log
cat gpurun_out/pytest_stripes.log
timeout 120 python scripts/time_stripe_modes.py C3 > gpurun_out/time_stripe_modes.txt 2>&1
timeout 120 python scripts/time_stripe_modes.py C2 >> gpurun_out/time_stripe_modes.txt 2>&1
cat gpurun_out/time_stripe_modes.txt
