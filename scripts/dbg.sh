cd /root/repo
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "test_cuda_grid_api_reproduces_reference_fixture and wavefront-grid_potts" 2>&1 | grep -v "^$" | head -60
