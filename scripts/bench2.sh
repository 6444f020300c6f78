timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
