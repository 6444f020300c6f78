timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "peak pool\|objects of size" | tail -25
