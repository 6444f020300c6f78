timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dbg2.py 70 45 16 trunclin f32 2>&1 | grep -v "peak pool\|objects of size" | tail -12
timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dbg2.py 37 53 32 potts f32 wavefront_cluster=4 2>&1 | grep -v "peak pool\|objects of size" | tail -6
