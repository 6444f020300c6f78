for f in 17 33; do
  timeout 120 python scripts/dbg_lean.py 24 40 16 trunclin fwd wavefront_max_ctas=1,wavefront_lean=$f 2>&1 | grep "theta differs\|a differs"
done
timeout 120 python scripts/dbg_lean.py 17 40 16 trunclin fwd wavefront_max_ctas=1 2>&1 | grep "theta differs\|a differs\|first a"
timeout 120 python scripts/dbg_lean.py 24 40 16 trunclin fwd wavefront_tile_cols=4 2>&1 | grep "theta differs\|a differs\|first a"
