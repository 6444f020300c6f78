# compute-sanitizer memcheck on small wavefront cases (anisotropic + uniform instances of the split kernel, clusters, several bands per CTA)
for c in "70 48 16 trunclin" "37 32 32 potts wavefront_cluster=4" "20 24 16 potts wavefront_max_ctas=1" "37 53 32 potts"; do
  timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_case.py $c 2>&1 | grep -v "peak pool\|objects of size" | tail -4
done
