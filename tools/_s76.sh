timeout 90 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "spatial" 2>&1 | tail -3
for sk in 0 2000 3500 5000; do
  echo "== skew $sk"
  SF_SPATIAL_SKEW=$sk timeout 40 python tools/kernel_bench.py --only attn --reps 50 2>&1 | grep "spatial attention"
done
