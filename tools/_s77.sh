for nrot in 4 1; do
  echo "== nrot $nrot"
  SF_KB_NROT=$nrot SF_SPATIAL_SKEW=2000 timeout 40 python tools/kernel_bench.py --only attn --reps 50 2>&1 | grep -E "spatial attention|temporal"
  SF_KB_NROT=$nrot SF_SPATIAL_ROW=0 timeout 40 python tools/kernel_bench.py --only attn --reps 50 2>&1 | grep -E "spatial attention:"
done
