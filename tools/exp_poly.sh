for n in 0 1 2 3; do
  SF_BUILD_ONLY=attention_tc.cu SF_NVCC_EXTRA="-DSF_EXP2_POLY_PER4=$n" python -m streamformer_b200.build --force > /dev/null 2>&1
  echo "== poly per 4: $n"
  python tools/kernel_bench.py --only attn --reps 50 2>&1 | grep "spatial attention:"
done
SF_BUILD_ONLY=attention_tc.cu SF_NVCC_EXTRA="-DSF_EXP2_POLY_PER4=1" python -m streamformer_b200.build --force > /dev/null 2>&1
python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "spatial" 2>&1 | tail -3
