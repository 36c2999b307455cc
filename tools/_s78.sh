cp streamformer_b200/lib/libsf_tl.so streamformer_b200/lib/libstreamformer_b200.so
SF_SPATIAL_SKEW=2000 timeout 40 python tools/kernel_bench.py --only attn --reps 1 2>&1 | grep -E "^tile" | head -22
