timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "pool or golden or stream or head" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py --only cfg3 2>&1 | tail -1
timeout 300 python tools/kernel_bench.py --only head --reps 30 2>&1 | grep -v "^{" | tail -5
