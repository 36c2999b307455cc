timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "gemm or stream or cfg3 or golden or oracle or head" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py --only cfg3 2>&1 | tail -1
SF_GEMM_BN64=0 timeout 300 python tools/bench_configs.py --only cfg3 2>&1 | tail -1
