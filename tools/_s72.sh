timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "decode" 2>&1 | tail -2
timeout 120 python tools/decode_bench.py
timeout 300 python tools/bench_configs.py --only cfg3 2>&1 | tail -1
