set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "spatial or decode" 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_step.csv python tools/one_step.py > gpurun_out/s71_a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tcgen05 -o gpurun_out/r2f_gemm_full -f python tools/one_step.py > gpurun_out/s71_b.log 2>&1
ncu -i gpurun_out/r2f_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2f_gemm_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2f_gemm_full.ncu-rep
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:spatial_attn_row -c 1 -o gpurun_out/r2f_spatial_row -f python tools/one_step.py > gpurun_out/s71_c.log 2>&1
SF_STREAM_GRAPH=0 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:temporal_decode_direct -c 1 -o gpurun_out/r2f_decode_direct -f python tools/stream_steps.py --steps 34 --profile-from 33 > gpurun_out/s71_d.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_smoke.py 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "temporal_decode_kernel and 3-33-33 and direct" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_spatial_attention and 5-193 and row" 2>&1 | tail -6
ls -la gpurun_out/ | grep r2f
