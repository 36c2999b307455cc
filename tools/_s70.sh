timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s70_bench.json 2> gpurun_out/s70_bench.err
tail -c 600 gpurun_out/s70_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/s70_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('value','ms_per_step','vs_baseline','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['clocks'])
        for k,v in d['configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','frames_per_s','ms_per_step_median','ms_per_step_mean','frac_of_floor','host_us_per_step')})
        print(d['cpu_baseline'])
PY
