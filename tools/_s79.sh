timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step') if k in d}, d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['attention_block']['frac_of_burst_peak_executed'], d['attention_block']['ms_per_layer'])
"
