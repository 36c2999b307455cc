"""A few streaming steps (BASELINE configs[2]: B=4, one frame per step, 12 layers) for profilers:
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:. --csv python tools/stream_steps.py --steps 40 --profile-from 38
runs `--steps` appends and brackets the last ones with cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=34)
ap.add_argument("--profile-from", type=int, default=32)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--layers", type=int, default=12)
a = ap.parse_args()
torch.manual_seed(0)
m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=a.layers, enable_causal_temporal=True, num_frames=64))
m = m.to("cuda", torch.bfloat16).eval()
x = torch.randn(a.batch, 1, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
cache = m.new_kv_cache(a.batch, max_frames=64)
with torch.no_grad():
    for s in range(a.steps):
        if s == a.profile_from:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        m(x, past_key_values=cache)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", cache.get_seq_length())
