"""Per-kernel-class GPU time of one training step (forward + native backward), CUDA events around every launch.
    python tools/train_profile.py [--batch 8]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200 import _native as N  # noqa: E402
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--layers", type=int, default=12)
a = ap.parse_args()
torch.manual_seed(0)
m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=a.layers, enable_causal_temporal=True)).to("cuda", torch.bfloat16).train()
x = torch.randn(a.batch, 16, 3, 224, 224, device="cuda", dtype=torch.bfloat16)


def step():
    for p in m.parameters():
        p.grad = None
    out = m(x)
    (out.pooler_output.float().sum() * 1e-3).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1))
N.profile(1)
step()
prof = N.profile_collect()
N.profile(0)
print(json.dumps({k: {"ms": round(v["ms"], 3), "launches": v["launches"], "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)} for k, v in prof.items() if v["launches"]}, indent=1))
