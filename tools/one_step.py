"""One cfg2 forward step (B=8, T=16, 12 layers, bf16) bracketed by cudaProfilerStart/Stop after warm-up, for
    ncu --profile-from-start off ... python tools/one_step.py [--train]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--train", action="store_true")
a = ap.parse_args()
torch.manual_seed(0)
m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=a.layers, enable_causal_temporal=True)).to("cuda", torch.bfloat16)
m = m.train() if a.train else m.eval()
xs = [torch.randn(a.batch, a.frames, 3, 224, 224, device="cuda", dtype=torch.bfloat16) for _ in range(3)]


def step(i):
    if a.train:
        for p in m.parameters():
            p.grad = None
        (m(xs[i % 3]).pooler_output.float().sum() * 1e-3).backward()
    else:
        with torch.no_grad():
            m(xs[i % 3])


for i in range(3):
    step(i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
