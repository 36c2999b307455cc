"""Ad-hoc robustness sweep: 1-layer model vs the numpy oracle over odd batch / frame counts (ragged GEMM
tiles on every tile shape, partial 16-row temporal tiles, bf16 and fp16).  python tools/fuzz_shapes.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import streamformer_oracle as O  # noqa: E402  (checker only)
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # noqa: E402


def rel_rms(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()))


worst = 0.0
for (B, T, dt) in [(3, 16, torch.bfloat16), (5, 7, torch.bfloat16), (7, 16, torch.bfloat16), (9, 3, torch.float16),
                   (3, 16, torch.float16), (11, 16, torch.bfloat16), (2, 33, torch.bfloat16)]:
    ocfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(ocfg, seed=B * 100 + T, style="stress")
    px = O.make_pixels(B, T, ocfg, seed=B * 100 + T)
    ref = O.forward(w, ocfg, px)
    model = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=1, enable_causal_temporal=True))
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items() if k in model.state_dict()}
    model.load_state_dict(sd, strict=False)
    model = model.to("cuda", dt).eval()
    with torch.no_grad():
        out = model(torch.from_numpy(px).cuda())
    r1 = rel_rms(out.last_hidden_state.float().cpu().numpy(), ref["last_hidden_state"])
    r2 = rel_rms(out.pooler_output.float().cpu().numpy(), ref["pooler_output"])
    tol = 2e-2 if dt == torch.bfloat16 else 3e-3
    print(f"B={B} T={T} {dt}: last_hidden_state {r1:.3e} pooler_output {r2:.3e} (tol {tol})", flush=True)
    assert r1 <= tol and r2 <= tol
    worst = max(worst, r1 / tol, r2 / tol)
print("ok, worst fraction of tolerance", round(worst, 3))
