"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
2 layers, B=2, T=4 one-shot forward (CTA-pair GEMM tiles are not reached at this size), a 3-step
stream, and one large-M GEMM of every epilogue flavour so the TMA-store epilogue is exercised too."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200 import ops  # noqa: E402
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # noqa: E402

torch.manual_seed(0)
m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=2, enable_causal_temporal=True)).to("cuda", torch.bfloat16).eval()
with torch.no_grad():
    x = torch.randn(2, 4, 3, 224, 224, device="cuda")
    out = m(x)
    cache = m.new_kv_cache(2, max_frames=4)
    for t in range(3):
        m(x[:, t:t + 1], past_key_values=cache)
    # CTA-pair tiles + TMA-store epilogue: M = 9600 (ragged last M tile), all epilogue flavours
    M, D = 9600 + 40, 768
    a = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(3 * D, D, device="cuda", dtype=torch.bfloat16) * 0.05
    bias = torch.randn(3 * D, device="cuda")
    st = ops.rowstats(a)
    ops.gemm(a, w, bias=bias, ln_stats=st, ln_colsum=torch.randn(3 * D, device="cuda"), ln_eps=1e-6)
    ops.gemm(a, w, bias=bias, act=1)
    res = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    so = torch.empty(ops.gemm_stats_parts(M, D), M, 2, device="cuda")
    ops.gemm(a, w[:D], bias=bias[:D], residual=res, gate=torch.tensor([0.3], device="cuda"), stats_out=so, out=res)
    torch.cuda.synchronize()
print("ok", float(out.pooler_output.float().abs().mean()))
