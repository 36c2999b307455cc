import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from tests.test_backward_model_gpu import _models, _loss
from oracle import streamformer_oracle as O
from streamformer_b200 import ops
orig = ops.gate_backward
def patched(dx, y, gate, dgate):
    out = orig(dx, y, gate, dgate)
    tg = torch.tanh(gate.double())
    ref = float((dx.double() * y.double()).sum() * (1 - tg * tg))
    sabs = float((dx.double() * y.double()).abs().sum())
    print("gate_backward kernel", float(dgate), "fp64 on same tensors", ref, "sum|terms|", sabs, "|dx||y|", float(dx.double().norm() * y.double().norm()))
    return out
ops.gate_backward = patched
for (L, B, T, dt, seed) in [(1, 1, 3, torch.float32, 63), (2, 2, 4, torch.bfloat16, 61)]:
    ocfg, ours, ref = _models(L, False, seed, dtype=dt)
    px = torch.from_numpy(O.make_pixels(B, T, ocfg, seed=seed)).cuda()
    g = torch.Generator(device="cpu").manual_seed(seed)
    w_pool = torch.randn(B, T, 768, generator=g).cuda() * 0.1
    w_tok = torch.randn(B, T, 196, 768, generator=g).cuda() * 0.01
    _loss(ours(px), w_pool, w_tok).backward()
    _loss(ref(px.float()), w_pool, w_tok).backward()
    for l in range(L):
        print("layer", l, "ours", float(ours.encoder.layer[l].temporal_attention_gating.grad), "ref", float(ref.encoder.layer[l].temporal_attention_gating.grad))
