"""Per-kernel checks of the backward-pass ops through the C ABI against torch fp32 autograd of the
same op on the same (16-bit rounded) inputs.  Floating-point kernels: tolerance = a few ulps of the 16-bit output."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from streamformer_b200 import ops
    return ops


def _close(got, want, dtype, what, scale=1.0):
    got, want = got.float(), want.float()
    tol = (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * scale
    err = float((got - want).abs().max())
    ref = float(want.abs().max())
    if ref < 1e-5:      # exactly-zero gradients (e.g. dq / dk of a one-key softmax): absolute check
        assert err <= 1e-5, f"{what}: max err {err:.4g} against a zero reference"
        return
    rel = float((got - want).norm() / want.norm().clamp_min(1e-30))
    assert err <= tol * ref * 2 and rel <= tol, f"{what}: max err {err:.4g} (ref max {ref:.4g}), rel-rms {rel:.4g}"


@pytest.mark.parametrize("M,N", [(25088, 768), (100, 8), (1571, 2304), (64, 64), (196 * 3, 3072)])
def test_transpose(M, N):
    ops = _ops()
    x = torch.randn(M, N, device=DEV).bfloat16()
    out = ops.transpose(x)
    Mp = (M + 7) // 8 * 8
    assert out.shape == (N, Mp)
    assert torch.equal(out[:, :M], x.t()) and float(out[:, M:].abs().sum()) == 0.0


def test_wgrad_via_transpose_and_gemm():
    """dW = dY^T . X as gemm(a = dY^T, w = X^T): the contraction runs over M (padded with zeros to a multiple of 8)."""
    ops = _ops()
    M, O, I = 196 * 5 + 0, 768, 256
    dY = (torch.randn(M, O, device=DEV) * 0.1).bfloat16()
    X = torch.randn(M, I, device=DEV).bfloat16()
    G = ops.gemm(ops.transpose(dY), ops.transpose(X))
    _close(G, dY.float().t() @ X.float(), torch.bfloat16, "wgrad")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,O,I", [(25088, 768, 768), (1571, 2304, 768), (588, 768, 3072), (100, 128, 256), (4100, 1536, 768), (33, 32, 768),
                                   (6272, 768, 40)])
def test_wgrad_tcgen05_mn_major(dtype, M, O, I):
    """G = dY^T . X with both operands read MN-major in place (ragged M / O / I tiles, more splits than row blocks)."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(M + O + I)
    dY = (torch.randn(M, O, generator=g) * 0.1).to(DEV, dtype)
    X = torch.randn(M, I, generator=g).to(DEV, dtype)
    parts = ops.wgrad(dY, X)
    assert parts.dtype == torch.float32 and parts.shape[1:] == (O, I)
    want = dY.float().t() @ X.float()
    got = parts.sum(0)
    err = float((got - want).abs().max())
    assert err <= 2e-3 * float(want.abs().max()) + 1e-4, f"max err {err:.4g} (ref max {float(want.abs().max()):.4g}, splits {parts.shape[0]})"
    dW = ops.wfold_finish_partials(parts, dtype, torch.float32)
    assert float((dW - got).abs().max()) <= 1e-5 * max(1.0, float(got.abs().max()))


@pytest.mark.parametrize("M,N", [(25088, 768), (777, 2304), (3, 64)])
def test_colsum(M, N):
    ops = _ops()
    x = torch.randn(M, N, device=DEV).bfloat16()
    got = ops.colsum(x)
    want = x.float().sum(0)
    assert float((got - want).abs().max()) <= 1e-3 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("residual", [True, False])
def test_ln_backward(dtype, residual):
    ops = _ops()
    M, D = 1571, 768
    x = (torch.randn(M, D, device=DEV) * 2 + 0.3).to(dtype)
    dn = torch.randn(M, D, device=DEV).to(dtype)
    dres = torch.randn(M, D, device=DEV).to(dtype) if residual else None
    got = ops.ln_backward(x, dn, 1e-6, dres)
    xr = x.float().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (D,), eps=1e-6)
    y.backward(dn.float())
    want = xr.grad + (dres.float() if residual else 0)
    _close(got, want, dtype, "ln backward")


@pytest.mark.parametrize("row_map,T,S", [(0, 1, 1), (2, 4, 49)])
def test_ln_affine_backward(row_map, T, S):
    ops = _ops()
    B, D = 3, 768
    M = B * T * S if row_map else 1000
    x = (torch.randn(M, D, device=DEV) * 1.5).bfloat16()
    gamma = (1 + 0.1 * torch.randn(D, device=DEV)).float()
    beta = (0.1 * torch.randn(D, device=DEV)).float()
    dy = torch.randn(M, D, device=DEV).bfloat16()          # rows in the OUTPUT order
    dgamma = torch.zeros(D, device=DEV)
    dbeta = torch.zeros(D, device=DEV)
    got = ops.ln_affine_backward(x, dy, gamma, 1e-6, dgamma, dbeta, row_map, T, S)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (D,), gr, br, eps=1e-6)
    if row_map == 2:   # (b,n,t) -> (b,t,n)
        y = y.view(B, S, T, D).permute(0, 2, 1, 3).reshape(M, D)
    y.backward(dy.float())
    _close(got, xr.grad, torch.bfloat16, "ln affine backward dx")
    assert float((dgamma - gr.grad).abs().max()) <= 2e-3 * float(gr.grad.abs().max())
    assert float((dbeta - br.grad).abs().max()) <= 2e-3 * float(br.grad.abs().max())


@pytest.mark.parametrize("act", [1, 2])
def test_gelu_backward(act):
    ops = _ops()
    a = (torch.randn(1000, 3072, device=DEV) * 2).bfloat16()
    dh = torch.randn(1000, 3072, device=DEV).bfloat16()
    ar = a.float().requires_grad_(True)
    h = torch.nn.functional.gelu(ar, approximate="none" if act == 1 else "tanh")
    h.backward(dh.float())
    a2, d2 = a.clone(), dh.clone()
    ops.gelu_backward_(a2, d2, act)
    _close(a2, h.detach(), torch.bfloat16, "gelu forward")
    _close(d2, ar.grad, torch.bfloat16, "gelu backward")


def test_gate_backward():
    ops = _ops()
    dx = torch.randn(3000, 768, device=DEV).bfloat16()
    y = torch.randn(3000, 768, device=DEV).bfloat16()
    gate = torch.tensor([0.37], device=DEV)
    dgate = torch.zeros(1, device=DEV)
    dy = ops.gate_backward(dx, y, gate, dgate)
    tg = torch.tanh(gate)
    _close(dy, dx.float() * tg, torch.bfloat16, "gate dy")
    want = float((dx.float() * y.float()).sum() * (1 - tg * tg))
    assert abs(float(dgate) - want) <= 1e-3 * abs(want) + 1e-2


def test_wfold_finish():
    ops = _ops()
    O, I = 2304, 768
    G = torch.randn(O, I, device=DEV).bfloat16()
    W = (torch.randn(O, I, device=DEV) * 0.05)
    gamma = (1 + 0.1 * torch.randn(I, device=DEV)).float()
    beta = (0.1 * torch.randn(I, device=DEV)).float()
    db = torch.randn(O, device=DEV).float()
    Wp = (W * gamma).bfloat16()
    dgamma, dbeta = torch.zeros(I, device=DEV), torch.zeros(I, device=DEV)
    dW = ops.wfold_finish(G, torch.float32, Wp, gamma, beta, db, dgamma, dbeta)
    Wr = Wp.float() / gamma
    assert float((dW - (G.float() * gamma + db[:, None] * beta[None, :])).abs().max()) <= 1e-4
    wg, wb = (G.float() * Wr).sum(0), (db[:, None] * Wr).sum(0)
    assert float((dgamma - wg).abs().max()) <= 2e-3 * float(wg.abs().max())
    assert float((dbeta - wb).abs().max()) <= 2e-3 * float(wb.abs().max())
    plain = ops.wfold_finish(G, torch.bfloat16)
    assert torch.equal(plain, G)


def test_embed_table_grads_and_rowperm():
    ops = _ops()
    B, T, S, D = 3, 5, 49, 768
    dx = torch.randn(B * S * T, D, device=DEV).bfloat16()       # rows (b, n, t)
    pos = torch.zeros(S, D, device=DEV)
    ops.embed_table_grad(dx, B, T, S, 0, pos)
    v = dx.float().view(B, S, T, D)
    assert float((pos - v.sum((0, 2))).abs().max()) <= 1e-3
    tidx = torch.tensor([0, 0, 1, 2, 2], device=DEV, dtype=torch.int32)
    tm = torch.zeros(3, D, device=DEV)
    ops.embed_table_grad(dx, B, T, S, 1, tm, tidx)
    per_t = v.sum((0, 1))
    want = torch.stack([per_t[0] + per_t[1], per_t[2], per_t[3] + per_t[4]])
    assert float((tm - want).abs().max()) <= 2e-3
    p = ops.rowperm(dx, 2, T, S)                                  # (b,n,t) -> (b,t,n)
    assert torch.equal(p, dx.view(B, S, T, D).permute(0, 2, 1, 3).reshape(-1, D))


def _attn_autograd(q, k, v, dout, scale, causal):
    q, k, v = [t.float().requires_grad_(True) for t in (q, k, v)]
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        L = s.shape[-1]
        s = s.masked_fill(~torch.tril(torch.ones(L, L, dtype=torch.bool, device=s.device)), float("-inf"))
    o = torch.softmax(s, -1) @ v
    o.backward(dout.float())
    return o.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sites,T,causal", [(50, 16, True), (33, 6, True), (20, 16, False), (7, 24, True), (3, 128, True), (5, 1, True)])
def test_temporal_attention_backward(dtype, sites, T, causal):
    ops = _ops()
    H, D = 12, 768
    qkv = torch.randn(sites * T, 3 * D, device=DEV).to(dtype)
    dout = torch.randn(sites * T, D, device=DEV).to(dtype)
    x = qkv.view(sites, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    do = dout.view(sites, T, H, 64).permute(0, 2, 1, 3)
    o, dq, dk, dv = _attn_autograd(x[0], x[1], x[2], do, 0.125, causal)
    out = o.permute(0, 2, 1, 3).reshape(sites * T, D).to(dtype)
    got = ops.attention_backward(0, qkv, out, dout, sites, H, T, 1, causal, 0.125)
    want = torch.stack([dq, dk, dv], 0).permute(1, 3, 0, 2, 4).reshape(sites * T, 3 * D)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        _close(got[:, sl], want[:, sl], dtype, f"temporal {name}", scale=4.0)


@pytest.mark.parametrize("B,T,S", [(2, 4, 196), (1, 1, 49), (1, 3, 64), (2, 1, 196)])
def test_spatial_attention_backward(B, T, S):
    """Rows in the residual stream's (b, n, t) order (T_inner = T): token n of frame (b, t) at row (b*S + n)*T + t."""
    ops = _ops()
    dtype, H, D = torch.bfloat16, 12, 768
    frames = B * T
    qkv = torch.randn(B * S * T, 3 * D, device=DEV).to(dtype)
    dout = torch.randn(B * S * T, D, device=DEV).to(dtype)
    x = qkv.view(B, S, T, 3, H, 64).permute(3, 0, 2, 4, 1, 5).reshape(3, frames, H, S, 64)
    do = dout.view(B, S, T, H, 64).permute(0, 2, 3, 1, 4).reshape(frames, H, S, 64)
    o, dq, dk, dv = _attn_autograd(x[0], x[1], x[2], do, 0.125, False)
    out = o.view(B, T, H, S, 64).permute(0, 3, 1, 2, 4).reshape(B * S * T, D).to(dtype)
    got = ops.attention_backward(1, qkv, out, dout, frames, H, S, T, False, 0.125)
    want = torch.stack([dq, dk, dv], 0).view(3, B, T, H, S, 64).permute(1, 4, 2, 0, 3, 5).reshape(B * S * T, 3 * D)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        _close(got[:, sl], want[:, sl], dtype, f"spatial {name}", scale=4.0)


def test_pool_attention_backward():
    ops = _ops()
    frames, H, S, D = 6, 12, 196, 768
    kv = torch.randn(frames * S, 2 * D, device=DEV).bfloat16()
    q = (torch.randn(D, device=DEV) * 0.3).float()
    dout = torch.randn(frames, D, device=DEV).bfloat16()
    dq = torch.zeros(D, device=DEV)
    got = ops.pool_attention_backward(kv, q, dout, frames, H, S, dq)
    kvr = kv.float().requires_grad_(True)
    qr = q.clone().requires_grad_(True)
    k = kvr[:, :D].view(frames, S, H, 64)
    v = kvr[:, D:].view(frames, S, H, 64)
    s = torch.einsum("hd,fnhd->fhn", qr.view(H, 64), k)
    o = torch.einsum("fhn,fnhd->fhd", torch.softmax(s, -1), v).reshape(frames, D)
    o.backward(dout.float())
    _close(got, kvr.grad, torch.bfloat16, "pool dkv", scale=2.0)
    assert float((dq - qr.grad).abs().max()) <= 5e-3 * float(qr.grad.abs().max())
