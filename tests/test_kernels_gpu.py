"""Per-kernel parity: each hand-written sm_100a kernel, called through the C ABI (sf_op_*), against a
plain PyTorch fp32 reference of the same op on the same seeded inputs.

Tolerances: operands are bf16/fp16, accumulation fp32, output rounded once to bf16/fp16, so the
bound is one output ulp (2^-8 relative for bf16, 2^-11 for fp16) plus accumulation-order noise.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from streamformer_b200 import ops
    return ops


def _close(out, ref, dtype, what, scale=1.0):
    out = out.float()
    ref = ref.float()
    tol = (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * scale
    err = (out - ref).abs().max().item()
    bound = tol * max(ref.abs().max().item(), 1.0)
    rel_rms = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)).item()
    assert math.isfinite(err), f"{what}: non-finite output"
    assert err <= bound, f"{what}: max abs err {err:.4g} > {bound:.4g} (rel rms {rel_rms:.3g})"
    assert rel_rms <= tol, f"{what}: rel rms {rel_rms:.3g} > {tol:.3g}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64),        # one tile, one k-block
    (128, 128, 768),       # narrow-tile path
    (300, 768, 768),       # ragged M
    (25088, 768, 768),     # cfg2 out-proj shape (wide tiles, ~4 waves)
    (3136, 2304, 768),     # cfg1 QKV
    (1000, 3072, 768),     # fc1
    (1000, 768, 3072),     # fc2 (long K)
    (6, 768, 768),         # pooling head at B=1, T=6
    (200, 264, 72),        # ragged N and K (K % 64 != 0, N % 128 != 0)
])
def test_gemm_bias(dtype, M, N, K):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(1234 + M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV, dtype)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, dtype)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias
    _close(out, ref, dtype, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_gelu(act):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(7)
    M, N, K = 1024, 3072, 768
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias=bias, act=act)
    pre = a.float() @ w.float().t() + bias
    ref = torch.nn.functional.gelu(pre, approximate="none" if act == 1 else "tanh")
    _close(out, ref, torch.bfloat16, f"gemm+gelu act={act}")


def test_gemm_gated_residual_inplace():
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(8)
    M, N, K = 3136, 768, 768
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV, torch.bfloat16)
    gate = torch.tensor([0.7], device=DEV)
    ref = res.float() + math.tanh(0.7) * (a.float() @ w.float().t() + bias)
    buf = res.clone()
    out = ops.gemm(a, w, bias=bias, residual=buf, gate=gate, out=buf)  # in place on the residual
    _close(out, ref, torch.bfloat16, "gemm+gated residual")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,act", [(3136, 2304, 768, 0), (25088, 3072, 768, 1), (300, 768, 768, 0), (784, 2304, 768, 0),
                                       (1000, 3072, 768, 2)])
def test_gemm_folded_layernorm(dtype, M, N, K, act):
    """LayerNorm folded into the projection: A holds the raw rows (large mean, wide spread), W is
    pre-scaled by gamma, the epilogue applies rstd * (acc - mean * colsum) + (b + W.beta).  The row
    statistics come from ops.rowstats (one partial) — reference is LN in fp32 followed by the Linear."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(21 + M + N)
    x = (torch.randn(M, K, generator=g) * (0.5 + 3 * torch.rand(M, 1, generator=g)) + 2 * torch.randn(M, 1, generator=g))
    x[:, 7] += 20.0  # a massive-activation channel as ViT residual streams have
    x = x.to(DEV, dtype)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    gamma = (1 + 0.3 * torch.randn(K, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(K, generator=g)).to(DEV)
    wp = (w * gamma).to(dtype)
    colsum = wp.float().sum(1)
    biasp = bias + w @ beta
    stats = ops.rowstats(x)
    out = ops.gemm(x, wp, bias=biasp, act=act, ln_stats=stats, ln_colsum=colsum, ln_eps=1e-6)
    ref = torch.nn.functional.layer_norm(x.float(), (K,), gamma, beta, 1e-6) @ w.t() + bias
    if act:
        ref = torch.nn.functional.gelu(ref, approximate="none" if act == 1 else "tanh")
    _close(out, ref, dtype, f"gemm+folded LN {M}x{N}x{K} act={act}", scale=2.0)


@pytest.mark.parametrize("M,N,K", [(25088, 768, 768), (3136, 768, 3072), (784, 768, 768), (300, 768, 768)])
def test_gemm_stats_out_feeds_folded_layernorm(M, N, K):
    """A residual-epilogue GEMM writes partial (sum, sumsq) of its output rows; they must equal the
    statistics of the stored (rounded) rows, and drive the next GEMM's folded LayerNorm."""
    ops = _ops()
    dtype = torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(31 + M)
    a = torch.randn(M, K, generator=g).to(DEV, dtype)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, dtype)
    res = (torch.randn(M, N, generator=g) * 2 + 1).to(DEV, dtype)
    bias = torch.randn(N, generator=g).to(DEV)
    parts = ops.gemm_stats_parts(M, N)
    stats = torch.full((parts, M, 2), float("nan"), device=DEV)
    y = ops.gemm(a, w, bias=bias, residual=res, stats_out=stats)
    tot = stats.sum(0)
    yf = y.float()
    assert torch.allclose(tot[:, 0], yf.sum(1), rtol=1e-4, atol=1e-2), (tot[:, 0] - yf.sum(1)).abs().max()
    assert torch.allclose(tot[:, 1], yf.pow(2).sum(1), rtol=1e-4, atol=1e-2)
    # consumer
    N2 = 2304
    w2 = (torch.randn(N2, N, generator=g) * 0.05).to(DEV)
    gamma = (1 + 0.3 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(N, generator=g)).to(DEV)
    wp = (w2 * gamma).to(dtype)
    out = ops.gemm(y, wp, bias=w2 @ beta, ln_stats=stats, ln_colsum=wp.float().sum(1), ln_eps=1e-6)
    ref = torch.nn.functional.layer_norm(yf, (N,), gamma, beta, 1e-6) @ w2.t()
    _close(out, ref, dtype, "stats_out -> folded LN", scale=2.0)


def test_gemm_embed_stats_out():
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(41)
    B, T, S, K, N = 2, 4, 196, 768, 768
    M = B * T * S
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    pos = torch.randn(S, N, generator=g).to(DEV)
    time = torch.randn(16, N, generator=g).to(DEV)
    stats = torch.zeros(ops.gemm_stats_parts(M, N), M, 2, device=DEV)
    y = ops.gemm(a, w, row_map=1, T=T, S=S, pos=pos, time_emb=time, time_total=T, stats_out=stats).float()
    tot = stats.sum(0)
    assert torch.allclose(tot[:, 0], y.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[:, 1], y.pow(2).sum(1), rtol=1e-4, atol=1e-2)


def test_gelu_erf_epilogue_accuracy():
    """The x*sigmoid(poly) form of the erf GELU stays within the bf16 output rounding of the exact one."""
    ops = _ops()
    K, N, M = 64, 256, 4096
    a = torch.zeros(M, K, device=DEV, dtype=torch.float16)
    a[:, 0] = 1.0
    w = torch.zeros(N, K, device=DEV, dtype=torch.float16)
    xs = torch.linspace(-9, 9, M * N, device=DEV).view(M, N)
    bias = xs[0].clone()          # per-column offsets ...
    a[:, 1] = (xs[:, 0] - xs[0, 0]).to(torch.float16)   # ... plus a per-row shift through the GEMM
    w[:, 1] = 1.0
    pre = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias=bias, act=1).float()
    ref = torch.nn.functional.gelu(pre.double(), approximate="none")
    err = (out.double() - ref).abs()
    bound = ref.abs() * 2.0 ** -10 + 4e-5     # fp16 rounding (2^-11) + the fit's 1.8e-4 relative / 3.5e-5 absolute error
    assert bool((err <= bound).all()), (err - bound).max()


@pytest.mark.parametrize("B,T,S", [(2, 16, 196), (1, 6, 196), (3, 5, 49)])
def test_gemm_row_maps(B, T, S):
    """QKV projection writes (b,t,n) rows from (b,n,t) input; out-proj maps back with residual."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(9)
    M, K, N = B * T * S, 768, 256
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    full = (a.float() @ w.float().t())
    # (b,n,t) -> (b,t,n)
    out = ops.gemm(a, w, row_map=2, T=T, S=S)
    ref = full.view(B, S, T, N).permute(0, 2, 1, 3).reshape(M, N)
    _close(out, ref, torch.bfloat16, "row_map BNT->BTN")
    # (b,t,n) -> (b,n,t) with residual indexed by the output row
    res = torch.randn(M, N, generator=g).to(DEV, torch.bfloat16)
    out = ops.gemm(a, w, row_map=1, T=T, S=S, residual=res)
    ref = full.view(B, T, S, N).permute(0, 2, 1, 3).reshape(M, N) + res.float()
    _close(out, ref, torch.bfloat16, "row_map BTN->BNT + residual")


@pytest.mark.parametrize("T,time_len,time_off,time_total", [(16, 16, 0, 16), (6, 16, 0, 6), (24, 16, 0, 24), (1, 16, 5, 6), (2, 16, 30, 40)])
def test_gemm_patch_embed_epilogue(T, time_len, time_off, time_total):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(10)
    B, S, K, N = 2, 196, 768, 768
    M = B * T * S
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    pos = torch.randn(S, N, generator=g).to(DEV)
    time = torch.randn(time_len, N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias=bias, row_map=1, T=T, S=S, pos=pos, time_emb=time, time_total=time_total, time_off=time_off)
    y = (a.float() @ w.float().t() + bias).view(B, T, S, N) + pos.view(1, 1, S, N)
    if time_total <= time_len:
        tsel = time[time_off:time_off + T]
    else:
        stretched = torch.nn.functional.interpolate(time.t().unsqueeze(0), size=time_total, mode="nearest")[0].t()
        tsel = stretched[time_off:time_off + T]
    y = y + tsel.view(1, T, 1, N)
    ref = y.permute(0, 2, 1, 3).reshape(M, N)
    _close(out, ref, torch.bfloat16, "patch-embed epilogue")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,D", [(1000, 768), (7, 768), (64, 1536), (33, 64)])
def test_layernorm(dtype, M, D):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(11)
    x = (torch.randn(M, D, generator=g) * 3 + 1).to(DEV, dtype)
    gamma = torch.randn(D, generator=g).to(DEV)
    beta = torch.randn(D, generator=g).to(DEV)
    y = ops.layernorm(x, gamma, beta, 1e-6)
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-6)
    _close(y, ref, dtype, "layernorm")


def test_layernorm_rowmap():
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(12)
    B, T, S, D = 2, 5, 49, 768
    x = torch.randn(B * S * T, D, generator=g).to(DEV, torch.bfloat16)
    gamma = torch.randn(D, generator=g).to(DEV)
    beta = torch.randn(D, generator=g).to(DEV)
    y = ops.layernorm(x, gamma, beta, 1e-6, row_map=2, T=T, S=S)
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-6)
    ref = ref.view(B, S, T, D).permute(0, 2, 1, 3).reshape(-1, D)
    _close(y, ref, torch.bfloat16, "layernorm BNT->BTN")


@pytest.mark.parametrize("pix_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_im2col_matches_conv(pix_dtype):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(13)
    BT, Cc, H, W, P, D = 3, 3, 224, 224, 16, 64
    pix = torch.randn(BT, Cc, H, W, generator=g).to(DEV, pix_dtype)
    wconv = (torch.randn(D, Cc, P, P, generator=g) * 0.05).to(DEV)
    a = ops.im2col(pix, P, torch.bfloat16)
    ref = torch.nn.functional.conv2d(pix.float().to(torch.bfloat16).float(), wconv, stride=P).flatten(2).transpose(1, 2).reshape(-1, D)
    out = a.float() @ wconv.reshape(D, -1).t()
    assert torch.allclose(out, ref, atol=2e-3, rtol=1e-3), (out - ref).abs().max()


def _attn_ref(q, k, v, scale, mask=None):
    s = (q @ k.transpose(-2, -1)) * scale
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    return s.softmax(-1) @ v


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sites,T,causal", [(392, 16, True), (50, 16, False), (30, 6, True), (20, 1, True), (12, 24, True), (6, 128, True), (5, 40, False)])
def test_temporal_attention(dtype, sites, T, causal):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(14)
    H, D = 12, 768
    qkv = torch.randn(sites * T, 3 * D, generator=g).to(DEV, dtype)
    out = ops.temporal_attention(qkv, sites, H, T, causal, 0.125)
    x = qkv.float().view(sites, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = torch.tril(torch.ones(T, T, dtype=torch.bool, device=DEV)) if causal else None
    ref = _attn_ref(x[0], x[1], x[2], 0.125, mask).transpose(1, 2).reshape(sites * T, D)
    _close(out, ref, dtype, "temporal attention", scale=2.0)


@pytest.mark.parametrize("Tnew,steps", [(1, 20), (8, 2), (3, 7)])
def test_temporal_attention_kv_cache(Tnew, steps):
    """Streaming: append T_new frames per step; each step must equal the matching rows of the
    one-shot causal attention over all frames."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(15)
    sites, H, D, cap = 40, 12, 768, 32
    total = Tnew * steps
    assert total <= cap
    qkv_all = torch.randn(sites, total, 3 * D, generator=g).to(DEV, torch.bfloat16)
    x = qkv_all.float().view(sites, total, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = torch.tril(torch.ones(total, total, dtype=torch.bool, device=DEV))
    ref_all = _attn_ref(x[0], x[1], x[2], 0.125, mask).transpose(1, 2).reshape(sites, total, D)
    kc = torch.zeros(sites, H, cap, 64, device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    for s in range(steps):
        pos0 = s * Tnew
        step = qkv_all[:, pos0:pos0 + Tnew].reshape(sites * Tnew, 3 * D).contiguous()
        ops.kv_append(step, kc, vc, sites, H, Tnew, pos0)
        out = ops.temporal_attention(step, sites, H, Tnew, True, 0.125, kcache=kc, vcache=vc, Tk=pos0 + Tnew, q_off=pos0)
        ref = ref_all[:, pos0:pos0 + Tnew].reshape(sites * Tnew, D)
        _close(out, ref, torch.bfloat16, f"kv-cache step {s}", scale=2.0)


@pytest.fixture(params=["direct", "tma_ring"])
def decode_kernel(request):
    """Both streaming decode kernels: the register-direct default and the TMA-ring / mma.sync one."""
    from streamformer_b200 import _native as N
    N.set_option("decode_tma", 1 if request.param == "tma_ring" else 0)
    yield request.param
    N.set_option("decode_tma", -1)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sites,cap,steps", [(784, 64, 64), (61, 96, 96), (5, 8, 8), (3, 33, 33)])
def test_temporal_decode_kernel(dtype, sites, cap, steps, decode_kernel):
    """Streaming decode kernels (history straight into registers / bulk-copy staged, fused append): every step equals
    the matching row of the one-shot causal attention; the cache it leaves behind equals what kv_append writes."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(17)
    H, D = 12, 768
    qkv_all = torch.randn(sites, steps, 3 * D, generator=g).to(DEV, dtype)
    x = qkv_all.float().view(sites, steps, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = torch.tril(torch.ones(steps, steps, dtype=torch.bool, device=DEV))
    ref_all = _attn_ref(x[0], x[1], x[2], 0.125, mask).transpose(1, 2).reshape(sites, steps, D)
    kc = torch.zeros(sites, H, cap, 64, device=DEV, dtype=dtype)
    vc = torch.zeros_like(kc)
    kc2, vc2 = torch.zeros_like(kc), torch.zeros_like(kc)
    for s in range(steps):
        step = qkv_all[:, s].contiguous()
        out = ops.temporal_decode(step, kc, vc, sites, H, s, 0.125)
        _close(out, ref_all[:, s], dtype, f"decode step {s}", scale=2.0)
        if s in (0, 1, steps // 2, steps - 1):
            ops.kv_append(step, kc2, vc2, sites, H, 1, s)
            assert torch.equal(kc[:, :, s], kc2[:, :, s]) and torch.equal(vc[:, :, s], vc2[:, :, s])
    with pytest.raises(Exception, match="capacity"):
        ops.temporal_decode(qkv_all[:, 0].contiguous(), kc, vc, sites, H, cap, 0.125)


@pytest.fixture(params=["row", "general"])
def spatial_kernel(request):
    """Both tcgen05 spatial kernels at S = 196: rows in registers (default) and the general two-pass one."""
    from streamformer_b200 import _native as N
    N.set_option("spatial_row", 1 if request.param == "row" else 0)
    yield request.param
    N.set_option("spatial_row", -1)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("frames,S", [(16, 196), (3, 49), (2, 392), (1, 64), (2, 1), (5, 193), (40, 200)])
def test_spatial_attention(dtype, frames, S, spatial_kernel):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(16)
    H, D = 12, 768
    qkv = torch.randn(frames * S, 3 * D, generator=g).to(DEV, dtype)
    out, probs = ops.spatial_attention(qkv, frames, H, S, 0.125, want_probs=True)
    x = qkv.float().view(frames, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(x[0], x[1], x[2], 0.125).transpose(1, 2).reshape(frames * S, D)
    _close(out, ref, dtype, "spatial attention", scale=2.0)
    pref = ((x[0] @ x[1].transpose(-2, -1)) * 0.125).softmax(-1)
    assert torch.allclose(probs, pref, atol=1e-4, rtol=1e-3), (probs - pref).abs().max()


@pytest.mark.parametrize("B,T,S", [(2, 16, 196), (1, 6, 196), (3, 5, 49), (4, 1, 196), (1, 24, 64)])
def test_spatial_attention_in_place_layout(B, T, S, spatial_kernel):
    """Frames read in place from the residual stream's (b,n,t) row order (row stride T), outputs
    written back in the same order: must equal the contiguous-frame result on permuted rows."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(18)
    H, D = 12, 768
    qkv_bnt = torch.randn(B * S * T, 3 * D, generator=g).to(DEV, torch.bfloat16)
    out, probs = ops.spatial_attention(qkv_bnt, B * T, H, S, 0.125, want_probs=True, T_inner=T)
    qkv_btn = qkv_bnt.view(B, S, T, 3 * D).permute(0, 2, 1, 3).reshape(B * T * S, 3 * D).contiguous()
    ref, pref = ops.spatial_attention(qkv_btn, B * T, H, S, 0.125, want_probs=True)
    ref_bnt = ref.view(B, T, S, D).permute(0, 2, 1, 3).reshape(B * S * T, D)
    assert torch.equal(out, ref_bnt)
    assert torch.equal(probs, pref)


@pytest.mark.parametrize("frames,S", [(16, 196), (5, 49), (1, 400)])
def test_pool_attention(frames, S):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(17)
    H, D = 12, 768
    kv = torch.randn(frames * S, 2 * D, generator=g).to(DEV, torch.bfloat16)
    q = (torch.randn(D, generator=g) * 0.125).to(DEV)
    out = ops.pool_attention(kv, q, frames, H, S)
    k = kv[:, :D].float().view(frames, S, H, 64).permute(0, 2, 1, 3)
    v = kv[:, D:].float().view(frames, S, H, 64).permute(0, 2, 1, 3)
    s = torch.einsum("hd,fhsd->fhs", q.view(H, 64), k).softmax(-1)
    ref = torch.einsum("fhs,fhsd->fhd", s, v).reshape(frames, D)
    _close(out, ref, torch.bfloat16, "pool attention", scale=2.0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("frames,S", [(16, 196), (5, 49), (3, 1), (2, 400)])
def test_pool_probe_matches_projected_attention(dtype, frames, S):
    """Collapsed probe pooling == softmax(q . (W_k x + b_k)) (W_v x + b_v) per head, computed the long way in fp32."""
    ops = _ops()
    heads, D = 12, 768
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(frames * S, D, generator=g).to(DEV, dtype)
    wk = (torch.randn(D, D, generator=g) * 0.05).to(DEV)
    wv = (torch.randn(D, D, generator=g) * 0.05).to(DEV, dtype)
    bk = torch.randn(D, generator=g).to(DEV)
    bv = torch.randn(D, generator=g).to(DEV)
    q = (torch.randn(D, generator=g) * 0.3).to(DEV)                  # already scaled probe query
    u = torch.einsum("hdk,hd->hk", wk.view(heads, 64, D), q.view(heads, 64)).contiguous()
    out = ops.pool_probe(x, u, wv, bv, frames, heads, S).float()
    xf = x.float().view(frames, S, D)
    k = (xf @ wk.t() + bk).view(frames, S, heads, 64)
    v = (xf @ wv.float().t() + bv).view(frames, S, heads, 64)
    p = torch.softmax(torch.einsum("fshd,hd->fhs", k, q.view(heads, 64)), dim=-1)
    ref = torch.einsum("fhs,fshd->fhd", p, v).reshape(frames, D)
    _close(out, ref, dtype, "pool_probe")
