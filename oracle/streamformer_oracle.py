"""CPU oracle for the StreamFormer encoder hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy (fp32) restatement of ``TimesformerMultiTaskingModelSigLIP.forward`` and of the
KV-cache twin, written independently of the CUDA implementation so that the two can be compared on
the same seeded inputs.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module; the product path (``streamformer_b200``) never does and fails loudly
when its CUDA extension is missing.

Parity pinning: the reference ships NO tests or golden vectors for this path, so the
oracle is pinned against *outputs of the reference itself*, produced in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference read-only) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every fixture on CPU.

Every function cites the reference lines it restates.  Paths are relative to the reference root:
  R  = models/modeling_timesformer_siglip.py
  KV = downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

try:  # exact erf for GELU; scipy is in the image, the fallback is the same function, slower
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float32])

F32 = np.float32


@dataclass
class OracleConfig:
    """Field-for-field subset of StreamformerConfig (models/configuration_streamformer.py:92-137)."""
    image_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    num_frames: int = 16
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    layer_norm_eps: float = 1e-6
    qkv_bias: bool = True
    attention_type: str = "divided_space_time"
    enable_causal_temporal: bool = True
    add_lora_spatial: bool = False
    lora_rank: int = 32

    @property
    def num_patches(self) -> int:
        g = self.image_size // self.patch_size
        return g * g

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


# --------------------------------------------------------------------------------------- weights
def make_weights(cfg: OracleConfig, seed: int = 0, style: str = "reference") -> Dict[str, np.ndarray]:
    """Deterministic fp32 weights under the reference's state-dict names.

    The reference initialises the temporal gates and time embeddings to ZERO (R:896, R:377), which
    would switch the whole temporal branch off, so the oracle weights draw gates from U(-1,1), time
    embeddings from N(0,0.02) and LoRA-B from N(0,0.02).  ``style="stress"`` widens the Q/K/V
    projections so the softmaxes are peaked rather than near-uniform.
    """
    rng = np.random.RandomState(seed)
    D, I, L, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.num_attention_heads
    P, C, N, F = cfg.patch_size, cfg.num_channels, cfg.num_patches, cfg.num_frames
    qkv_std = 0.02 if style == "reference" else 0.06
    w: Dict[str, np.ndarray] = {}

    def normal(shape, std):
        return (rng.standard_normal(shape) * std).astype(F32)

    def linear(prefix, out_f, in_f, std=0.02, bias=True):
        w[prefix + ".weight"] = np.clip(normal((out_f, in_f), std), -2 * std, 2 * std)  # trunc_normal_ (R:1079)
        if bias:
            w[prefix + ".bias"] = normal((out_f,), 0.02)

    def lnorm(prefix):
        w[prefix + ".weight"] = (1.0 + 0.1 * rng.standard_normal(D)).astype(F32)
        w[prefix + ".bias"] = normal((D,), 0.05)

    w["embeddings.position_embeddings"] = normal((1, N, D), 0.02)
    w["embeddings.time_embeddings"] = normal((1, F, D), 0.02)
    w["embeddings.patch_embeddings.projection.weight"] = normal((D, C, P, P), 0.02)
    w["embeddings.patch_embeddings.projection.bias"] = normal((D,), 0.02)
    for l in range(L):
        p = f"encoder.layer.{l}."
        w[p + "temporal_attention_gating"] = np.asarray(rng.uniform(-1.0, 1.0), dtype=F32)
        linear(p + "attention.attention.qkv", 3 * D, D, qkv_std, cfg.qkv_bias)
        linear(p + "attention.output.dense", D, D)
        linear(p + "intermediate.dense", I, D)
        linear(p + "output.dense", D, I)
        lnorm(p + "layernorm_before")
        lnorm(p + "layernorm_after")
        lnorm(p + "temporal_layernorm")
        w[p + "temporal_attention.attention.mask"] = np.tril(np.ones((F, F), dtype=F32))  # unused buffer (R:515-517)
        linear(p + "temporal_attention.attention.qkv", 3 * D, D, qkv_std, cfg.qkv_bias)
        linear(p + "temporal_attention.output.dense", D, D)
        linear(p + "temporal_dense", D, D)
        if cfg.add_lora_spatial:
            r = cfg.lora_rank
            w[p + "attention.attention.qkv_lora_a.weight"] = normal((r, D), 0.02)
            w[p + "attention.attention.qkv_lora_b.weight"] = normal((3 * D, r), 0.02)
            w[p + "attention.output.dense_lora_a.weight"] = normal((r, D), 0.02)
            w[p + "attention.output.dense_lora_b.weight"] = normal((D, r), 0.02)
    lnorm("post_layernorm")
    w["head.probe"] = normal((1, 1, D), 1.0)  # torch.randn (R:1134)
    w["head.attention.in_proj_weight"] = normal((3 * D, D), 0.03)
    w["head.attention.in_proj_bias"] = normal((3 * D,), 0.02)
    linear("head.attention.out_proj", D, D)
    lnorm("head.layernorm")
    linear("head.mlp.fc1", I, D)
    linear("head.mlp.fc2", D, I)
    return w


def make_pixels(B: int, T: int, cfg: OracleConfig, seed: int = 0, H: Optional[int] = None,
                W: Optional[int] = None) -> np.ndarray:
    rng = np.random.RandomState(10_000 + seed)
    H = H or cfg.image_size
    W = W or cfg.image_size
    return rng.standard_normal((B, T, cfg.num_channels, H, W)).astype(F32)


# --------------------------------------------------------------------------------------- primitives
def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    """nn.Linear: x @ W^T + b."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y.astype(F32, copy=False)


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    """nn.LayerNorm over the last dim, biased variance (R:860-880)."""
    x64 = x.astype(np.float64)
    mu = x64.mean(-1, keepdims=True)
    var = ((x64 - mu) ** 2).mean(-1, keepdims=True)
    return (((x64 - mu) / np.sqrt(var + eps)) * g + b).astype(F32)


def gelu(x: np.ndarray, kind: str = "gelu") -> np.ndarray:
    """ACT2FN["gelu"] = exact erf GELU (R:814-817); "gelu_pytorch_tanh" = tanh approximation."""
    if kind == "gelu":
        return (0.5 * x * (1.0 + _erf(x * np.float32(0.7071067811865476)))).astype(F32)
    if kind in ("gelu_pytorch_tanh", "gelu_new"):
        return (0.5 * x * (1.0 + np.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))).astype(F32)
    raise ValueError(kind)


def softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(-1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(x - m)
    return (e / e.sum(-1, keepdims=True)).astype(F32)


def split_heads(qkv: np.ndarray, heads: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """reshape(B, L, 3, heads, hd).permute(2, 0, 3, 1, 4) (R:577-588, 690-701)."""
    Bx, Lx, D3 = qkv.shape
    hd = D3 // 3 // heads
    t = qkv.reshape(Bx, Lx, 3, heads, hd).transpose(2, 0, 3, 1, 4)
    return t[0], t[1], t[2]


def merge_heads(ctx: np.ndarray) -> np.ndarray:
    """(B, heads, L, hd).transpose(1, 2).reshape(B, L, D) (R:605-609)."""
    Bx, Hh, Lx, hd = ctx.shape
    return ctx.transpose(0, 2, 1, 3).reshape(Bx, Lx, Hh * hd)


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """F.interpolate(mode="nearest") source indices: min(floor(dst * float(in)/out), in-1) in fp32."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


# --------------------------------------------------------------------------------------- modules
def patch_embeddings(w: Dict[str, np.ndarray], cfg: OracleConfig, pixels: np.ndarray) -> np.ndarray:
    """TimesformerPatchEmbeddings.forward (R:336-350): Conv2d(k=P, s=P) as a patch GEMM.
    pixels [B,T,C,H,W] -> [B*T, N, D] with N ordered (row, col) and K ordered (c, kh, kw)."""
    B, T, C, H, W = pixels.shape
    P = cfg.patch_size
    gh, gw = H // P, W // P
    x = pixels.reshape(B * T, C, gh, P, gw, P).transpose(0, 2, 4, 1, 3, 5).reshape(B * T, gh * gw, C * P * P)
    wt = w["embeddings.patch_embeddings.projection.weight"].reshape(cfg.hidden_size, -1)
    return linear(x, wt, w["embeddings.patch_embeddings.projection.bias"])


def _cubic_aa_filter(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    """Keys cubic convolution kernel with a = -0.5, the filter of ATen's anti-aliased bicubic resize."""
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    far = (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def _aa_resize_matrix(in_size: int, out_size: int) -> np.ndarray:
    """[out, in] weights of one separable pass of F.interpolate(mode="bicubic", antialias=True,
    align_corners=False): support 2*max(scale,1) taps around centre scale*(i+0.5), weights normalised per
    output sample (torch ATen UpSampleKernel `_compute_indices_weights_aa`; torch is the third-party
    dependency the reference calls at R:402-407, pinned 2.5.1 in requirements.txt:26)."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    inv = 1.0 / scale if scale >= 1.0 else 1.0
    m = np.zeros((out_size, in_size), dtype=np.float64)
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        j = np.arange(lo, hi)
        wgt = _cubic_aa_filter((j - center + 0.5) * inv)
        m[i, lo:hi] = wgt / wgt.sum()
    return m


def position_table(w: Dict[str, np.ndarray], cfg: OracleConfig, npatch: int, H: int, W: int) -> np.ndarray:
    """interpolate_pos_encoding (R:380-411).  Identity when npatch == N and W == H; otherwise the
    [M, M] table is resampled (bicubic, antialias) to size (w0, h0) = (W // P, H // P) — in THAT order,
    as the reference passes it (R:389-401) — and flattened row-major, so for a non-square input the
    table is laid out [w0, h0] while the patch tokens are laid out [H // P, W // P]."""
    pos = w["embeddings.position_embeddings"]
    if npatch == pos.shape[1] and W == H:
        return pos
    Np, D = pos.shape[1], pos.shape[2]
    M = int(math.sqrt(Np))
    assert Np == M * M
    w0, h0 = W // cfg.patch_size, H // cfg.patch_size
    grid = pos.reshape(M, M, D).astype(np.float64)
    rows = _aa_resize_matrix(M, w0)            # first spatial axis  -> w0
    cols = _aa_resize_matrix(M, h0)            # second spatial axis -> h0
    out = np.einsum("im,mnd->ind", rows, grid)
    out = np.einsum("jn,ind->ijd", cols, out)
    assert out.shape[0] * out.shape[1] == npatch
    return out.reshape(1, npatch, D).astype(F32)


def embeddings(w: Dict[str, np.ndarray], cfg: OracleConfig, pixels: np.ndarray, past_frames: int = 0,
               time_total: Optional[int] = None) -> np.ndarray:
    """TimesformerEmbeddingsSigLIP.forward (R:413-457) and its KV twin (KV:307-375).
    Returns [B, N*T, D] with token index n*T + t."""
    B, T, _, H, W = pixels.shape
    x = patch_embeddings(w, cfg, pixels)                                   # (B*T, N, D)
    N, D = x.shape[1], x.shape[2]
    x = x + position_table(w, cfg, N, H, W)                                # R:418-420
    x = x.reshape(B, T, N, D).transpose(0, 2, 1, 3).reshape(B * N, T, D)   # R:427-433
    te = w["embeddings.time_embeddings"]                                   # (1, F, D)
    F = te.shape[1]
    end = past_frames + T
    total = end if time_total is None else max(time_total, end)
    if total <= F:
        t_sel = te[:, past_frames:end, :]                                  # R:436-439 / KV:354-356
    else:
        idx = nearest_index(total, F)                                      # R:441-447 / KV:340-352
        t_sel = te[:, idx[past_frames:end], :]
    x = x + t_sel
    return x.reshape(B, N * T, D).astype(F32)                              # R:452-454


@dataclass
class TemporalCache:
    """What transformers.DynamicCache holds for the KV twin (KV:517-518): per layer K and V
    [B*N, heads, seen, hd], concatenated on the time axis."""
    keys: List[Optional[np.ndarray]] = field(default_factory=list)
    values: List[Optional[np.ndarray]] = field(default_factory=list)

    def seq_len(self) -> int:
        return 0 if not self.keys or self.keys[0] is None else self.keys[0].shape[2]

    def update(self, k: np.ndarray, v: np.ndarray, layer: int) -> Tuple[np.ndarray, np.ndarray]:
        while len(self.keys) <= layer:
            self.keys.append(None)
            self.values.append(None)
        if self.keys[layer] is None:
            self.keys[layer], self.values[layer] = k, v
        else:
            self.keys[layer] = np.concatenate([self.keys[layer], k], axis=2)
            self.values[layer] = np.concatenate([self.values[layer], v], axis=2)
        return self.keys[layer], self.values[layer]


def temporal_attention(w, cfg: OracleConfig, p: str, x: np.ndarray, cache: Optional[TemporalCache], layer: int,
                       start_pos: int) -> np.ndarray:
    """TimesformerCausalSelfAttention.forward (R:575-615; cached variant KV:491-560) followed by
    TimesformerSelfOutput (R:759-763).  x: [B*N, T, D] (already layer-normed).
    With enable_causal_temporal=False the reference uses the unmasked TimesformerSelfAttention (R:894)."""
    heads = cfg.num_attention_heads
    scale = F32(cfg.head_dim ** -0.5)
    qkv = linear(x, w[p + "temporal_attention.attention.qkv.weight"], w.get(p + "temporal_attention.attention.qkv.bias"))
    q, k, v = split_heads(qkv, heads)
    if cache is not None:
        k, v = cache.update(k, v, layer)                                   # KV:517-518
    s = (q @ k.transpose(0, 1, 3, 2)) * scale                              # scale AFTER q@k^T (R:590)
    if cfg.enable_causal_temporal:
        Tq, Tk = s.shape[-2], s.shape[-1]
        i = np.arange(Tq)[:, None]
        j = np.arange(Tk)[None, :]
        mask = j <= (start_pos + i)                                        # R:593-601 / KV:533-537
        s = np.where(mask, s, -np.inf)
    ctx = merge_heads(softmax(s) @ v)                                      # R:602-609
    return linear(ctx, w[p + "temporal_attention.output.dense.weight"], w[p + "temporal_attention.output.dense.bias"])


def spatial_attention(w, cfg: OracleConfig, p: str, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """TimesformerSelfAttention.forward (R:688-717) or lora_forward (R:649-683), then
    TimesformerSelfOutput.forward / lora_forward (R:748-763).  x: [B*T, N, D] (layer-normed).
    Returns (output, attention_probs)."""
    heads = cfg.num_attention_heads
    scale = F32(cfg.head_dim ** -0.5)
    qkv = linear(x, w[p + "attention.attention.qkv.weight"], w.get(p + "attention.attention.qkv.bias"))
    if (p + "attention.attention.qkv_lora_a.weight") in w:
        qkv = qkv + (x @ w[p + "attention.attention.qkv_lora_a.weight"].T) @ w[p + "attention.attention.qkv_lora_b.weight"].T
    q, k, v = split_heads(qkv, heads)
    probs = softmax((q @ k.transpose(0, 1, 3, 2)) * scale)
    ctx = merge_heads(probs @ v)
    out = linear(ctx, w[p + "attention.output.dense.weight"], w[p + "attention.output.dense.bias"])
    if (p + "attention.output.dense_lora_a.weight") in w:
        out = out + (ctx @ w[p + "attention.output.dense_lora_a.weight"].T) @ w[p + "attention.output.dense_lora_b.weight"].T
    return out.astype(F32), probs


def layer_forward(w, cfg: OracleConfig, layer: int, x: np.ndarray, T: int, cache: Optional[TemporalCache] = None,
                  start_pos: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """TimesformerLayerSigLIP.forward, divided_space_time branch (R:934-1004).  x: [B, N*T, D]."""
    p = f"encoder.layer.{layer}."
    B, NT, D = x.shape
    N = NT // T
    eps = cfg.layer_norm_eps
    # temporal (R:937-958)
    xt = x.reshape(B * N, T, D)
    att = temporal_attention(w, cfg, p, layer_norm(xt, w[p + "temporal_layernorm.weight"], w[p + "temporal_layernorm.bias"], eps),
                             cache, layer, start_pos)
    res_t = linear(att.reshape(B, NT, D), w[p + "temporal_dense.weight"], w[p + "temporal_dense.bias"])
    h1 = x + np.tanh(w[p + "temporal_attention_gating"]).astype(F32) * res_t
    # spatial (R:960-991)
    xs = h1.reshape(B, N, T, D).transpose(0, 2, 1, 3).reshape(B * T, N, D)
    att_s, probs = spatial_attention(w, cfg, p, layer_norm(xs, w[p + "layernorm_before.weight"], w[p + "layernorm_before.bias"], eps))
    res_s = att_s.reshape(B, T, N, D).transpose(0, 2, 1, 3).reshape(B, NT, D)
    # MLP (R:993-1000)
    h2 = h1 + res_s
    y = layer_norm(h2, w[p + "layernorm_after.weight"], w[p + "layernorm_after.bias"], eps)
    y = gelu(linear(y, w[p + "intermediate.dense.weight"], w[p + "intermediate.dense.bias"]), cfg.hidden_act)
    y = linear(y, w[p + "output.dense.weight"], w[p + "output.dense.bias"])
    return (h2 + y).astype(F32), probs


def pooling_head(w, cfg: OracleConfig, x: np.ndarray) -> np.ndarray:
    """TimesformerSiglipMultiheadAttentionPoolingHead.forward (R:1141-1154): nn.MultiheadAttention with the
    learned probe as the only query (F.multi_head_attention_forward: packed in_proj, q scaled by hd**-0.5),
    then r + SiglipMLP(LayerNorm(r)) (R:1113-1125).  x: [B*T, N, D] -> [B*T, D]."""
    D, heads, hd = cfg.hidden_size, cfg.num_attention_heads, cfg.head_dim
    ipw, ipb = w["head.attention.in_proj_weight"], w["head.attention.in_proj_bias"]
    probe = w["head.probe"].reshape(1, D)
    q = linear(probe, ipw[:D], ipb[:D]) * F32(hd ** -0.5)                        # [1, D]
    k = linear(x, ipw[D:2 * D], ipb[D:2 * D])                                    # [BT, N, D]
    v = linear(x, ipw[2 * D:], ipb[2 * D:])
    BT, N, _ = x.shape
    qh = q.reshape(heads, hd)
    kh = k.reshape(BT, N, heads, hd)
    vh = v.reshape(BT, N, heads, hd)
    s = np.einsum("hd,bnhd->bhn", qh, kh)
    pr = softmax(s)
    ctx = np.einsum("bhn,bnhd->bhd", pr, vh).reshape(BT, D)
    r = linear(ctx, w["head.attention.out_proj.weight"], w["head.attention.out_proj.bias"])
    y = layer_norm(r, w["head.layernorm.weight"], w["head.layernorm.bias"], cfg.layer_norm_eps)
    y = gelu(linear(y, w["head.mlp.fc1.weight"], w["head.mlp.fc1.bias"]), cfg.hidden_act)
    y = linear(y, w["head.mlp.fc2.weight"], w["head.mlp.fc2.bias"])
    return (r + y).astype(F32)


def forward(w, cfg: OracleConfig, pixels: np.ndarray, output_hidden_states: bool = False,
            output_attentions: bool = False, cache: Optional[TemporalCache] = None,
            time_total: Optional[int] = None) -> Dict[str, object]:
    """TimesformerMultiTaskingModelSigLIP.forward (R:1299-1354); with ``cache`` the KV twin's forward
    (KV:1316-1392: cache_position = arange(seen, seen+T), time embeddings offset by frames seen).
    Returns last_hidden_state [B,T,N,D], pooler_output [B,T,D], hidden_states, attentions."""
    B, T = pixels.shape[:2]
    seen = cache.seq_len() if cache is not None else 0
    x = embeddings(w, cfg, pixels, past_frames=seen, time_total=time_total)   # R:1319
    D = x.shape[-1]
    hs = [x] if output_hidden_states else None
    atts = [] if output_attentions else None
    for l in range(cfg.num_hidden_layers):                                    # R:1030-1048
        x, probs = layer_forward(w, cfg, l, x, T, cache, seen)
        if hs is not None:
            hs.append(x)
        if atts is not None:
            atts.append(probs)
    x = layer_norm(x, w["post_layernorm.weight"], w["post_layernorm.bias"], cfg.layer_norm_eps)  # R:1330
    N = x.shape[1] // T
    seq = x.reshape(B, N, T, D).transpose(0, 2, 1, 3)                          # (B, T, N, D)  R:1332-1346
    pooled = pooling_head(w, cfg, seq.reshape(B * T, N, D)).reshape(B, T, D)   # R:1338-1340
    return {
        "last_hidden_state": np.ascontiguousarray(seq, dtype=F32),
        "pooler_output": pooled,
        "hidden_states": hs,
        "attentions": atts,
    }


# --------------------------------------------------------------------------------------- accounting
def attention_block_flops_per_clip(cfg: OracleConfig, T: int) -> float:
    """Algorithmic FLOPs of ONE layer's space-time attention block for one clip (SURVEY.md 8d:
    temporal QKV + attention + out-proj + temporal_dense, spatial QKV + attention + out-proj; 35.34 G
    at T=16)."""
    D, N = cfg.hidden_size, cfg.num_patches
    M = T * N
    return float(2 * M * D * 3 * D * 2 + 2 * M * D * D * 3 + 2 * 2 * M * T * D + 2 * 2 * M * N * D)


def flops_per_clip(cfg: OracleConfig, T: int) -> float:
    """Algorithmic FLOPs (2*M*N*K, full non-causal attention count, no LoRA, un-folded graph) of one
    clip of T frames — the figure BASELINE.md §4 and bench.py's roofline use."""
    D, I, N, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_patches, cfg.num_hidden_layers
    M = T * N
    Kp = cfg.num_channels * cfg.patch_size ** 2
    per_layer = 2 * M * D * 3 * D * 2          # temporal + spatial QKV
    per_layer += 2 * M * D * D * 3             # temporal out, temporal_dense, spatial out
    per_layer += 2 * 2 * M * T * D             # temporal QK^T + PV
    per_layer += 2 * 2 * M * N * D             # spatial QK^T + PV
    per_layer += 2 * 2 * M * D * I             # MLP
    head = 2 * M * D * 2 * D + 2 * 2 * T * N * D + 2 * T * D * D + 2 * 2 * T * D * I
    return float(2 * M * Kp * D + L * per_layer + head)


# --------------------------------------------------------------------------------------- task heads
def log_sigmoid(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float64)
    return np.minimum(x, 0.0) - np.log1p(np.exp(-np.abs(x)))


def classification_head(pooler_output: np.ndarray, label_embeddings: np.ndarray, logit_scale: float, logit_bias: float,
                        labels: np.ndarray) -> Tuple[float, np.ndarray]:
    """TimesformerVideoClassificationHead.forward (R:1704-1726): last-frame pooled feature, L2-normalised; the label
    embeddings are used as stored (already normalised / averaged, R:1665-1672).
    Returns (loss, logits_per_image [B, L])."""
    img = pooler_output[:, -1, :].astype(np.float64)
    img = img / np.linalg.norm(img, axis=-1, keepdims=True)                         # R:1711
    logits_per_text = label_embeddings.astype(np.float64) @ img.T * math.exp(logit_scale) + logit_bias   # R:1714-1718
    logits = logits_per_text.T
    target = -np.ones_like(logits)                                                  # R:1721-1724
    target[np.arange(labels.shape[0]), labels] = 1.0
    loss = -log_sigmoid(target * logits).sum() / labels.shape[0]                    # R:1725
    return float(loss), logits.astype(F32)


def siglip_loss(image_features: np.ndarray, text_features: np.ndarray, logit_scale_exp: float, logit_bias: float,
                negative_only: bool = False) -> float:
    """SigLipLoss._loss (R:220-243): features arrive normalised, logit_scale already exponentiated (R:2341-2343)."""
    logits = logit_scale_exp * image_features.astype(np.float64) @ text_features.astype(np.float64).T + logit_bias
    n = image_features.shape[0]
    labels = -np.ones((n, text_features.shape[0]))
    if not negative_only:
        labels = labels + 2.0 * np.eye(n, text_features.shape[0])
    return float(-log_sigmoid(labels * logits).sum() / n)


def siglip_loss_world(image_features: List[np.ndarray], text_features: List[np.ndarray], logit_scale_exp: float,
                      logit_bias: float) -> List[float]:
    """SigLipLoss.forward with world_size = len(image_features) (R:245-297): rank r's loss is its own block plus a
    negatives-only term against every other rank's text features (what the ring exchange accumulates)."""
    out = []
    for r, img in enumerate(image_features):
        loss = siglip_loss(img, text_features[r], logit_scale_exp, logit_bias)
        for q, txt in enumerate(text_features):
            if q != r:
                loss += siglip_loss(img, txt, logit_scale_exp, logit_bias, negative_only=True)
        out.append(loss)
    return out
