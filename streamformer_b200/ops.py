"""torch.Tensor wrappers over the single-kernel C-ABI entry points (sf_op_*).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation happens in
the hand-written sm_100a kernels.  Every wrapper requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as N

_DT = {torch.bfloat16: N.SF_BF16, torch.float16: N.SF_F16, torch.float32: N.SF_F32}


def sf_dtype(t: torch.dtype) -> int:
    if t not in _DT:
        raise TypeError(f"unsupported dtype {t}; the encoder runs in bfloat16 or float16")
    return _DT[t]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise N.NativeError("streamformer_b200 kernels need CUDA tensors (no CPU fallback)")


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = N.SF_ACT_NONE,
         residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
         row_map: int = N.SF_ROW_IDENTITY, T: int = 1, S: int = 1, pos: Optional[torch.Tensor] = None,
         time_emb: Optional[torch.Tensor] = None, time_total: int = 0, time_off: int = 0,
         out: Optional[torch.Tensor] = None, ln_stats: Optional[torch.Tensor] = None,
         ln_colsum: Optional[torch.Tensor] = None, ln_eps: float = 0.0,
         stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = epilogue(a[m] @ w.T); a [M,K], w [N,K] (nn.Linear layout), fp32 bias/pos/time/gate.
    ln_stats [parts, M, 2] + ln_colsum [N]: LayerNorm folded into the GEMM (w pre-scaled by gamma).
    stats_out [gemm_stats_parts(M, N), M, 2]: partial row statistics of the output."""
    _req(a, w, bias, residual, gate, pos, time_emb, out, ln_stats, ln_colsum, stats_out)
    M, K = a.shape
    Nn = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty(M, Nn, dtype=a.dtype, device=a.device)
    e = N.SfGemmEpilogue()
    e.bias = _p(bias); e.act = act
    e.residual = _p(residual); e.ldr = residual.stride(0) if residual is not None else 0
    e.gate = _p(gate); e.row_map = row_map; e.T = T; e.S = S
    e.pos = _p(pos); e.time_emb = _p(time_emb)
    e.time_len = time_emb.shape[0] if time_emb is not None else 0
    e.time_total = time_total; e.time_off = time_off
    e.ln_stats = _p(ln_stats); e.ln_parts = ln_stats.shape[0] if ln_stats is not None else 0
    e.ln_colsum = _p(ln_colsum); e.ln_eps = ln_eps; e.stats_out = _p(stats_out)
    for t in (bias, gate, pos, time_emb, ln_stats, ln_colsum, stats_out):
        assert t is None or t.dtype == torch.float32
    N.check(N.load().sf_op_gemm(_stream(), sf_dtype(a.dtype), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0),
                                out.data_ptr(), out.stride(0), M, Nn, K, e), "sf_op_gemm")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, row_map: int = 0, T: int = 1,
              S: int = 1) -> torch.Tensor:
    _req(x, gamma, beta)
    M, D = x.shape
    y = torch.empty(M, D, dtype=x.dtype, device=x.device)
    N.check(N.load().sf_op_layernorm(_stream(), sf_dtype(x.dtype), x.data_ptr(), x.stride(0), gamma.data_ptr(),
                                     beta.data_ptr(), eps, y.data_ptr(), y.stride(0), M, D, row_map, T, S),
            "sf_op_layernorm")
    return y


def im2col(pixels: torch.Tensor, patch: int, act_dtype: torch.dtype) -> torch.Tensor:
    _req(pixels)
    BT, Cc, H, W = pixels.shape
    S = (H // patch) * (W // patch)
    out = torch.empty(BT * S, Cc * patch * patch, dtype=act_dtype, device=pixels.device)
    N.check(N.load().sf_op_im2col(_stream(), sf_dtype(pixels.dtype), pixels.data_ptr(), sf_dtype(act_dtype),
                                  out.data_ptr(), BT, Cc, H, W, patch), "sf_op_im2col")
    return out


def temporal_attention(qkv: torch.Tensor, sites: int, heads: int, Tq: int, causal: bool, scale: float,
                       kcache: Optional[torch.Tensor] = None, vcache: Optional[torch.Tensor] = None, Tk: int = 0,
                       q_off: int = 0) -> torch.Tensor:
    _req(qkv, kcache, vcache)
    D = heads * 64
    out = torch.empty(sites * Tq, D, dtype=qkv.dtype, device=qkv.device)
    Tcap = kcache.shape[2] if kcache is not None else 0
    N.check(N.load().sf_op_temporal_attention(_stream(), sf_dtype(qkv.dtype), qkv.data_ptr(), qkv.stride(0),
                                              _p(kcache), _p(vcache), Tcap, out.data_ptr(), out.stride(0), sites,
                                              heads, Tq, Tk if kcache is not None else Tq, q_off, int(causal), scale),
            "sf_op_temporal_attention")
    return out


def temporal_decode(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, sites: int, heads: int, seen: int,
                    scale: float) -> torch.Tensor:
    """One new frame per site: appends its K/V row at cache index ``seen`` and attends to rows 0..seen."""
    _req(qkv, kcache, vcache)
    out = torch.empty(sites, heads * 64, dtype=qkv.dtype, device=qkv.device)
    N.check(N.load().sf_op_temporal_decode(_stream(), sf_dtype(qkv.dtype), qkv.data_ptr(), qkv.stride(0), kcache.data_ptr(),
                                           vcache.data_ptr(), kcache.shape[2], out.data_ptr(), out.stride(0), sites, heads,
                                           seen, scale), "sf_op_temporal_decode")
    return out


def kv_append(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, sites: int, heads: int, Tq: int,
              pos0: int) -> None:
    _req(qkv, kcache, vcache)
    N.check(N.load().sf_op_kv_append(_stream(), sf_dtype(qkv.dtype), qkv.data_ptr(), qkv.stride(0), kcache.data_ptr(),
                                     vcache.data_ptr(), kcache.shape[2], sites, heads, Tq, pos0), "sf_op_kv_append")


def gemm_stats_parts(M: int, Nn: int) -> int:
    return int(N.load().sf_op_gemm_stats_parts(M, Nn))


def rowstats(x: torch.Tensor) -> torch.Tensor:
    """[1, M, 2] fp32 (sum, sum of squares) per row."""
    _req(x)
    M, D = x.shape
    st = torch.empty(1, M, 2, dtype=torch.float32, device=x.device)
    N.check(N.load().sf_op_rowstats(_stream(), sf_dtype(x.dtype), x.data_ptr(), x.stride(0), M, D, st.data_ptr()),
            "sf_op_rowstats")
    return st


def spatial_attention(qkv: torch.Tensor, frames: int, heads: int, S: int, scale: float,
                      want_probs: bool = False, T_inner: int = 1):
    _req(qkv)
    D = heads * 64
    out = torch.empty(frames * S, D, dtype=qkv.dtype, device=qkv.device)
    probs = torch.empty(frames, heads, S, S, dtype=torch.float32, device=qkv.device) if want_probs else None
    N.check(N.load().sf_op_spatial_attention(_stream(), sf_dtype(qkv.dtype), qkv.data_ptr(), qkv.stride(0),
                                             out.data_ptr(), out.stride(0), frames, heads, S, T_inner, scale,
                                             _p(probs)),
            "sf_op_spatial_attention")
    return (out, probs) if want_probs else out


def pool_attention(kv: torch.Tensor, q: torch.Tensor, frames: int, heads: int, S: int) -> torch.Tensor:
    _req(kv, q)
    assert q.dtype == torch.float32
    out = torch.empty(frames, heads * 64, dtype=kv.dtype, device=kv.device)
    N.check(N.load().sf_op_pool_attention(_stream(), sf_dtype(kv.dtype), kv.data_ptr(), kv.stride(0), q.data_ptr(),
                                          out.data_ptr(), out.stride(0), frames, heads, S), "sf_op_pool_attention")
    return out


def pool_probe(tokens: torch.Tensor, u: torch.Tensor, wv: torch.Tensor, bv: torch.Tensor, frames: int, heads: int,
               S: int) -> torch.Tensor:
    """Pooling attention with a single probe, K/V projections collapsed: tokens [frames*S, D] -> [frames, D]."""
    _req(tokens, u, wv, bv)
    assert u.dtype == torch.float32 and bv.dtype == torch.float32 and wv.dtype == tokens.dtype
    D = heads * 64
    out = torch.empty(frames, D, dtype=tokens.dtype, device=tokens.device)
    N.check(N.load().sf_op_pool_probe(_stream(), sf_dtype(tokens.dtype), tokens.data_ptr(), tokens.stride(0), u.data_ptr(),
                                      wv.data_ptr(), bv.data_ptr(), out.data_ptr(), out.stride(0), frames, heads, S),
            "sf_op_pool_probe")
    return out


# ------------------------------------------------------------------------------------- backward-pass ops
def transpose(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[M, N] -> [N, Mpad] (Mpad = M rounded up to 8, padding zeroed): the wgrad operand layout."""
    _req(x, out)
    M, Nn = x.shape
    Mp = (M + 7) // 8 * 8
    if out is None:
        out = torch.empty(Nn, Mp, dtype=x.dtype, device=x.device)
    N.check(N.load().sf_op_transpose(_stream(), sf_dtype(x.dtype), x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), M, Nn),
            "sf_op_transpose")
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    _req(x)
    M, Nn = x.shape
    out = torch.empty(Nn, dtype=torch.float32, device=x.device)
    N.check(N.load().sf_op_colsum(_stream(), sf_dtype(x.dtype), x.data_ptr(), x.stride(0), M, Nn, out.data_ptr()), "sf_op_colsum")
    return out


def ln_backward(x: torch.Tensor, dn: torch.Tensor, eps: float, dres: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, dn, dres, out)
    M, D = x.shape
    if out is None:
        out = torch.empty(M, D, dtype=x.dtype, device=x.device)
    N.check(N.load().sf_op_ln_backward(_stream(), sf_dtype(x.dtype), x.data_ptr(), x.stride(0), dn.data_ptr(), dn.stride(0), eps,
                                       _p(dres), dres.stride(0) if dres is not None else 0, out.data_ptr(), out.stride(0), M, D),
            "sf_op_ln_backward")
    return out


def ln_affine_backward(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float, dgamma: torch.Tensor, dbeta: torch.Tensor,
                       row_map: int = 0, T: int = 1, S: int = 1) -> torch.Tensor:
    _req(x, dy, gamma, dgamma, dbeta)
    assert gamma.dtype == torch.float32 and dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32
    M, D = x.shape
    out = torch.empty(M, D, dtype=x.dtype, device=x.device)
    N.check(N.load().sf_op_ln_affine_backward(_stream(), sf_dtype(x.dtype), x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0),
                                              gamma.data_ptr(), eps, out.data_ptr(), out.stride(0), M, D, row_map, T, S,
                                              dgamma.data_ptr(), dbeta.data_ptr()), "sf_op_ln_affine_backward")
    return out


def gelu(a: torch.Tensor, act: int = N.SF_ACT_GELU) -> torch.Tensor:
    _req(a)
    assert a.is_contiguous()
    h = torch.empty_like(a)
    N.check(N.load().sf_op_gelu(_stream(), sf_dtype(a.dtype), a.data_ptr(), h.data_ptr(), a.numel(), act), "sf_op_gelu")
    return h


def gelu_backward_(a_h: torch.Tensor, dh_dpre: torch.Tensor, act: int = N.SF_ACT_GELU) -> None:
    """In place: a -> GELU(a), dh -> dh * GELU'(a)."""
    _req(a_h, dh_dpre)
    assert a_h.is_contiguous() and dh_dpre.is_contiguous() and a_h.numel() == dh_dpre.numel()
    N.check(N.load().sf_op_gelu_backward(_stream(), sf_dtype(a_h.dtype), a_h.data_ptr(), dh_dpre.data_ptr(), a_h.numel(), act),
            "sf_op_gelu_backward")


def gate_backward(dx: torch.Tensor, y: torch.Tensor, gate: torch.Tensor, dgate: torch.Tensor) -> torch.Tensor:
    _req(dx, y, gate, dgate)
    assert dx.is_contiguous() and y.is_contiguous() and gate.dtype == torch.float32 and dgate.dtype == torch.float32
    dy = torch.empty_like(dx)
    N.check(N.load().sf_op_gate_backward(_stream(), sf_dtype(dx.dtype), dx.data_ptr(), y.data_ptr(), gate.data_ptr(), dy.data_ptr(),
                                         dx.numel(), dgate.data_ptr()), "sf_op_gate_backward")
    return dy


def wfold_finish(G: torch.Tensor, out_dtype: torch.dtype, Wp: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
                 beta: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None, dgamma: Optional[torch.Tensor] = None,
                 dbeta: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(G, Wp, gamma, beta, db, dgamma, dbeta)
    O, I = G.shape
    dW = torch.empty(O, I, dtype=out_dtype, device=G.device)
    N.check(N.load().sf_op_wfold_finish(_stream(), sf_dtype(G.dtype), G.data_ptr(), G.stride(0), _p(Wp), Wp.stride(0) if Wp is not None else 0,
                                        _p(gamma), _p(beta), _p(db), dW.data_ptr(), sf_dtype(out_dtype), I, O, I, _p(dgamma), _p(dbeta)),
            "sf_op_wfold_finish")
    return dW


def wgrad(dY: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """fp32 split-K partials [splits, O, I] of G = dY^T . X (tcgen05, MN-major operands read in place)."""
    _req(dY, X)
    M, O = dY.shape
    I = X.shape[1]
    assert X.shape[0] == M and dY.stride(1) == 1 and X.stride(1) == 1
    splits = int(N.load().sf_op_wgrad_splits(M, O, I))
    parts = torch.empty(splits, O, I, dtype=torch.float32, device=dY.device)
    N.check(N.load().sf_op_wgrad(_stream(), sf_dtype(dY.dtype), dY.data_ptr(), dY.stride(0), X.data_ptr(), X.stride(0), M, O, I,
                                 parts.data_ptr()), "sf_op_wgrad")
    return parts


def wfold_finish_partials(parts: torch.Tensor, act_dtype: torch.dtype, out_dtype: torch.dtype, Wp: Optional[torch.Tensor] = None,
                          gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None,
                          dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(parts, Wp, gamma, beta, db, dgamma, dbeta)
    splits, O, I = parts.shape
    dW = torch.empty(O, I, dtype=out_dtype, device=parts.device)
    N.check(N.load().sf_op_wfold_finish_partials(_stream(), sf_dtype(act_dtype), parts.data_ptr(), splits, _p(Wp),
                                                 Wp.stride(0) if Wp is not None else 0, _p(gamma), _p(beta), _p(db), dW.data_ptr(),
                                                 sf_dtype(out_dtype), I, O, I, _p(dgamma), _p(dbeta)), "sf_op_wfold_finish_partials")
    return dW


def embed_table_grad(dx: torch.Tensor, B: int, T: int, S: int, mode: int, out: torch.Tensor, tidx: Optional[torch.Tensor] = None) -> None:
    _req(dx, out, tidx)
    assert out.dtype == torch.float32 and (tidx is None or tidx.dtype == torch.int32)
    D = dx.shape[-1]
    N.check(N.load().sf_op_embed_table_grad(_stream(), sf_dtype(dx.dtype), dx.data_ptr(), D, B, T, S, D, mode, _p(tidx), out.data_ptr()),
            "sf_op_embed_table_grad")


def rowperm(x: torch.Tensor, row_map: int, T: int, S: int) -> torch.Tensor:
    _req(x)
    assert x.is_contiguous() and x.dim() == 2
    out = torch.empty_like(x)
    N.check(N.load().sf_op_rowperm(_stream(), x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1] * x.element_size(), row_map, T, S),
            "sf_op_rowperm")
    return out


def attention_backward(mode: int, qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, groups: int, heads: int, L: int,
                       T_inner: int, causal: bool, scale: float, dqkv: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mode 0 temporal (groups = sites, L = T), mode 1 spatial (groups = frames, L = S)."""
    _req(qkv, out, dout, dqkv)
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    N.check(N.load().sf_op_attention_backward(_stream(), sf_dtype(qkv.dtype), mode, qkv.data_ptr(), qkv.stride(0), out.data_ptr(),
                                              out.stride(0), dout.data_ptr(), dout.stride(0), dqkv.data_ptr(), dqkv.stride(0), groups,
                                              heads, L, T_inner, int(causal), scale), "sf_op_attention_backward")
    return dqkv


def pool_attention_backward(kv: torch.Tensor, q: torch.Tensor, dout: torch.Tensor, frames: int, heads: int, S: int,
                            dq: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(kv, q, dout, dq)
    dkv = torch.empty_like(kv)
    N.check(N.load().sf_op_pool_attention_backward(_stream(), sf_dtype(kv.dtype), kv.data_ptr(), kv.stride(0), q.data_ptr(), dout.data_ptr(),
                                                   dout.stride(0), dkv.data_ptr(), dkv.stride(0), _p(dq), frames, heads, S),
            "sf_op_pool_attention_backward")
    return dkv
