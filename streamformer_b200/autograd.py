"""Differentiable encoder forward: the training step of the reference
(tools/finetune_tools.py:543-573, run_finetuning_multi_task.py:420-433) back-propagates through
``TimesformerMultiTaskingModelSigLIP.forward`` with torch.autograd; here the same gradients come from the
hand-written sm_100a kernels behind the C ABI, composed in ONE ``torch.autograd.Function`` around ``sf_forward``.

Host code stays Python (as the north star asks); every FLOP that scales with the batch runs in native kernels:
  * contractions: dgrad = dY . W and wgrad = dY^T . X on the tcgen05 GEMM (``sf_op_gemm``), wgrad operands made
    M-contiguous by ``sf_op_transpose``;
  * attention cores: ``sf_op_attention_backward`` (temporal-causal and spatial), ``sf_op_pool_attention_backward``;
  * row-wise: LayerNorm backward (folded / affine), GELU, gate, bias column sums, embedding-table sums.
Only parameter-sized glue (LoRA factor gradients, the constant probe query of the pooling head, concatenations)
is left to torch.

Memory/recompute: the forward keeps the L+1 layer-boundary activations (the ``hidden_states`` the forward can
already emit, 154 MB each at 32 clips per GPU) and the backward recomputes one layer's intermediates at a time
(activation checkpointing by layer) with the LayerNorms un-folded — the normalised rows are wgrad operands.

Scope of this round: one-shot forward (no KV cache), default resolution, attention groups of up to 208 tokens
(T <= 208 frames, S <= 208 patches).  Anything else raises instead of returning wrong gradients.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
from transformers.modeling_outputs import BaseModelOutputWithPooling

from . import _native as N
from . import ops
from .ops import sf_dtype

_LAYER_MATS = ["t_qkv", "t_out", "t_dense", "s_qkv", "s_out", "fc1", "fc2"]
_HEAD_MATS = ["head_kv", "head_out", "head_fc1", "head_fc2"]


class _TrainPack:
    """Packed matrices of the bound context as torch tensors (what the forward kernels consume: LayerNorm gamma
    folded in, LoRA merged) plus their transposes, the B operands of the dgrad GEMMs."""

    def __init__(self, eng, cfg):
        self.binds = eng.binds
        D, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        P = cfg.patch_size
        Kp = cfg.num_channels * P * P
        shapes = {"t_qkv": (3 * D, D), "t_out": (D, D), "t_dense": (D, D), "s_qkv": (3 * D, D), "s_out": (D, D), "fc1": (I, D),
                  "fc2": (D, I), "head_kv": (2 * D, D), "head_out": (D, D), "head_fc1": (I, D), "head_fc2": (D, I), "patch": (D, Kp)}
        dev, dt = eng.device, eng.dtype
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.layers: List[Dict[str, torch.Tensor]] = []

        def export(layer, name):
            rows, cols = shapes[name]
            w = torch.empty(rows, cols, dtype=dt, device=dev)
            wt = torch.empty(cols, rows, dtype=dt, device=dev)
            b = torch.empty(rows, dtype=torch.float32, device=dev)
            for t, nm, tr in ((w, name, 0), (wt, name, 1), (b, name + "_b", 0)):
                N.check(eng.lib.sf_export_packed(eng.handle, stream, layer, nm.encode(), tr, t.data_ptr(), t.numel() * t.element_size()),
                        "sf_export_packed")
            return w, wt, b

        for l in range(L):
            d = {}
            for name in _LAYER_MATS:
                d[name], d[name + "_T"], d[name + "_b"] = export(l, name)
            self.layers.append(d)
        self.head = {}
        for name in _HEAD_MATS + ["patch"]:
            self.head[name], self.head[name + "_T"], self.head[name + "_b"] = export(-1, name)
        q = torch.empty(D, dtype=torch.float32, device=dev)
        N.check(eng.lib.sf_export_packed(eng.handle, stream, -1, b"head_q", 0, q.data_ptr(), q.numel() * 4), "sf_export_packed")
        self.head["head_q"] = q
        self.unit = torch.ones(D, dtype=torch.float32, device=dev)
        self.zero = torch.zeros(D, dtype=torch.float32, device=dev)


def _train_pack(eng, cfg) -> _TrainPack:
    tp = getattr(eng, "_train_pack", None)
    if tp is None or tp.binds != eng.binds:
        tp = _TrainPack(eng, cfg)
        eng._train_pack = tp
    return tp


def _wgrad(dY: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """G[O, I] = dY^T . X  (dY [M, O], X [M, I]) on the tcgen05 GEMM: both operands made M-contiguous first."""
    return ops.gemm(ops.transpose(dY), ops.transpose(X))


def _vec(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return t.to(like.dtype).reshape(like.shape)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, eng, pixel_values, pix_dtype, H, W, names, *params):
        cfg = model.config
        B, T = pixel_values.shape[:2]
        P = cfg.patch_size
        S = (H // P) * (W // P)
        D, L = cfg.hidden_size, cfg.num_hidden_layers
        if S != model.embeddings.position_embeddings.shape[1] or H != W:
            raise NotImplementedError("training at a non-default resolution (gradients through the bicubic position-table "
                                      "resampling) is not implemented; run it under torch.no_grad()")
        if T > 208 or S > 208:
            raise NotImplementedError(f"backward supports attention groups of up to 208 tokens (T={T}, S={S})")
        dev = pixel_values.device
        last_hidden = torch.empty(B, T, S, D, dtype=eng.dtype, device=dev)
        pooled = torch.empty(B, T, D, dtype=eng.dtype, device=dev)
        hs = [torch.empty(B, S * T, D, dtype=eng.dtype, device=dev) for _ in range(L + 1)]
        ws = eng.get_workspace(B, T, H, W)
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(eng.lib.sf_forward(eng.handle, stream, pixel_values.data_ptr(), pix_dtype, B, T, H, W, last_hidden.data_ptr(),
                                   pooled.data_ptr(), N.ptr_array([t.data_ptr() for t in hs]), None, ws.data_ptr(), ws.numel()),
                "sf_forward")
        ctx.model, ctx.eng, ctx.names, ctx.pix_dtype = model, eng, names, pix_dtype
        ctx.geom = (B, T, S, H, W)
        ctx.save_for_backward(pixel_values, last_hidden, *hs)
        ctx.mark_non_differentiable(*hs)
        return (last_hidden, pooled, *hs)

    @staticmethod
    def backward(ctx, d_lhs, d_pooled, *_unused):
        model, eng, names = ctx.model, ctx.eng, ctx.names
        cfg = model.config
        pixel_values, last_hidden, *hs = ctx.saved_tensors
        B, T, S, H, W = ctx.geom
        D, I, L, heads = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.num_attention_heads
        M = B * S * T
        F = B * T
        dt = eng.dtype
        dev = last_hidden.device
        eps = float(cfg.layer_norm_eps)
        act = N.SF_ACT_GELU if cfg.hidden_act == "gelu" else N.SF_ACT_GELU_TANH
        causal = bool(cfg.enable_causal_temporal)
        tp = _train_pack(eng, cfg)
        params = dict(zip(names, [p for p in model.parameters()]))
        grads: Dict[str, Optional[torch.Tensor]] = {n: None for n in names}

        def wants(name):
            p = params.get(name)
            return p is not None and p.requires_grad

        def f32(n):
            return torch.zeros(n, dtype=torch.float32, device=dev)

        def put(name, g):
            if wants(name):
                grads[name] = _vec(g, params[name])

        def lin_grads(prefix, dY, X, pack=None, gamma_name=None, lora=None, w_key=None, lw=None):
            """Gradients of one nn.Linear (weight [+ LoRA factors], bias, and — when its input LayerNorm is folded into
            it — that LayerNorm's gamma / beta) from dY [M, O] and its input X [M, I] (the NORMALISED rows when folded)."""
            wname, bname = prefix + ".weight", prefix + ".bias"
            need_w = wants(wname) or (lora is not None and any(wants(n) for n in lora))
            need_ln = gamma_name is not None and (wants(gamma_name + ".weight") or wants(gamma_name + ".bias"))
            db = ops.colsum(dY) if (wants(bname) or need_ln or need_w and gamma_name is not None) else None
            if wants(bname):
                put(bname, db)
            if not (need_w or need_ln):
                return
            G = _wgrad(dY, X)
            pdt = params[wname].dtype
            if gamma_name is not None:
                gm = params[gamma_name + ".weight"].detach().float().contiguous()
                bt = params[gamma_name + ".bias"].detach().float().contiguous()
                dg, dbt = f32(gm.numel()), f32(gm.numel())
                dW = ops.wfold_finish(G, pdt, lw[w_key], gm, bt, db, dg, dbt)
                put(gamma_name + ".weight", dg)
                put(gamma_name + ".bias", dbt)
            else:
                dW = ops.wfold_finish(G, pdt)
            if wants(wname):
                grads[wname] = dW
            if lora is not None and all(n in params for n in lora):
                a_name, b_name = lora                       # W_merged = W + B . A  (…siglip.py:653-654, 749-751)
                A, Bm = params[a_name].detach(), params[b_name].detach()
                dWl = dW.to(dt).contiguous()
                if wants(b_name):                            # dB [O, r] = dW [O, I] . A^T
                    grads[b_name] = ops.gemm(dWl, A.to(dt).contiguous()).to(Bm.dtype)
                if wants(a_name):                            # dA [r, I] = B^T [r, O] . dW [O, I]
                    grads[a_name] = ops.gemm(ops.transpose(Bm.to(dt).contiguous()), ops.transpose(dWl)).to(A.dtype)

        # ------------------------------------------------------------------ pooling head + post_layernorm
        lhs2d = last_hidden.reshape(M, D)
        xL = hs[L].reshape(M, D)
        d_tokens = d_lhs.to(dt).contiguous().reshape(M, D) if d_lhs is not None else None
        hp = tp.head
        if d_pooled is not None:
            dp = d_pooled.to(dt).contiguous().reshape(F, D)
            kv = ops.gemm(lhs2d, hp["head_kv"], bias=hp["head_kv_b"])
            pc = ops.pool_attention(kv, hp["head_q"], F, heads, S)
            r = ops.gemm(pc, hp["head_out"], bias=hp["head_out_b"])
            g_h = params["head.layernorm.weight"].detach().float().contiguous()
            b_h = params["head.layernorm.bias"].detach().float().contiguous()
            lnr = ops.layernorm(r, g_h, b_h, eps)
            a1 = ops.gemm(lnr, hp["head_fc1"], bias=hp["head_fc1_b"])
            dhh = ops.gemm(dp, hp["head_fc2_T"])
            ops.gelu_backward_(a1, dhh, act)                 # a1 -> hidden, dhh -> dpre
            lin_grads("head.mlp.fc2", dp, a1)
            lin_grads("head.mlp.fc1", dhh, lnr)
            dlnr = ops.gemm(dhh, hp["head_fc1_T"])
            dg, dbt = f32(D), f32(D)
            dr = ops.ln_affine_backward(r, dlnr, g_h, eps, dg, dbt) + dp          # [frames, D]: parameter-scale glue
            put("head.layernorm.weight", dg)
            put("head.layernorm.bias", dbt)
            lin_grads("head.attention.out_proj", dr, pc)
            dpc = ops.gemm(dr, hp["head_out_T"])
            dq = f32(D)
            dkv = ops.pool_attention_backward(kv, hp["head_q"], dpc, F, heads, S, dq)
            # in_proj = [W_q; W_k; W_v]: K/V rows from the token GEMM, Q rows through the constant probe query
            # q = (W_q probe + b_q) / 8  (…siglip.py:1141-1148; F.multi_head_attention_forward)
            ipw, ipb, probe = params["head.attention.in_proj_weight"], params["head.attention.in_proj_bias"], params["head.probe"]
            need_ip = wants("head.attention.in_proj_weight") or wants("head.attention.in_proj_bias") or wants("head.probe")
            if need_ip:
                dq8 = dq * 0.125
                Gkv = _wgrad(dkv, lhs2d)
                dbkv = ops.colsum(dkv)
                pr = probe.detach().float().reshape(D)
                if wants("head.attention.in_proj_weight"):
                    grads["head.attention.in_proj_weight"] = torch.cat([torch.outer(dq8, pr).to(ipw.dtype), Gkv.to(ipw.dtype)], 0)
                if wants("head.attention.in_proj_bias"):
                    grads["head.attention.in_proj_bias"] = torch.cat([dq8, dbkv]).to(ipb.dtype)
                if wants("head.probe"):
                    grads["head.probe"] = (ipw.detach()[:D].float().t() @ dq8).to(probe.dtype).reshape(probe.shape)
            d_tokens = ops.gemm(dkv, hp["head_kv_T"], residual=d_tokens)
        if d_tokens is None:
            return (None,) * 7 + tuple(None for _ in names)
        g_p = params["post_layernorm.weight"].detach().float().contiguous()
        dg, dbt = f32(D), f32(D)
        dx = ops.ln_affine_backward(xL, d_tokens, g_p, eps, dg, dbt, N.SF_ROW_BNT_TO_BTN if T > 1 else N.SF_ROW_IDENTITY, T, S)
        put("post_layernorm.weight", dg)
        put("post_layernorm.bias", dbt)
        del d_tokens, lhs2d

        # ------------------------------------------------------------------ layers, last to first
        for l in range(L - 1, -1, -1):
            lw = tp.layers[l]
            p = f"encoder.layer.{l}."
            x0 = hs[l].reshape(M, D)
            gate = params[p + "temporal_attention_gating"].detach().float().reshape(1).contiguous()
            # ---- recompute the layer with the LayerNorms un-folded (…siglip.py:934-1004)
            n_t = ops.layernorm(x0, tp.unit, tp.zero, eps)
            qkv_t = ops.gemm(n_t, lw["t_qkv"], bias=lw["t_qkv_b"])
            ctx_t = ops.temporal_attention(qkv_t, B * S, heads, T, causal, 0.125)
            u_t = ops.gemm(ctx_t, lw["t_out"], bias=lw["t_out_b"])
            y_t = ops.gemm(u_t, lw["t_dense"], bias=lw["t_dense_b"])
            x1 = ops.gemm(u_t, lw["t_dense"], bias=lw["t_dense_b"], residual=x0, gate=gate)
            n_s = ops.layernorm(x1, tp.unit, tp.zero, eps)
            qkv_s = ops.gemm(n_s, lw["s_qkv"], bias=lw["s_qkv_b"])
            ctx_s = ops.spatial_attention(qkv_s, F, heads, S, 0.125, T_inner=T)
            x2 = ops.gemm(ctx_s, lw["s_out"], bias=lw["s_out_b"], residual=x1)
            n_a = ops.layernorm(x2, tp.unit, tp.zero, eps)
            a1 = ops.gemm(n_a, lw["fc1"], bias=lw["fc1_b"])
            # ---- MLP
            dh = ops.gemm(dx, lw["fc2_T"])
            ops.gelu_backward_(a1, dh, act)                  # a1 -> h, dh -> dpre
            lin_grads(p + "output.dense", dx, a1)
            lin_grads(p + "intermediate.dense", dh, n_a, gamma_name=p + "layernorm_after", w_key="fc1", lw=lw)
            dn = ops.gemm(dh, lw["fc1_T"])
            del dh, a1, n_a
            dx2 = ops.ln_backward(x2, dn, eps, dres=dx)
            # ---- spatial branch
            lin_grads(p + "attention.output.dense", dx2, ctx_s,
                      lora=(p + "attention.output.dense_lora_a.weight", p + "attention.output.dense_lora_b.weight"))
            dctx = ops.gemm(dx2, lw["s_out_T"])
            dqkv = ops.attention_backward(1, qkv_s, ctx_s, dctx, F, heads, S, T, False, 0.125)
            lin_grads(p + "attention.attention.qkv", dqkv, n_s, gamma_name=p + "layernorm_before", w_key="s_qkv", lw=lw,
                      lora=(p + "attention.attention.qkv_lora_a.weight", p + "attention.attention.qkv_lora_b.weight"))
            dn = ops.gemm(dqkv, lw["s_qkv_T"])
            del dqkv, qkv_s, ctx_s, n_s, x2
            dx1 = ops.ln_backward(x1, dn, eps, dres=dx2)
            del dx2
            # ---- temporal branch: x1 = x0 + tanh(g) * temporal_dense(out_proj(attn))   (…siglip.py:937-958)
            dgate = f32(1)
            dy = ops.gate_backward(dx1, y_t, gate, dgate)
            put(p + "temporal_attention_gating", dgate)
            lin_grads(p + "temporal_dense", dy, u_t)
            du = ops.gemm(dy, lw["t_dense_T"])
            lin_grads(p + "temporal_attention.output.dense", du, ctx_t)
            dctx = ops.gemm(du, lw["t_out_T"])
            dqkv = ops.attention_backward(0, qkv_t, ctx_t, dctx, B * S, heads, T, 1, causal, 0.125)
            lin_grads(p + "temporal_attention.attention.qkv", dqkv, n_t, gamma_name=p + "temporal_layernorm", w_key="t_qkv", lw=lw)
            dn = ops.gemm(dqkv, lw["t_qkv_T"])
            dx = ops.ln_backward(x0, dn, eps, dres=dx1)
            del dqkv, qkv_t, ctx_t, u_t, y_t, n_t, x1, dx1, dn, dy, du, dctx

        # ------------------------------------------------------------------ embeddings (…siglip.py:413-457)
        emb = "embeddings."
        if wants(emb + "position_embeddings"):
            g = torch.zeros(S, D, dtype=torch.float32, device=dev)
            ops.embed_table_grad(dx, B, T, S, 0, g)
            put(emb + "position_embeddings", g)
        if wants(emb + "time_embeddings"):
            Fr = params[emb + "time_embeddings"].shape[1]
            if T <= Fr:
                tidx = torch.arange(T, dtype=torch.int32, device=dev)                 # [:, :T] slice (…siglip.py:436-439)
            else:                                                                     # nearest map (…siglip.py:441-447)
                tidx = torch.clamp(torch.floor(torch.arange(T, dtype=torch.float32, device=dev) * (float(Fr) / float(T))), max=Fr - 1).to(torch.int32)
            g = torch.zeros(Fr, D, dtype=torch.float32, device=dev)
            ops.embed_table_grad(dx, B, T, S, 1, g, tidx)
            put(emb + "time_embeddings", g)
        wn, bn = emb + "patch_embeddings.projection.weight", emb + "patch_embeddings.projection.bias"
        if wants(wn) or wants(bn):
            dxp = ops.rowperm(dx, N.SF_ROW_BNT_TO_BTN, T, S) if T > 1 else dx           # rows (b,t,n): the patch GEMM's order
            if wants(bn):
                put(bn, ops.colsum(dxp))
            if wants(wn):
                px = pixel_values if ctx.pix_dtype not in (N.SF_U8, N.SF_U8_HWC) else None
                if px is None:
                    raise NotImplementedError("gradient of the patch projection with uint8 inputs: pass float pixels for training")
                patches = ops.im2col(px.reshape(F, cfg.num_channels, H, W), cfg.patch_size, dt)
                grads[wn] = ops.wfold_finish(_wgrad(dxp, patches), params[wn].dtype).reshape(params[wn].shape)
        return (None,) * 7 + tuple(grads[n] for n in names)


def encoder_forward_with_grad(model, eng, pixel_values, pix_dtype, H, W, output_hidden_states, return_dict):
    """Training-mode ``model(pixel_values)``: same outputs as the inference path, differentiable w.r.t. every
    parameter that requires grad (``last_hidden_state`` and ``pooler_output``; ``hidden_states`` are returned detached)."""
    named = [(n, p) for n, p in model.named_parameters()]
    names = tuple(n for n, _ in named)
    out = _EncoderFn.apply(model, eng, pixel_values, pix_dtype, H, W, names, *[p for _, p in named])
    last_hidden, pooled, hs = out[0], out[1], out[2:]
    out_dtype = model.embeddings.position_embeddings.dtype
    if out_dtype != eng.dtype:
        last_hidden, pooled = last_hidden.to(out_dtype), pooled.to(out_dtype)
        hs = tuple(h.to(out_dtype) for h in hs)
    hs_t = tuple(hs) if output_hidden_states else None
    if not return_dict:
        return tuple(v for v in [last_hidden, hs_t] if v is not None)
    return BaseModelOutputWithPooling(last_hidden_state=last_hidden, pooler_output=pooled, hidden_states=hs_t, attentions=None)
