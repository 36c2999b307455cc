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
    folded in, LoRA merged) plus their transposes, the B operands of the dgrad GEMMs.  Exported lazily per
    parameter group, so a stand-alone encoder / head / embedding module exports only what it owns."""

    def __init__(self, eng, cfg):
        self.eng, self.binds = eng, eng.binds
        D, I = cfg.hidden_size, cfg.intermediate_size
        P = cfg.patch_size[0] if isinstance(cfg.patch_size, (tuple, list)) else cfg.patch_size
        Kp = cfg.num_channels * P * P
        self.shapes = {"t_qkv": (3 * D, D), "t_out": (D, D), "t_dense": (D, D), "s_qkv": (3 * D, D), "s_out": (D, D), "fc1": (I, D),
                       "fc2": (D, I), "head_kv": (2 * D, D), "head_out": (D, D), "head_fc1": (I, D), "head_fc2": (D, I), "patch": (D, Kp)}
        self.D = D
        self._layers: Dict[int, Dict[str, torch.Tensor]] = {}
        self._head: Optional[Dict[str, torch.Tensor]] = None
        self._patch: Optional[Dict[str, torch.Tensor]] = None
        self.unit = torch.ones(D, dtype=torch.float32, device=eng.device)
        self.zero = torch.zeros(D, dtype=torch.float32, device=eng.device)

    def _export(self, layer, name, d):
        eng = self.eng
        rows, cols = self.shapes[name]
        dev, dt = eng.device, eng.dtype
        stream = torch.cuda.current_stream(dev).cuda_stream
        w = torch.empty(rows, cols, dtype=dt, device=dev)
        wt = torch.empty(cols, rows, dtype=dt, device=dev)
        b = torch.empty(rows, dtype=torch.float32, device=dev)
        for t, nm, tr in ((w, name, 0), (wt, name, 1), (b, name + "_b", 0)):
            N.check(eng.lib.sf_export_packed(eng.handle, stream, layer, nm.encode(), tr, t.data_ptr(), t.numel() * t.element_size()),
                    "sf_export_packed")
        d[name], d[name + "_T"], d[name + "_b"] = w, wt, b

    def layer(self, l: int) -> Dict[str, torch.Tensor]:
        if l not in self._layers:
            d: Dict[str, torch.Tensor] = {}
            for name in _LAYER_MATS:
                self._export(l, name, d)
            self._layers[l] = d
        return self._layers[l]

    def head(self) -> Dict[str, torch.Tensor]:
        if self._head is None:
            d: Dict[str, torch.Tensor] = {}
            for name in _HEAD_MATS:
                self._export(-1, name, d)
            q = torch.empty(self.D, dtype=torch.float32, device=self.eng.device)
            stream = torch.cuda.current_stream(self.eng.device).cuda_stream
            N.check(self.eng.lib.sf_export_packed(self.eng.handle, stream, -1, b"head_q", 0, q.data_ptr(), q.numel() * 4), "sf_export_packed")
            d["head_q"] = q
            self._head = d
        return self._head


def _train_pack(eng, cfg) -> _TrainPack:
    tp = getattr(eng, "_train_pack", None)
    if tp is None or tp.binds != eng.binds:
        tp = _TrainPack(eng, cfg)
        eng._train_pack = tp
    return tp


def _wgrad(dY: torch.Tensor, X: torch.Tensor, out_dtype: torch.dtype, **fold) -> torch.Tensor:
    """dW [O, I] = dY^T . X  (dY [M, O], X [M, I]): the tcgen05 wgrad kernel reads both activations MN-major in place and
    leaves fp32 split-K partials; wfold_finish sums them, applies the LayerNorm fold (``fold``: Wp, gamma, beta, db,
    dgamma, dbeta) and casts to the parameter dtype."""
    return ops.wfold_finish_partials(ops.wgrad(dY, X), dY.dtype, out_dtype, **fold)


class _Bwd:
    """State of one backward pass: the engine, its training pack, the parameters by reference state-dict name and
    the gradients collected so far."""

    def __init__(self, root, eng, cfg, names):
        self.eng, self.cfg, self.names = eng, cfg, names
        self.tp = _train_pack(eng, cfg)
        self.params = dict(zip(names, [p for p in root.parameters()]))
        self.grads: Dict[str, Optional[torch.Tensor]] = {n: None for n in names}
        self.dt, self.dev = eng.dtype, eng.device
        self.eps = float(cfg.layer_norm_eps)
        self.act = N.SF_ACT_GELU if cfg.hidden_act == "gelu" else N.SF_ACT_GELU_TANH
        self.causal = bool(cfg.enable_causal_temporal)
        self.D, self.I, self.heads = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads

    def wants(self, name):
        p = self.params.get(name)
        return p is not None and p.requires_grad

    def f32(self, n):
        return torch.zeros(n, dtype=torch.float32, device=self.dev)

    def put(self, name, g):
        if self.wants(name):
            like = self.params[name]
            self.grads[name] = g.to(like.dtype).reshape(like.shape)

    def result(self):
        return tuple(self.grads[n] for n in self.names)

    def lin_grads(self, prefix, dY, X, gamma_name=None, lora=None, packed=None):
        """Gradients of one nn.Linear (weight [+ LoRA factors], bias, and — when its input LayerNorm is folded into
        it — that LayerNorm's gamma / beta) from dY [M, O] and its input X [M, I] (the NORMALISED rows when folded;
        `packed` is then the gamma-scaled matrix the forward used)."""
        wants, params, grads, dt = self.wants, self.params, self.grads, self.dt
        wname, bname = prefix + ".weight", prefix + ".bias"
        need_w = wants(wname) or (lora is not None and any(wants(n) for n in lora))
        need_ln = gamma_name is not None and (wants(gamma_name + ".weight") or wants(gamma_name + ".bias"))
        folded = gamma_name is not None
        db = ops.colsum(dY) if (wants(bname) or need_ln or (need_w and folded)) else None
        if wants(bname):
            self.put(bname, db)
        if not (need_w or need_ln):
            return
        pdt = params[wname].dtype
        if folded:
            gm = params[gamma_name + ".weight"].detach().float().contiguous()
            bt = params[gamma_name + ".bias"].detach().float().contiguous()
            dg, dbt = self.f32(gm.numel()), self.f32(gm.numel())
            dW = _wgrad(dY, X, pdt, Wp=packed, gamma=gm, beta=bt, db=db, dgamma=dg, dbeta=dbt)
            self.put(gamma_name + ".weight", dg)
            self.put(gamma_name + ".bias", dbt)
        else:
            dW = _wgrad(dY, X, pdt)
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

    # ------------------------------------------------------------------ one divided space-time block
    _ACTS = ("n_t", "qkv_t", "ctx_t", "u_t", "y_t", "x1", "n_s", "qkv_s", "ctx_s", "x2", "n_a", "a1")

    def layer_forward(self, l, x0, B, T, S, want_output=False):
        """The layer with its LayerNorms un-folded (…siglip.py:934-1004): every intermediate the backward needs (the
        normalised rows are wgrad operands, a1 is the MLP pre-activation).  With want_output also the layer output."""
        lw, tp, heads, eps = self.tp.layer(l), self.tp, self.heads, self.eps
        gate = self.params[f"encoder.layer.{l}.temporal_attention_gating"].detach().float().reshape(1).contiguous()
        a = {}
        a["n_t"] = ops.layernorm(x0, tp.unit, tp.zero, eps)
        a["qkv_t"] = ops.gemm(a["n_t"], lw["t_qkv"], bias=lw["t_qkv_b"])
        a["ctx_t"] = ops.temporal_attention(a["qkv_t"], B * S, heads, T, self.causal, 0.125)
        a["u_t"] = ops.gemm(a["ctx_t"], lw["t_out"], bias=lw["t_out_b"])
        a["y_t"] = ops.gemm(a["u_t"], lw["t_dense"], bias=lw["t_dense_b"])
        a["x1"] = ops.gemm(a["u_t"], lw["t_dense"], bias=lw["t_dense_b"], residual=x0, gate=gate)
        a["n_s"] = ops.layernorm(a["x1"], tp.unit, tp.zero, eps)
        a["qkv_s"] = ops.gemm(a["n_s"], lw["s_qkv"], bias=lw["s_qkv_b"])
        a["ctx_s"] = ops.spatial_attention(a["qkv_s"], B * T, heads, S, 0.125, T_inner=T)
        a["x2"] = ops.gemm(a["ctx_s"], lw["s_out"], bias=lw["s_out_b"], residual=a["x1"])
        a["n_a"] = ops.layernorm(a["x2"], tp.unit, tp.zero, eps)
        a["a1"] = ops.gemm(a["n_a"], lw["fc1"], bias=lw["fc1_b"])
        out = None
        if want_output:
            out = ops.gemm(ops.gelu(a["a1"], self.act), lw["fc2"], bias=lw["fc2_b"], residual=a["x2"])
        return a, out

    def layer(self, l, x0, dx, B, T, S, acts=None):
        """dx (gradient w.r.t. the layer's output [M, D]) -> gradient w.r.t. its input x0; parameter gradients collected.
        acts: the intermediates kept by the forward, or None to recompute them (activation checkpointing by layer)."""
        lw = self.tp.layer(l)
        p = f"encoder.layer.{l}."
        heads, eps = self.heads, self.eps
        F = B * T
        gate = self.params[p + "temporal_attention_gating"].detach().float().reshape(1).contiguous()
        if acts is None:
            acts, _ = self.layer_forward(l, x0, B, T, S)
        n_t, qkv_t, ctx_t, u_t, y_t, x1, n_s, qkv_s, ctx_s, x2, n_a, a1 = [acts[k] for k in self._ACTS]
        del acts
        # ---- MLP
        dh = ops.gemm(dx, lw["fc2_T"])
        ops.gelu_backward_(a1, dh, self.act)                  # a1 -> h, dh -> dpre
        self.lin_grads(p + "output.dense", dx, a1)
        self.lin_grads(p + "intermediate.dense", dh, n_a, gamma_name=p + "layernorm_after", packed=lw["fc1"])
        dn = ops.gemm(dh, lw["fc1_T"])
        del dh, a1, n_a
        dx2 = ops.ln_backward(x2, dn, eps, dres=dx)
        # ---- spatial branch
        self.lin_grads(p + "attention.output.dense", dx2, ctx_s,
                       lora=(p + "attention.output.dense_lora_a.weight", p + "attention.output.dense_lora_b.weight"))
        dctx = ops.gemm(dx2, lw["s_out_T"])
        dqkv = ops.attention_backward(1, qkv_s, ctx_s, dctx, F, heads, S, T, False, 0.125)
        self.lin_grads(p + "attention.attention.qkv", dqkv, n_s, gamma_name=p + "layernorm_before", packed=lw["s_qkv"],
                       lora=(p + "attention.attention.qkv_lora_a.weight", p + "attention.attention.qkv_lora_b.weight"))
        dn = ops.gemm(dqkv, lw["s_qkv_T"])
        del dqkv, qkv_s, ctx_s, n_s, x2
        dx1 = ops.ln_backward(x1, dn, eps, dres=dx2)
        del dx2
        # ---- temporal branch: x1 = x0 + tanh(g) * temporal_dense(out_proj(attn))   (…siglip.py:937-958)
        dgate = self.f32(1)
        dy = ops.gate_backward(dx1, y_t, gate, dgate)
        self.put(p + "temporal_attention_gating", dgate)
        self.lin_grads(p + "temporal_dense", dy, u_t)
        du = ops.gemm(dy, lw["t_dense_T"])
        self.lin_grads(p + "temporal_attention.output.dense", du, ctx_t)
        dctx = ops.gemm(du, lw["t_out_T"])
        dqkv = ops.attention_backward(0, qkv_t, ctx_t, dctx, B * S, heads, T, 1, self.causal, 0.125)
        self.lin_grads(p + "temporal_attention.attention.qkv", dqkv, n_t, gamma_name=p + "temporal_layernorm", packed=lw["t_qkv"])
        dn = ops.gemm(dqkv, lw["t_qkv_T"])
        return ops.ln_backward(x0, dn, eps, dres=dx1)

    # ------------------------------------------------------------------ pooling head (…siglip.py:1141-1154)
    def head(self, tokens2d, dp, F, S, d_tokens=None):
        """dp [F, D] (gradient w.r.t. the pooled output) -> gradient w.r.t. the tokens [F*S, D] (added to d_tokens)."""
        hp = self.tp.head()
        params, D, heads, eps = self.params, self.D, self.heads, self.eps
        kv = ops.gemm(tokens2d, hp["head_kv"], bias=hp["head_kv_b"])
        pc = ops.pool_attention(kv, hp["head_q"], F, heads, S)
        r = ops.gemm(pc, hp["head_out"], bias=hp["head_out_b"])
        g_h = params["head.layernorm.weight"].detach().float().contiguous()
        b_h = params["head.layernorm.bias"].detach().float().contiguous()
        lnr = ops.layernorm(r, g_h, b_h, eps)
        a1 = ops.gemm(lnr, hp["head_fc1"], bias=hp["head_fc1_b"])
        dhh = ops.gemm(dp, hp["head_fc2_T"])
        ops.gelu_backward_(a1, dhh, self.act)                 # a1 -> hidden, dhh -> dpre
        self.lin_grads("head.mlp.fc2", dp, a1)
        self.lin_grads("head.mlp.fc1", dhh, lnr)
        dlnr = ops.gemm(dhh, hp["head_fc1_T"])
        dg, dbt = self.f32(D), self.f32(D)
        dr = ops.ln_affine_backward(r, dlnr, g_h, eps, dg, dbt) + dp          # [frames, D]: parameter-scale glue
        self.put("head.layernorm.weight", dg)
        self.put("head.layernorm.bias", dbt)
        self.lin_grads("head.attention.out_proj", dr, pc)
        dpc = ops.gemm(dr, hp["head_out_T"])
        dq = self.f32(D)
        dkv = ops.pool_attention_backward(kv, hp["head_q"], dpc, F, heads, S, dq)
        # in_proj = [W_q; W_k; W_v]: K/V rows from the token GEMM, Q rows through the constant probe query
        # q = (W_q probe + b_q) / 8  (…siglip.py:1141-1148; F.multi_head_attention_forward)
        wn, bn, pn = "head.attention.in_proj_weight", "head.attention.in_proj_bias", "head.probe"
        if self.wants(wn) or self.wants(bn) or self.wants(pn):
            ipw, ipb, probe = params[wn], params[bn], params[pn]
            dq8 = dq * 0.125
            pr = probe.detach().float().reshape(D)
            if self.wants(wn):
                self.grads[wn] = torch.cat([torch.outer(dq8, pr).to(ipw.dtype), _wgrad(dkv, tokens2d, ipw.dtype)], 0)
            if self.wants(bn):
                self.grads[bn] = torch.cat([dq8, ops.colsum(dkv)]).to(ipb.dtype)
            if self.wants(pn):
                self.grads[pn] = (ipw.detach()[:D].float().t() @ dq8).to(probe.dtype).reshape(probe.shape)
        return ops.gemm(dkv, hp["head_kv_T"], residual=d_tokens)

    # ------------------------------------------------------------------ embeddings (…siglip.py:413-457)
    def embeddings(self, dx, pixel_values, pix_dtype, B, T, S, H, W):
        cfg, D, dev = self.cfg, self.D, self.dev
        emb = "embeddings."
        if self.wants(emb + "position_embeddings"):
            g = torch.zeros(S, D, dtype=torch.float32, device=dev)
            ops.embed_table_grad(dx, B, T, S, 0, g)
            self.put(emb + "position_embeddings", g)
        if self.wants(emb + "time_embeddings"):
            Fr = self.params[emb + "time_embeddings"].shape[1]
            if T <= Fr:
                tidx = torch.arange(T, dtype=torch.int32, device=dev)                 # [:, :T] slice (…siglip.py:436-439)
            else:                                                                     # nearest map (…siglip.py:441-447)
                tidx = torch.clamp(torch.floor(torch.arange(T, dtype=torch.float32, device=dev) * (float(Fr) / float(T))), max=Fr - 1).to(torch.int32)
            g = torch.zeros(Fr, D, dtype=torch.float32, device=dev)
            ops.embed_table_grad(dx, B, T, S, 1, g, tidx)
            self.put(emb + "time_embeddings", g)
        wn, bn = emb + "patch_embeddings.projection.weight", emb + "patch_embeddings.projection.bias"
        if self.wants(wn) or self.wants(bn):
            dxp = ops.rowperm(dx, N.SF_ROW_BNT_TO_BTN, T, S) if T > 1 else dx           # rows (b,t,n): the patch GEMM's order
            if self.wants(bn):
                self.put(bn, ops.colsum(dxp))
            if self.wants(wn):
                if pix_dtype in (N.SF_U8, N.SF_U8_HWC):
                    raise NotImplementedError("gradient of the patch projection with uint8 inputs: pass float pixels for training")
                P = cfg.patch_size[0] if isinstance(cfg.patch_size, (tuple, list)) else cfg.patch_size
                patches = ops.im2col(pixel_values.reshape(B * T, cfg.num_channels, H, W), P, self.dt)
                self.grads[wn] = _wgrad(dxp, patches, self.params[wn].dtype).reshape(self.params[wn].shape)


def _check_geometry(T, S):
    if T > 208 or S > 208:
        raise NotImplementedError(f"backward supports attention groups of up to 208 tokens (T={T}, S={S})")


def _names(root, eng):
    return tuple(eng.prefix + n for n, _ in root.named_parameters())


def _keep_activations(cfg, M, L, dev) -> bool:
    """Keep every layer's intermediates for the backward (29 KB per token row and layer: 35 GB at 32 clips per GPU —
    a B200 has 180 GB) instead of recomputing them, unless config.training_recompute forces one way or memory is short."""
    mode = str(getattr(cfg, "training_recompute", "auto"))
    if mode in ("always", "True", "true", "1"):
        return False
    if mode in ("never", "False", "false", "0"):
        return True
    need = 2 * M * L * (9 * cfg.hidden_size + 6 * cfg.hidden_size + cfg.intermediate_size)
    free, _total = torch.cuda.mem_get_info(dev)
    return need < 0.45 * free


class _EncoderFn(torch.autograd.Function):
    """TimesformerMultiTaskingModelSigLIP.forward: pixels -> (last_hidden_state, pooler_output)."""

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
        _check_geometry(T, S)
        dev = pixel_values.device
        M = B * S * T
        last_hidden = torch.empty(B, T, S, D, dtype=eng.dtype, device=dev)
        pooled = torch.empty(B, T, D, dtype=eng.dtype, device=dev)
        hs = [torch.empty(B, S * T, D, dtype=eng.dtype, device=dev) for _ in range(L + 1)]
        ws = eng.get_workspace(B, T, H, W)
        stream = torch.cuda.current_stream(dev).cuda_stream
        keep = _keep_activations(cfg, M, L, dev)
        acts_flat = []
        if keep:
            # training forward with the LayerNorms un-folded, every intermediate kept: no recompute in the backward
            bw = _Bwd(model, eng, cfg, names)
            N.check(eng.lib.sf_embed_forward(eng.handle, stream, pixel_values.data_ptr(), pix_dtype, B, T, H, W, 0, T, hs[0].data_ptr(),
                                             ws.data_ptr(), ws.numel()), "sf_embed_forward")
            for l in range(L):
                acts, out = bw.layer_forward(l, hs[l].reshape(M, D), B, T, S, want_output=True)
                hs[l + 1] = out.reshape(B, S * T, D)
                acts_flat.extend(acts[k] for k in _Bwd._ACTS)
            N.check(eng.lib.sf_final_norm(eng.handle, stream, hs[L].data_ptr(), B, T, S, last_hidden.data_ptr()), "sf_final_norm")
            N.check(eng.lib.sf_head_forward(eng.handle, stream, last_hidden.data_ptr(), B * T, S, pooled.data_ptr(), ws.data_ptr(),
                                            ws.numel()), "sf_head_forward")
        else:
            N.check(eng.lib.sf_forward(eng.handle, stream, pixel_values.data_ptr(), pix_dtype, B, T, H, W, last_hidden.data_ptr(),
                                       pooled.data_ptr(), N.ptr_array([t.data_ptr() for t in hs]), None, ws.data_ptr(), ws.numel()),
                    "sf_forward")
        ctx.model, ctx.eng, ctx.names, ctx.pix_dtype = model, eng, names, pix_dtype
        ctx.geom = (B, T, S, H, W)
        ctx.kept = keep
        ctx.save_for_backward(pixel_values, last_hidden, *hs, *acts_flat)
        ctx.mark_non_differentiable(*hs)
        return (last_hidden, pooled, *hs)

    @staticmethod
    def backward(ctx, d_lhs, d_pooled, *_unused):
        model, eng = ctx.model, ctx.eng
        L = model.config.num_hidden_layers
        saved = ctx.saved_tensors
        pixel_values, last_hidden = saved[0], saved[1]
        hs = saved[2:2 + L + 1]
        acts_flat = list(saved[2 + L + 1:])
        B, T, S, H, W = ctx.geom
        bw = _Bwd(model, eng, model.config, ctx.names)
        D = bw.D
        M, F = B * S * T, B * T
        nact = len(_Bwd._ACTS)
        d_tokens = d_lhs.to(bw.dt).contiguous().reshape(M, D) if d_lhs is not None else None
        if d_pooled is not None:
            d_tokens = bw.head(last_hidden.reshape(M, D), d_pooled.to(bw.dt).contiguous().reshape(F, D), F, S, d_tokens)
        if d_tokens is None:
            return (None,) * 7 + bw.result()
        g_p = bw.params["post_layernorm.weight"].detach().float().contiguous()
        dg, dbt = bw.f32(D), bw.f32(D)
        dx = ops.ln_affine_backward(hs[L].reshape(M, D), d_tokens, g_p, bw.eps, dg, dbt,
                                    N.SF_ROW_BNT_TO_BTN if T > 1 else N.SF_ROW_IDENTITY, T, S)
        bw.put("post_layernorm.weight", dg)
        bw.put("post_layernorm.bias", dbt)
        del d_tokens
        for l in range(L - 1, -1, -1):
            acts = dict(zip(_Bwd._ACTS, acts_flat[l * nact:(l + 1) * nact])) if ctx.kept else None
            dx = bw.layer(l, hs[l].reshape(M, D), dx, B, T, S, acts)
        bw.embeddings(dx, pixel_values, ctx.pix_dtype, B, T, S, H, W)
        return (None,) * 7 + bw.result()


def encoder_forward_with_grad(model, eng, pixel_values, pix_dtype, H, W, output_hidden_states, return_dict):
    """Training-mode ``model(pixel_values)``: same outputs as the inference path, differentiable w.r.t. every
    parameter that requires grad (``last_hidden_state`` and ``pooler_output``; ``hidden_states`` are returned detached)."""
    names = _names(model, eng)
    out = _EncoderFn.apply(model, eng, pixel_values, pix_dtype, H, W, names, *list(model.parameters()))
    last_hidden, pooled, hs = out[0], out[1], out[2:]
    out_dtype = model.embeddings.position_embeddings.dtype
    if out_dtype != eng.dtype:
        last_hidden, pooled = last_hidden.to(out_dtype), pooled.to(out_dtype)
        hs = tuple(h.to(out_dtype) for h in hs)
    hs_t = tuple(hs) if output_hidden_states else None
    if not return_dict:
        return tuple(v for v in [last_hidden, hs_t] if v is not None)
    return BaseModelOutputWithPooling(last_hidden_state=last_hidden, pooler_output=pooled, hidden_states=hs_t, attentions=None)


# ------------------------------------------------------------------------------------- block-level API
# The sub-modules downstream code composes on its own (downstream/AR/models/modeling_timesformer_video_classification.py:
# 42-133 fine-tunes embeddings -> encoder -> its own torch post_layernorm -> pooling head): each is differentiable
# w.r.t. its input activations and its own parameters.
class _StackFn(torch.autograd.Function):
    """TimesformerEncoder.forward (all layers) or one TimesformerLayerSigLIP (layers = [index])."""

    @staticmethod
    def forward(ctx, root, eng, names, layers, num_frames, x, *params):
        cfg = root.config
        B, NT, D = x.shape
        T = num_frames
        S = NT // T
        _check_geometry(T, S)
        x = x.contiguous()
        hs = [x] + [torch.empty_like(x) for _ in layers]
        P = cfg.patch_size[0] if isinstance(cfg.patch_size, (tuple, list)) else cfg.patch_size
        ws = eng.get_workspace(B, T, P, P * S)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        if len(layers) == cfg.num_hidden_layers and len(layers) > 1:
            N.check(eng.lib.sf_encoder_forward(eng.handle, stream, x.data_ptr(), B, T, S, None, hs[-1].data_ptr(),
                                               N.ptr_array([t.data_ptr() for t in hs]), None, ws.data_ptr(), ws.numel()), "sf_encoder_forward")
        else:
            for i, l in enumerate(layers):
                N.check(eng.lib.sf_layer_forward(eng.handle, stream, l, hs[i].data_ptr(), hs[i + 1].data_ptr(), B, T, S, None, None,
                                                 ws.data_ptr(), ws.numel()), "sf_layer_forward")
        ctx.root, ctx.eng, ctx.names, ctx.layers, ctx.geom = root, eng, names, layers, (B, T, S)
        ctx.save_for_backward(*hs)
        ctx.mark_non_differentiable(*hs[1:-1])
        return tuple(hs[1:][::-1])            # (final output, ..., first layer's output)

    @staticmethod
    def backward(ctx, d_out, *_unused):
        hs = ctx.saved_tensors
        B, T, S = ctx.geom
        bw = _Bwd(ctx.root, ctx.eng, ctx.root.config, ctx.names)
        M = B * S * T
        dx = d_out.to(bw.dt).contiguous().reshape(M, bw.D)
        for i in range(len(ctx.layers) - 1, -1, -1):
            dx = bw.layer(ctx.layers[i], hs[i].reshape(M, bw.D), dx, B, T, S)
        return (None,) * 5 + (dx.reshape(B, S * T, bw.D),) + bw.result()


def stack_forward_with_grad(root, eng, layers, hidden_states, num_frames):
    """Differentiable layers over ``hidden_states`` [B, N*T, D]: returns the list of layer outputs (last = result)."""
    names = _names(root, eng)
    outs = _StackFn.apply(root, eng, names, tuple(layers), num_frames, hidden_states.to(eng.dtype), *list(root.parameters()))
    return list(outs[::-1])


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, root, eng, names, pixel_values, pix_dtype, H, W, *params):
        cfg = root.config
        B, T = pixel_values.shape[:2]
        P = cfg.patch_size[0] if isinstance(cfg.patch_size, (tuple, list)) else cfg.patch_size
        S = (H // P) * (W // P)
        x = torch.empty(B, S * T, cfg.hidden_size, dtype=eng.dtype, device=pixel_values.device)
        ws = eng.get_workspace(B, T, H, W)
        N.check(eng.lib.sf_embed_forward(eng.handle, torch.cuda.current_stream(x.device).cuda_stream, pixel_values.data_ptr(), pix_dtype,
                                         B, T, H, W, 0, T, x.data_ptr(), ws.data_ptr(), ws.numel()), "sf_embed_forward")
        ctx.root, ctx.eng, ctx.names, ctx.pix_dtype, ctx.geom = root, eng, names, pix_dtype, (B, T, S, H, W)
        ctx.save_for_backward(pixel_values)
        return x

    @staticmethod
    def backward(ctx, dx):
        (pixel_values,) = ctx.saved_tensors
        B, T, S, H, W = ctx.geom
        bw = _Bwd(ctx.root, ctx.eng, ctx.root.config, ctx.names)
        bw.embeddings(dx.to(bw.dt).contiguous().reshape(B * S * T, bw.D), pixel_values, ctx.pix_dtype, B, T, S, H, W)
        return (None,) * 7 + bw.result()


def embed_forward_with_grad(root, eng, emb, pixel_values, pix_dtype, H, W):
    S = (H // emb.patch_embeddings.patch_size[0]) * (W // emb.patch_embeddings.patch_size[1])
    if S != emb.position_embeddings.shape[1] or H != W:
        raise NotImplementedError("training at a non-default resolution is not implemented; run it under torch.no_grad()")
    return _EmbedFn.apply(root, eng, _names(root, eng), pixel_values, pix_dtype, H, W, *list(root.parameters()))


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, root, eng, names, tokens, *params):
        frames, S, D = tokens.shape
        tokens = tokens.contiguous()
        out = torch.empty(frames, D, dtype=eng.dtype, device=tokens.device)
        cfg = root.config
        P = cfg.patch_size[0] if isinstance(cfg.patch_size, (tuple, list)) else cfg.patch_size
        ws = eng.get_workspace(frames, 1, P, P * S)
        N.check(eng.lib.sf_head_forward(eng.handle, torch.cuda.current_stream(tokens.device).cuda_stream, tokens.data_ptr(), frames, S,
                                        out.data_ptr(), ws.data_ptr(), ws.numel()), "sf_head_forward")
        ctx.root, ctx.eng, ctx.names = root, eng, names
        ctx.save_for_backward(tokens)
        return out

    @staticmethod
    def backward(ctx, dp):
        (tokens,) = ctx.saved_tensors
        frames, S, D = tokens.shape
        bw = _Bwd(ctx.root, ctx.eng, ctx.root.config, ctx.names)
        d_tokens = bw.head(tokens.reshape(frames * S, D), dp.to(bw.dt).contiguous(), frames, S)
        return (None,) * 3 + (d_tokens.reshape(frames, S, D),) + bw.result()


def head_forward_with_grad(root, eng, tokens):
    return _HeadFn.apply(root, eng, _names(root, eng), tokens.to(eng.dtype), *list(root.parameters()))
