"""HF-style boundary of the B200-native StreamFormer encoder.

Same class names, constructor, ``from_pretrained()/forward()`` signature, outputs and state-dict
names as the reference (models/modeling_timesformer_siglip.py:300-1354 and the KV-cache twin
downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py:194-1392), so
``run_finetuning_multi_task.py``, ``extract_oad_feature.py`` and the ``downstream/`` pipelines can
import it unchanged.  The modules below only *hold parameters under the reference's names*; all
arithmetic happens in the hand-written sm_100a kernels behind the C ABI
(include/streamformer_b200.h) — there is no eager / CPU fallback, and a forward on a non-CUDA
tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import torch
from torch import nn
from transformers.modeling_outputs import BaseModelOutput, BaseModelOutputWithPooling, ModelOutput
from transformers.modeling_utils import PreTrainedModel

from . import _native as N
from .configuration_streamformer import StreamformerConfig
from .ops import sf_dtype

__all__ = [
    "StreamformerConfig",
    "TimesformerPreTrainedModel",
    "TimesformerPatchEmbeddings",
    "TimesformerEmbeddingsSigLIP",
    "TimesformerCausalSelfAttention",
    "TimesformerSelfAttention",
    "TimesformerSelfOutput",
    "TimeSformerCausalAttention",
    "TimeSformerAttention",
    "TimesformerIntermediate",
    "TimesformerOutput",
    "TimesformerLayerSigLIP",
    "TimesformerEncoder",
    "SiglipMLP",
    "TimesformerSiglipMultiheadAttentionPoolingHead",
    "TimesformerMultiTaskingModelSigLIP",
    "StreamformerKVCache",
    "StreamformerOutputWithPast",
]

_ACTS = {"gelu": N.SF_ACT_GELU, "gelu_pytorch_tanh": N.SF_ACT_GELU_TANH, "gelu_new": N.SF_ACT_GELU_TANH}


# =====================================================================================  native engine
class _Engine:
    """Owns the sf_ctx of one *root* module (the full model, or a stand-alone embeddings / encoder / layer /
    pooling head composed inside someone else's model) on one device and keeps its packed weights in sync
    with the nn.Parameters.

    Change detection is cheap on purpose (it runs on every forward, and the streaming / OAD paths are
    host-bound below ~1 ms): the parameter list is collected once and only the version counters are
    compared per call; anything that replaces storages or bypasses the counters goes through
    ``mark_dirty()`` — ``Module._apply`` (``.to()/.cuda()/.half()``), ``load_state_dict``, LoRA insertion
    and the public ``rebind_weights()`` (needed after ``p.data.copy_()/p.data = ...`` style updates,
    which PyTorch does not version)."""

    def __init__(self, root: nn.Module, prefix: str, config, device: torch.device, dtype: torch.dtype):
        cfg = config
        if cfg.attention_type != "divided_space_time":
            raise NotImplementedError("only attention_type='divided_space_time' is on the StreamFormer hot path")
        if cfg.hidden_act not in _ACTS:
            raise NotImplementedError(f"hidden_act={cfg.hidden_act!r} is not supported (gelu / gelu_pytorch_tanh)")
        self.lib = N.load()
        self.device = device
        self.dtype = dtype
        self.prefix = prefix
        c = N.SfConfig()
        img = cfg.image_size[0] if isinstance(cfg.image_size, (tuple, list)) else cfg.image_size
        pat = cfg.patch_size[0] if isinstance(cfg.patch_size, (tuple, list)) else cfg.patch_size
        c.image_size = int(img); c.patch_size = int(pat); c.num_channels = int(cfg.num_channels)
        c.num_frames = int(cfg.num_frames); c.hidden_size = int(cfg.hidden_size)
        c.num_hidden_layers = int(cfg.num_hidden_layers); c.num_attention_heads = int(cfg.num_attention_heads)
        c.intermediate_size = int(cfg.intermediate_size); c.hidden_act = _ACTS[cfg.hidden_act]
        c.layer_norm_eps = float(cfg.layer_norm_eps); c.causal_temporal = int(bool(cfg.enable_causal_temporal))
        c.dtype = sf_dtype(dtype); c.fold_temporal_proj = int(bool(getattr(cfg, "fold_temporal_proj", True)))
        handle = C.c_void_p()
        N.check(self.lib.sf_create(C.byref(c), device.index or 0, C.byref(handle)), "sf_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, self.lib.sf_destroy, handle)
        mean, std = getattr(cfg, "image_mean", None), getattr(cfg, "image_std", None)
        if mean is not None or std is not None:
            self.set_pixel_norm(mean if mean is not None else 0.5, std if std is not None else 0.5)
        self.params: Optional[List[torch.Tensor]] = None
        self.names: List[str] = []
        self.versions: Optional[List[int]] = None
        self.workspace: Optional[torch.Tensor] = None
        self.pos_key = None
        self.binds = 0

    def set_pixel_norm(self, mean, std) -> None:
        mean = [float(mean)] if not isinstance(mean, (tuple, list)) else [float(v) for v in mean]
        std = [float(std)] if not isinstance(std, (tuple, list)) else [float(v) for v in std]
        n = max(len(mean), len(std))
        mean, std = (mean * n)[:n], (std * n)[:n]
        N.check(self.lib.sf_set_pixel_norm(self.handle, (C.c_float * n)(*mean), (C.c_float * n)(*std), n), "sf_set_pixel_norm")

    # -- weights ----------------------------------------------------------------------------------
    def mark_dirty(self) -> None:
        self.params = None
        self.versions = None

    def sync_weights(self, root: nn.Module) -> None:
        params = self.params
        if params is not None:
            versions = [p._version for p in params]
            if versions == self.versions:
                return
        else:
            named = list(root.named_parameters())
            self.names = [self.prefix + n for n, _ in named]
            params = self.params = [p for _, p in named]
            versions = [p._version for p in params]
        descs = (N.SfWeightDesc * len(params))()
        keep = []
        for i, (name, p) in enumerate(zip(self.names, params)):
            t = p.detach()
            if t.device != self.device:
                raise N.NativeError(f"parameter {name} lives on {t.device}, engine on {self.device}")
            if not t.is_contiguous():
                t = t.contiguous()
            keep.append(t)
            bname = name.encode()
            keep.append(bname)
            descs[i].name = bname
            descs[i].data = t.data_ptr()
            descs[i].dtype = sf_dtype(t.dtype)
            descs[i].ndim = min(t.dim(), 4)
            shape = list(t.shape)[:4] if t.dim() <= 4 else [t.numel()]
            if t.dim() > 4:
                descs[i].ndim = 1
            for j, sz in enumerate(shape):
                descs[i].shape[j] = sz
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.sf_bind_weights(self.handle, stream, descs, len(params)), "sf_bind_weights")
        self.versions = versions
        self.pos_key = None
        self.binds += 1

    # -- scratch ----------------------------------------------------------------------------------
    def get_workspace(self, B: int, T: int, H: int, W: int) -> torch.Tensor:
        need = C.c_size_t()
        N.check(self.lib.sf_workspace_bytes(self.handle, B, T, H, W, C.byref(need)), "sf_workspace_bytes")
        if self.workspace is None or self.workspace.numel() < need.value:
            self.workspace = None
            self.workspace = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return self.workspace

    def ensure_pos_table(self, emb: "TimesformerEmbeddingsSigLIP", H: int, W: int) -> None:
        """Non-default resolution: bicubic-antialias resampling of the position table exactly as the
        reference does it (…siglip.py:380-411, a rare path left to PyTorch), handed to the runtime."""
        P = emb.patch_embeddings.patch_size[0]
        S = (H // P) * (W // P)
        pe = emb.position_embeddings
        if S == pe.shape[1] and H == W:
            return
        key = (H, W, pe.data_ptr(), pe._version)
        if key == self.pos_key:
            return
        with torch.no_grad():
            table = emb.interpolate_pos_encoding(None, W, H, npatch=S).float().reshape(S, -1).contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.sf_set_pos_embed(self.handle, stream, table.data_ptr(), S), "sf_set_pos_embed")
        self.pos_key = key


class _NativeRoot:
    """Mixin of every module that can own a native engine: the full model, and each sub-module the
    reference's downstream code instantiates on its own (downstream/AR/models/
    modeling_timesformer_video_classification.py:42-56, models/modeling_timesformer_siglip_adapter.py:
    481-482).  ``_sf_prefix`` maps the module's parameter names onto the reference state-dict names."""
    _sf_prefix = ""

    def _sf_engines(self) -> dict:
        d = self.__dict__.get("_engines")
        if d is None:
            d = {}
            object.__setattr__(self, "_engines", d)
        return d

    def _sf_compute_dtype(self) -> torch.dtype:
        pd = next(self.parameters()).dtype
        if pd in (torch.bfloat16, torch.float16):
            return pd
        if torch.is_autocast_enabled():
            return torch.get_autocast_gpu_dtype()
        return torch.float16 if str(getattr(self.config, "compute_dtype", "bfloat16")) in ("float16", "fp16", "half") \
            else torch.bfloat16

    def _engine(self, device: torch.device) -> _Engine:
        if device.type != "cuda":
            raise N.NativeError("streamformer_b200 runs on CUDA (sm_100a) only; there is no CPU fallback — "
                                "move the model and inputs to a B200")
        dtype = self._sf_compute_dtype()
        key = (device.index if device.index is not None else torch.cuda.current_device(), dtype)
        engines = self._sf_engines()
        eng = engines.get(key)
        if eng is None:
            eng = _Engine(self, self._sf_prefix, self.config, torch.device("cuda", key[0]), dtype)
            engines[key] = eng
        eng.sync_weights(self)
        return eng

    def rebind_weights(self) -> None:
        """Force the packed / folded / LoRA-merged device copies to be rebuilt from the parameters on the
        next forward.  Needed only after updates PyTorch does not version: ``p.data.copy_(...)``,
        ``p.data = ...``, ``m.weight.data.normal_()`` (EMA swaps, the reference adapter's init, some
        DeepSpeed / FSDP utilities).  ``.to()``, ``load_state_dict`` and in-place ops on the parameters
        themselves are detected automatically."""
        for eng in self._sf_engines().values():
            eng.mark_dirty()
        for m in self.modules():
            if m is not self and isinstance(m, _NativeRoot):
                for eng in m._sf_engines().values():
                    eng.mark_dirty()

    def _sf_mark_dirty(self) -> None:
        for eng in self._sf_engines().values():
            eng.mark_dirty()

    def _apply(self, fn, *args, **kwargs):   # .to() / .cuda() / .half(): storages are replaced
        r = super()._apply(fn, *args, **kwargs)
        self._sf_mark_dirty()
        return r

    def _sf_install_hooks(self) -> None:
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._sf_mark_dirty())


def _root_of(module: nn.Module) -> "_NativeRoot":
    """The module whose engine serves ``module``: its owner when it was built inside a larger native
    module (full model, stand-alone encoder), else the module itself."""
    ref = module.__dict__.get("_sf_owner")
    owner = ref() if ref is not None else None
    return owner if owner is not None else module


def _adopt(owner: nn.Module, *children: nn.Module) -> None:
    ref = weakref.ref(owner)
    for m in children:
        object.__setattr__(m, "_sf_owner", ref)


class StreamformerKVCache:
    """Pre-allocated temporal KV cache (replaces transformers.DynamicCache in the KV twin,
    …timesformer_encoder.py:517-518): [layer][K|V][B*N][heads][max_frames][64], appended in place."""

    def __init__(self, engine: _Engine, batch_size: int, num_patches: int, max_frames: int, time_horizon: int = 0):
        self._engine = engine
        self.batch_size, self.num_patches, self.max_frames = batch_size, num_patches, max_frames
        handle = C.c_void_p()
        N.check(engine.lib.sf_kv_create(engine.handle, batch_size, num_patches, max_frames, time_horizon,
                                        C.byref(handle)), "sf_kv_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, engine.lib.sf_kv_destroy, handle)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return int(self._engine.lib.sf_kv_seq_len(self.handle))

    def reset(self) -> None:
        self._engine.lib.sf_kv_reset(self.handle)

    def advance(self, frames: int) -> None:
        """Block-level streaming (``encoder.layer[i](x, T, past_key_value=cache)`` called layer by layer):
        every layer of a step appends at the same position; call this once after the last layer.
        ``model(...)`` and ``encoder(...)`` advance the cache themselves."""
        N.check(self._engine.lib.sf_kv_advance(self.handle, int(frames)), "sf_kv_advance")

    @property
    def graph_launches(self) -> int:
        """Steps served by replaying the captured CUDA graph of a streaming step."""
        return int(self._engine.lib.sf_kv_graph_launches(self.handle))

    def __len__(self) -> int:
        return self.get_seq_length()


@dataclass
class StreamformerOutputWithPast(ModelOutput):
    """BaseModelOutputWithPooling + the cache (the twin returns BaseModelOutputWithPast without the
    pooled output, …timesformer_encoder.py:1387-1392; here both are available)."""
    last_hidden_state: Optional[torch.Tensor] = None
    pooler_output: Optional[torch.Tensor] = None
    past_key_values: Optional[StreamformerKVCache] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _pixel_format(pixel_values: torch.Tensor, num_channels: int) -> Tuple[torch.Tensor, int, int, int]:
    """Validates [B,T,C,H,W] float / uint8 pixels (or interleaved uint8 [B,T,H,W,C] as a decoder hands them
    over) and returns (contiguous tensor, sf pixel dtype, H, W)."""
    if pixel_values.dim() != 5:
        raise ValueError(f"pixel_values must be [B, T, C, H, W], got {tuple(pixel_values.shape)}")
    if pixel_values.dtype == torch.uint8:
        if pixel_values.shape[2] != num_channels and pixel_values.shape[4] == num_channels:
            return pixel_values.contiguous(), N.SF_U8_HWC, pixel_values.shape[2], pixel_values.shape[3]
        return pixel_values.contiguous(), N.SF_U8, pixel_values.shape[3], pixel_values.shape[4]
    if pixel_values.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        pixel_values = pixel_values.float()
    return pixel_values.contiguous(), sf_dtype(pixel_values.dtype), pixel_values.shape[3], pixel_values.shape[4]


def _needs_grad(module: nn.Module, *inputs) -> bool:
    """Autograd is wanted when grad mode is on and either a parameter of the module or an input requires grad."""
    if not torch.is_grad_enabled():
        return False
    return any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs) or any(p.requires_grad for p in module.parameters())


def _no_autograd(what: str, *flags) -> None:
    if any(flags):
        raise NotImplementedError(f"{what}: the differentiable path covers neither the KV cache nor output_attentions; "
                                  "wrap such calls in torch.no_grad()")


# =====================================================================================  parameter holders
class TimesformerPatchEmbeddings(nn.Module):
    """Image to Patch Embedding (reference …siglip.py:300-350)."""

    def __init__(self, config):
        super().__init__()
        image_size = config.image_size if isinstance(config.image_size, (tuple, list)) else (config.image_size,) * 2
        patch_size = config.patch_size if isinstance(config.patch_size, (tuple, list)) else (config.patch_size,) * 2
        self.image_size, self.patch_size = tuple(image_size), tuple(patch_size)
        self.num_patches = (image_size[1] // patch_size[1]) * (image_size[0] // patch_size[0])
        self.projection = nn.Conv2d(config.num_channels, config.hidden_size, kernel_size=patch_size, stride=patch_size)


class TimesformerEmbeddingsSigLIP(_NativeRoot, nn.Module):
    """Patch + position + time embeddings (reference …siglip.py:353-457).  Works inside the full model and
    stand-alone (downstream/AR/models/modeling_timesformer_video_classification.py:48)."""
    _sf_prefix = "embeddings."

    def __init__(self, config):
        super().__init__()
        self.config = config
        self._sf_install_hooks()
        self.attention_type = config.attention_type
        self.patch_embeddings = TimesformerPatchEmbeddings(config)
        self.num_patches = self.patch_embeddings.num_patches
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.num_patches, config.hidden_size))
        self.pos_drop = nn.Dropout(p=config.hidden_dropout_prob)
        if config.attention_type != "space_only":
            self.time_embeddings = nn.Parameter(torch.zeros(1, config.num_frames, config.hidden_size))
            self.time_drop = nn.Dropout(p=config.hidden_dropout_prob)

    def interpolate_pos_encoding(self, x, w, h, npatch: Optional[int] = None):
        """Same resampling as the reference (…siglip.py:380-411); used for non-default resolutions."""
        if npatch is None:
            npatch = x.shape[1]
        Np = self.position_embeddings.shape[1]
        if npatch == Np and w == h:
            return self.position_embeddings
        pos = self.position_embeddings.float()
        dim = pos.shape[-1]
        w0 = w // self.patch_embeddings.patch_size[0]
        h0 = h // self.patch_embeddings.patch_size[1]
        M = int(math.sqrt(Np))
        assert Np == M * M
        pos = nn.functional.interpolate(pos.reshape(1, M, M, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=True,
                                        size=(w0, h0))
        assert (w0, h0) == pos.shape[-2:]
        return pos.permute(0, 2, 3, 1).reshape(1, -1, dim).to(self.position_embeddings.dtype)

    def forward(self, pixel_values, return_size=False, past_key_values=None):
        """pixels [B,T,C,H,W] (float, or uint8 planar / interleaved [B,T,H,W,C]) -> [B, N*T, D] in the
        reference's (b, n, t) token order; ``past_key_values`` offsets the time embedding (KV:336-366)."""
        pixel_values, pix_dtype, H, W = _pixel_format(pixel_values, self.config.num_channels)
        root = _root_of(self)
        eng = root._engine(pixel_values.device)
        B, T = pixel_values.shape[:2]
        eng.ensure_pos_table(self, H, W)
        p = self.patch_embeddings.patch_size
        if _needs_grad(self):
            _no_autograd("TimesformerEmbeddingsSigLIP.forward", past_key_values is not None)
            from .autograd import embed_forward_with_grad
            x = embed_forward_with_grad(root, eng, self, pixel_values, pix_dtype, H, W)
            x = x.to(self.position_embeddings.dtype) if self.position_embeddings.dtype != eng.dtype else x
            return (x, H // p[0], W // p[1]) if return_size else x
        S = (H // p[0]) * (W // p[1])
        x = torch.empty(B, S * T, self.config.hidden_size, dtype=eng.dtype, device=pixel_values.device)
        ws = eng.get_workspace(B, T, H, W)
        off = past_key_values.get_seq_length() if past_key_values is not None else 0
        N.check(eng.lib.sf_embed_forward(eng.handle, _stream(pixel_values), pixel_values.data_ptr(), pix_dtype, B, T, H, W,
                                         off, off + T, x.data_ptr(), ws.data_ptr(), ws.numel()), "sf_embed_forward")
        out_dtype = self.position_embeddings.dtype
        if out_dtype != eng.dtype:
            x = x.to(out_dtype)
        if return_size:
            return x, H // p[0], W // p[1]
        return x


class _LoraMixin:
    def _add_lora_pair(self, base: nn.Linear, a_name: str, b_name: str, rank: int):
        for p in base.parameters():
            p.requires_grad = False
        a = nn.Linear(base.in_features, rank, bias=False)
        b = nn.Linear(rank, base.out_features, bias=False)
        a.to(base.weight.device, base.weight.dtype)
        b.to(base.weight.device, base.weight.dtype)
        if a.weight.device.type != "meta":
            nn.init.normal_(a.weight, std=0.02)
            nn.init.zeros_(b.weight)
        self.add_module(a_name, a)
        self.add_module(b_name, b)


class TimesformerCausalSelfAttention(nn.Module, _LoraMixin):
    """Parameters of the temporal-causal attention (reference …siglip.py:502-615)."""

    def __init__(self, config):
        super().__init__()
        self.num_heads = config.num_attention_heads
        self.scale = (config.hidden_size // config.num_attention_heads) ** -0.5
        self.qkv = nn.Linear(config.hidden_size, config.hidden_size * 3, bias=config.qkv_bias)
        self.attn_drop = nn.Dropout(config.attention_probs_dropout_prob)
        # unused persistent buffer kept for state-dict compatibility (…siglip.py:515-517)
        self.register_buffer("mask", torch.tril(torch.ones(config.num_frames, config.num_frames)))

    def _add_lora(self, lora_rank):
        self._add_lora_pair(self.qkv, "qkv_lora_a", "qkv_lora_b", lora_rank)


class TimesformerSelfAttention(nn.Module, _LoraMixin):
    """Parameters of the spatial attention (reference …siglip.py:618-717)."""

    def __init__(self, config):
        super().__init__()
        self.num_heads = config.num_attention_heads
        self.scale = (config.hidden_size // config.num_attention_heads) ** -0.5
        self.qkv = nn.Linear(config.hidden_size, config.hidden_size * 3, bias=config.qkv_bias)
        self.attn_drop = nn.Dropout(config.attention_probs_dropout_prob)

    def _add_lora(self, lora_rank):
        self._add_lora_pair(self.qkv, "qkv_lora_a", "qkv_lora_b", lora_rank)


class TimesformerSelfOutput(nn.Module, _LoraMixin):
    """Attention output projection (reference …siglip.py:720-763)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def _add_lora(self, lora_rank: int = 32):
        self._add_lora_pair(self.dense, "dense_lora_a", "dense_lora_b", lora_rank)


class TimeSformerCausalAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = TimesformerCausalSelfAttention(config)
        self.output = TimesformerSelfOutput(config)


class TimeSformerAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = TimesformerSelfAttention(config)
        self.output = TimesformerSelfOutput(config)


class TimesformerIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class TimesformerOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class TimesformerLayerSigLIP(_NativeRoot, nn.Module):
    """One divided space-time block (reference …siglip.py:840-1004).  ``forward(x[B,N*T,D], T)`` runs
    the fused native layer; drop-path is identity at the shipped rate 0 and in eval."""

    def __init__(self, config, layer_index: int):
        super().__init__()
        self._sf_prefix = f"encoder.layer.{layer_index}."
        self._sf_install_hooks()
        if config.attention_type not in ["divided_space_time", "space_only", "joint_space_time"]:
            raise ValueError("Unknown attention type: {}".format(config.attention_type))
        # stochastic-depth rule of the reference without its .item() (meta-device safe)
        L = max(config.num_hidden_layers - 1, 1)
        self.drop_path_rate = float(config.drop_path_rate) * layer_index / L
        self.drop_path = nn.Identity()
        self.attention = TimeSformerAttention(config)
        self.intermediate = TimesformerIntermediate(config)
        self.output = TimesformerOutput(config)
        self.layernorm_before = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.layernorm_after = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.config = config
        self.attention_type = config.attention_type
        self.layer_index = layer_index
        if self.attention_type == "divided_space_time":
            self.temporal_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
            if config.enable_causal_temporal:
                self.temporal_attention = TimeSformerCausalAttention(config)
            else:
                self.temporal_attention = TimeSformerAttention(config)
            self.temporal_dense = nn.Linear(config.hidden_size, config.hidden_size)
            self.temporal_attention_gating = nn.Parameter(torch.tensor(0.0))

    def forward(self, hidden_states: torch.Tensor, num_frames: int, output_attentions: bool = False,
                past_key_value: Optional[StreamformerKVCache] = None):
        """(…siglip.py:900-1004).  With ``past_key_value`` the temporal K/V of the new frames are appended at
        the cache's current position; the caller advances the cache once per step (``cache.advance(T)``)
        after the LAST layer — all layers of a step share the same position."""
        root = _root_of(self)
        eng = root._engine(hidden_states.device)
        B, NT, D = hidden_states.shape
        if NT % num_frames:
            raise ValueError(f"sequence length {NT} is not a multiple of num_frames={num_frames}")
        S = NT // num_frames
        in_dtype = hidden_states.dtype
        if _needs_grad(self, hidden_states):
            _no_autograd("TimesformerLayerSigLIP.forward", past_key_value is not None, output_attentions)
            from .autograd import stack_forward_with_grad
            out = stack_forward_with_grad(root, eng, [self.layer_index], hidden_states, num_frames)[-1]
            return (out.to(in_dtype) if in_dtype != eng.dtype and in_dtype.is_floating_point else out,)
        x = hidden_states.to(eng.dtype).contiguous()
        out = torch.empty_like(x)
        heads = self.config.num_attention_heads
        probs = torch.empty(B * num_frames, heads, S, S, dtype=torch.float32, device=x.device) if output_attentions else None
        P = eng_patch(self.config)
        ws = eng.get_workspace(B, num_frames, P, P * S)  # any geometry with S patches per frame
        if past_key_value is not None and past_key_value._engine is not eng:
            raise ValueError("past_key_value was created by another module's engine; allocate it with new_kv_cache() of the "
                             "module that runs the layers")
        N.check(eng.lib.sf_layer_forward(eng.handle, _stream(x), self.layer_index, x.data_ptr(), out.data_ptr(), B,
                                         num_frames, S, past_key_value.handle if past_key_value is not None else None,
                                         probs.data_ptr() if probs is not None else None, ws.data_ptr(), ws.numel()),
                "sf_layer_forward")
        out = out.to(in_dtype) if in_dtype != eng.dtype and in_dtype.is_floating_point else out
        return (out, probs.to(out.dtype)) if output_attentions else (out,)


def eng_patch(config) -> int:
    return config.patch_size[0] if isinstance(config.patch_size, (tuple, list)) else config.patch_size


class TimesformerEncoder(_NativeRoot, nn.Module):
    """The layer stack (reference …siglip.py:1007-1063) with the reference's forward signature, usable inside
    the full model and stand-alone (downstream/AR/…video_classification.py:49, 116-122; the OVIS ViT-adapter,
    models/modeling_timesformer_siglip_adapter.py:481-482, 629-640)."""
    _sf_prefix = "encoder."

    def __init__(self, config):
        super().__init__()
        self.config = config
        self._sf_install_hooks()
        self.layer = nn.ModuleList([TimesformerLayerSigLIP(config, i) for i in range(config.num_hidden_layers)])
        _adopt(self, *self.layer)
        self.gradient_checkpointing = False

    def new_kv_cache(self, batch_size: int, num_patches: int, max_frames: int, time_horizon: int = 0,
                     device: Optional[torch.device] = None) -> StreamformerKVCache:
        device = torch.device(device) if device is not None else next(self.parameters()).device
        return StreamformerKVCache(_root_of(self)._engine(device), batch_size, num_patches, max_frames, time_horizon)

    def forward(self, hidden_states: torch.Tensor, num_frames: int, output_attentions: bool = False,
                output_hidden_states: bool = False, return_dict: bool = True,
                past_key_values: Optional[StreamformerKVCache] = None):
        root = _root_of(self)
        eng = root._engine(hidden_states.device)
        B, NT, D = hidden_states.shape
        if NT % num_frames:
            raise ValueError(f"sequence length {NT} is not a multiple of num_frames={num_frames}")
        S, L = NT // num_frames, len(self.layer)
        in_dtype = hidden_states.dtype
        if _needs_grad(self, hidden_states):
            _no_autograd("TimesformerEncoder.forward", past_key_values is not None, output_attentions)
            from .autograd import stack_forward_with_grad
            outs = stack_forward_with_grad(root, eng, list(range(L)), hidden_states, num_frames)
            cast = (lambda t: t.to(in_dtype)) if in_dtype != eng.dtype and in_dtype.is_floating_point else (lambda t: t)
            hs_t = tuple([hidden_states] + [cast(o) for o in outs]) if output_hidden_states else None
            if not return_dict:
                return tuple(v for v in [cast(outs[-1]), hs_t] if v is not None)
            return BaseModelOutput(last_hidden_state=cast(outs[-1]), hidden_states=hs_t, attentions=None)
        x = hidden_states.to(eng.dtype).contiguous()
        heads = self.config.num_attention_heads
        hs = [x] + [torch.empty_like(x) for _ in range(L)] if output_hidden_states else None
        out = torch.empty_like(x) if hs is None else hs[-1]
        atts = [torch.empty(B * num_frames, heads, S, S, dtype=torch.float32, device=x.device) for _ in range(L)] \
            if output_attentions else None
        P = eng_patch(self.config)
        ws = eng.get_workspace(B, num_frames, P, P * S)
        if past_key_values is not None and past_key_values._engine is not eng:
            raise ValueError("past_key_values was created by another module's engine")
        N.check(eng.lib.sf_encoder_forward(eng.handle, _stream(x), x.data_ptr(), B, num_frames, S,
                                           past_key_values.handle if past_key_values is not None else None, out.data_ptr(),
                                           N.ptr_array([t.data_ptr() for t in hs]) if hs is not None else None,
                                           N.ptr_array([t.data_ptr() for t in atts]) if atts is not None else None,
                                           ws.data_ptr(), ws.numel()), "sf_encoder_forward")
        if in_dtype != eng.dtype and in_dtype.is_floating_point:
            out = out.to(in_dtype)
            hs = [h.to(in_dtype) for h in hs] if hs is not None else None
        hs_t = tuple(hs) if hs is not None else None
        at_t = tuple(a.to(out.dtype) for a in atts) if atts is not None else None
        if not return_dict:
            return tuple(v for v in [out, hs_t, at_t] if v is not None)
        return BaseModelOutput(last_hidden_state=out, hidden_states=hs_t, attentions=at_t)


class SiglipMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class TimesformerSiglipMultiheadAttentionPoolingHead(_NativeRoot, nn.Module):
    """Multihead attention pooling with a learned probe (reference …siglip.py:1128-1154)."""
    _sf_prefix = "head."

    def __init__(self, config):
        super().__init__()
        self.config = config
        self._sf_install_hooks()
        self.probe = nn.Parameter(torch.randn(1, 1, config.hidden_size))
        self.attention = torch.nn.MultiheadAttention(config.hidden_size, config.num_attention_heads, batch_first=True)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.mlp = SiglipMLP(config)

    def forward(self, hidden_state):  # (B*T, N, D) -> (B*T, D)
        root = _root_of(self)
        eng = root._engine(hidden_state.device)
        frames, S, D = hidden_state.shape
        in_dtype = hidden_state.dtype
        if _needs_grad(self, hidden_state):
            from .autograd import head_forward_with_grad
            out = head_forward_with_grad(root, eng, hidden_state)
            return out.to(in_dtype) if in_dtype != eng.dtype and in_dtype.is_floating_point else out
        x = hidden_state.to(eng.dtype).contiguous()
        out = torch.empty(frames, D, dtype=eng.dtype, device=x.device)
        P = eng_patch(self.config)
        ws = eng.get_workspace(frames, 1, P, P * S)
        N.check(eng.lib.sf_head_forward(eng.handle, _stream(x), x.data_ptr(), frames, S, out.data_ptr(), ws.data_ptr(),
                                        ws.numel()), "sf_head_forward")
        return out.to(in_dtype) if in_dtype != eng.dtype and in_dtype.is_floating_point else out


# =====================================================================================  models
class TimesformerPreTrainedModel(PreTrainedModel):
    """Base class (reference …siglip.py:1066-1109); downstream models subclass it and compose the native
    sub-modules above (downstream/AR/models/modeling_timesformer_video_classification.py:42)."""
    config_class = StreamformerConfig
    base_model_prefix = "timesformer"
    main_input_name = "pixel_values"
    supports_gradient_checkpointing = True
    _no_split_modules = ["TimesformerLayerSigLIP"]

    def _init_weights(self, module):
        """Reference initialisation (…siglip.py:1077-1109)."""
        std = self.config.initializer_range
        if isinstance(module, (nn.Linear, nn.Conv2d)):
            nn.init.trunc_normal_(module.weight, std=std)
            if module.bias is not None:
                nn.init.constant_(module.bias, 0)
        elif isinstance(module, nn.LayerNorm):
            nn.init.constant_(module.bias, 0)
            nn.init.constant_(module.weight, 1.0)
        elif isinstance(module, TimesformerEmbeddingsSigLIP):
            nn.init.trunc_normal_(module.position_embeddings, std=std)
            module.patch_embeddings.apply(self._init_weights)


class TimesformerMultiTaskingModelSigLIP(_NativeRoot, TimesformerPreTrainedModel):
    """Drop-in for the reference class of the same name (…siglip.py:1241-1354) and for its KV-cache
    twin (…timesformer_encoder.py:1255-1392): ``past_key_values / use_cache`` select the streaming path."""
    _sf_prefix = ""

    def __init__(self, config: StreamformerConfig):
        super().__init__(config)
        self.config = config
        self._sf_install_hooks()
        self.embeddings = TimesformerEmbeddingsSigLIP(config)
        self.encoder = TimesformerEncoder(config)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.head = TimesformerSiglipMultiheadAttentionPoolingHead(config)
        if config.add_lora_spatial:
            self.add_lora_spatial()
        # sub-modules called on their own (model.embeddings(px), model.encoder.layer[i](x, T), model.head(x))
        # are served by THIS module's engine: one set of packed weights per replica
        _adopt(self, self.embeddings, self.encoder, self.head, *self.encoder.layer)
        self.post_init()

    # ---- reference utility surface -------------------------------------------------------------
    def get_input_embeddings(self):
        return self.embeddings.patch_embeddings

    def _prune_heads(self, heads_to_prune):
        raise NotImplementedError("head pruning is not on the StreamFormer hot path")

    def add_lora_spatial(self):
        """Rank-32 LoRA on every spatial qkv and out-proj (reference …siglip.py:1271-1282)."""
        assert self.encoder.layer[0].attention_type == "divided_space_time"
        for layer in self.encoder.layer:
            if not hasattr(layer.attention.attention, "qkv_lora_a"):
                layer.attention.attention._add_lora(32)
                layer.attention.output._add_lora(32)
        self.rebind_weights()

    def frozen_spatial(self):
        for layer in self.encoder.layer:
            for p in layer.attention.attention.qkv.parameters():
                p.requires_grad = False
            for p in layer.attention.output.dense.parameters():
                p.requires_grad = False

    def set_pixel_normalization(self, mean=0.5, std=0.5) -> None:
        """uint8 inputs are normalised on the GPU as (x / 255 - mean) / std — the loaders'
        ClipToTensor + Normalize (extract_oad_feature.py:42-48; SigLIP: 0.5 / 0.5, the default)."""
        self.config.image_mean, self.config.image_std = mean, std
        for eng in self._sf_engines().values():
            eng.set_pixel_norm(mean, std)

    def new_kv_cache(self, batch_size: int, max_frames: Optional[int] = None, image_size: Optional[Tuple[int, int]] = None,
                     time_horizon: int = 0, device: Optional[torch.device] = None) -> StreamformerKVCache:
        """Allocate a streaming cache (≙ DynamicCache()).  ``time_horizon`` > 0 fixes the number of
        frames the time-embedding table is stretched over, making streaming identical to a one-shot
        forward of that many frames even beyond ``config.num_frames``."""
        device = device or self.embeddings.position_embeddings.device
        eng = self._engine(torch.device(device))
        P = self.config.patch_size
        H, W = image_size or (self.config.image_size, self.config.image_size)
        return StreamformerKVCache(eng, batch_size, (H // P) * (W // P), max_frames or self.config.kv_cache_max_frames,
                                   time_horizon)

    # ---- the hot path --------------------------------------------------------------------------
    def forward(
        self,
        pixel_values: torch.Tensor,  # (B, T, 3, H, W) float, or uint8 (B, T, 3, H, W) / (B, T, H, W, 3)
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        past_key_values: Optional[StreamformerKVCache] = None,
        use_cache: bool = False,
        cache_position: Optional[torch.LongTensor] = None,
    ) -> Union[Tuple[torch.Tensor, ...], BaseModelOutputWithPooling, StreamformerOutputWithPast]:
        cfg = self.config
        output_attentions = output_attentions if output_attentions is not None else cfg.output_attentions
        output_hidden_states = output_hidden_states if output_hidden_states is not None else cfg.output_hidden_states
        return_dict = return_dict if return_dict is not None else getattr(cfg, "return_dict", True)

        pixel_values, pix_dtype, H, W = _pixel_format(pixel_values, cfg.num_channels)
        eng = self._engine(pixel_values.device)
        B, T = pixel_values.shape[:2]
        P = cfg.patch_size
        S = (H // P) * (W // P)
        D = cfg.hidden_size
        dev = pixel_values.device
        eng.ensure_pos_table(self.embeddings, H, W)

        streaming = use_cache or past_key_values is not None
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            if streaming or output_attentions:
                raise NotImplementedError("the differentiable path covers the one-shot forward (no KV cache, no "
                                          "output_attentions); wrap streaming / attention-map calls in torch.no_grad()")
            from .autograd import encoder_forward_with_grad
            return encoder_forward_with_grad(self, eng, pixel_values, pix_dtype, H, W, output_hidden_states, return_dict)
        if streaming:
            if past_key_values is None:
                past_key_values = self.new_kv_cache(B, image_size=(H, W), device=dev)
            if not isinstance(past_key_values, StreamformerKVCache):
                raise TypeError("past_key_values must be a StreamformerKVCache (model.new_kv_cache(...)); "
                                "transformers.DynamicCache objects are not supported by the native runtime")
            if past_key_values._engine is not eng:
                raise ValueError("past_key_values belongs to another engine (device / dtype changed since new_kv_cache())")
            if cache_position is not None and int(cache_position[0]) != past_key_values.get_seq_length():
                raise ValueError("cache_position must continue the cache (start at past_key_values.get_seq_length())")
            if output_attentions:
                raise NotImplementedError("output_attentions is not available on the streaming path")

        last_hidden = torch.empty(B, T, S, D, dtype=eng.dtype, device=dev)
        pooled = torch.empty(B, T, D, dtype=eng.dtype, device=dev)
        hs: Optional[List[torch.Tensor]] = None
        atts: Optional[List[torch.Tensor]] = None
        if output_hidden_states:
            hs = [torch.empty(B, S * T, D, dtype=eng.dtype, device=dev) for _ in range(cfg.num_hidden_layers + 1)]
        if output_attentions:
            atts = [torch.empty(B * T, cfg.num_attention_heads, S, S, dtype=torch.float32, device=dev)
                    for _ in range(cfg.num_hidden_layers)]
        ws = eng.get_workspace(B, T, H, W)
        stream = torch.cuda.current_stream(dev).cuda_stream
        hs_ptrs = N.ptr_array([t.data_ptr() for t in hs]) if hs is not None else None
        if streaming:
            N.check(eng.lib.sf_forward_stream(eng.handle, stream, past_key_values.handle, pixel_values.data_ptr(),
                                              pix_dtype, B, T, H, W, last_hidden.data_ptr(),
                                              pooled.data_ptr(), hs_ptrs, ws.data_ptr(), ws.numel()),
                    "sf_forward_stream")
        else:
            at_ptrs = N.ptr_array([t.data_ptr() for t in atts]) if atts is not None else None
            N.check(eng.lib.sf_forward(eng.handle, stream, pixel_values.data_ptr(), pix_dtype, B, T,
                                       H, W, last_hidden.data_ptr(), pooled.data_ptr(), hs_ptrs, at_ptrs, ws.data_ptr(),
                                       ws.numel()), "sf_forward")

        out_dtype = self.embeddings.position_embeddings.dtype  # output dtype = parameter dtype
        if out_dtype != eng.dtype:
            last_hidden, pooled = last_hidden.to(out_dtype), pooled.to(out_dtype)
            hs = [h.to(out_dtype) for h in hs] if hs is not None else None
        hs_t = tuple(hs) if hs is not None else None
        at_t = tuple(a.to(out_dtype) for a in atts) if atts is not None else None
        if not return_dict:
            return tuple(v for v in [last_hidden, hs_t, at_t] if v is not None)
        if streaming:
            return StreamformerOutputWithPast(last_hidden_state=last_hidden, pooler_output=pooled,
                                              past_key_values=past_key_values, hidden_states=hs_t, attentions=at_t)
        return BaseModelOutputWithPooling(last_hidden_state=last_hidden, pooler_output=pooled, hidden_states=hs_t,
                                          attentions=at_t)
