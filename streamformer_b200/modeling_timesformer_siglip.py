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
from transformers.modeling_outputs import BaseModelOutputWithPooling, ModelOutput
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
    """Owns the sf_ctx of one model replica on one device and keeps its packed weights in sync with
    the nn.Parameters (re-binds when any parameter's version counter or storage changes)."""

    def __init__(self, model: "TimesformerMultiTaskingModelSigLIP", device: torch.device, dtype: torch.dtype):
        cfg = model.config
        if cfg.attention_type != "divided_space_time":
            raise NotImplementedError("only attention_type='divided_space_time' is on the StreamFormer hot path")
        if cfg.hidden_act not in _ACTS:
            raise NotImplementedError(f"hidden_act={cfg.hidden_act!r} is not supported (gelu / gelu_pytorch_tanh)")
        self.lib = N.load()
        self.device = device
        self.dtype = dtype
        c = N.SfConfig()
        c.image_size = int(cfg.image_size); c.patch_size = int(cfg.patch_size); c.num_channels = int(cfg.num_channels)
        c.num_frames = int(cfg.num_frames); c.hidden_size = int(cfg.hidden_size)
        c.num_hidden_layers = int(cfg.num_hidden_layers); c.num_attention_heads = int(cfg.num_attention_heads)
        c.intermediate_size = int(cfg.intermediate_size); c.hidden_act = _ACTS[cfg.hidden_act]
        c.layer_norm_eps = float(cfg.layer_norm_eps); c.causal_temporal = int(bool(cfg.enable_causal_temporal))
        c.dtype = sf_dtype(dtype); c.fold_temporal_proj = int(bool(getattr(cfg, "fold_temporal_proj", True)))
        handle = C.c_void_p()
        N.check(self.lib.sf_create(C.byref(c), device.index or 0, C.byref(handle)), "sf_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, self.lib.sf_destroy, handle)
        self.bound_key = None
        self.workspace: Optional[torch.Tensor] = None
        self.pos_key = None

    # -- weights ----------------------------------------------------------------------------------
    @staticmethod
    def _key(tensors) -> Tuple:
        return tuple((t.data_ptr(), t._version) for _, t in tensors)

    def sync_weights(self, model: nn.Module) -> None:
        tensors = [(n, p) for n, p in model.named_parameters()]
        key = self._key(tensors)
        if key == self.bound_key:
            return
        descs = (N.SfWeightDesc * len(tensors))()
        keep = []
        for i, (name, p) in enumerate(tensors):
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
            for j, s in enumerate(shape):
                descs[i].shape[j] = s
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.sf_bind_weights(self.handle, stream, descs, len(tensors)), "sf_bind_weights")
        self.bound_key = key
        self.pos_key = None

    # -- scratch ----------------------------------------------------------------------------------
    def get_workspace(self, B: int, T: int, H: int, W: int) -> torch.Tensor:
        need = C.c_size_t()
        N.check(self.lib.sf_workspace_bytes(self.handle, B, T, H, W, C.byref(need)), "sf_workspace_bytes")
        if self.workspace is None or self.workspace.numel() < need.value:
            self.workspace = None
            self.workspace = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return self.workspace

    def ensure_pos_table(self, model, H: int, W: int) -> None:
        """Non-default resolution: bicubic-antialias resampling of the position table exactly as the
        reference does it (…siglip.py:380-411, a rare path left to PyTorch), handed to the runtime."""
        cfg = model.config
        P = cfg.patch_size
        S = (H // P) * (W // P)
        pe = model.embeddings.position_embeddings
        if S == pe.shape[1] and H == W:
            return
        key = (H, W, pe.data_ptr(), pe._version)
        if key == self.pos_key:
            return
        table = model.embeddings.interpolate_pos_encoding(None, W, H, npatch=S).float().reshape(S, -1).contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.sf_set_pos_embed(self.handle, stream, table.data_ptr(), S), "sf_set_pos_embed")
        self.pos_key = key


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


def _owner(module: nn.Module) -> "TimesformerMultiTaskingModelSigLIP":
    ref = getattr(module, "_sf_owner", None)
    owner = ref() if ref is not None else None
    if owner is None:
        raise NotImplementedError(
            f"{type(module).__name__} was built stand-alone; in this round sub-modules run only as part of "
            "TimesformerMultiTaskingModelSigLIP (which owns the native engine)")
    return owner


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


class TimesformerEmbeddingsSigLIP(nn.Module):
    """Patch + position + time embeddings (reference …siglip.py:353-457)."""

    def __init__(self, config):
        super().__init__()
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
        owner = _owner(self)
        x = owner._embed(pixel_values, past_key_values)
        if return_size:
            p = self.patch_embeddings.patch_size
            return x, pixel_values.shape[-2] // p[0], pixel_values.shape[-1] // p[1]
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


class TimesformerLayerSigLIP(nn.Module):
    """One divided space-time block (reference …siglip.py:840-1004).  ``forward(x[B,N*T,D], T)`` runs
    the fused native layer; drop-path is identity at the shipped rate 0 and in eval."""

    def __init__(self, config, layer_index: int):
        super().__init__()
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
        owner = _owner(self)
        return owner._layer(self.layer_index, hidden_states, num_frames, output_attentions, past_key_value)


class TimesformerEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([TimesformerLayerSigLIP(config, i) for i in range(config.num_hidden_layers)])
        self.gradient_checkpointing = False


class SiglipMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class TimesformerSiglipMultiheadAttentionPoolingHead(nn.Module):
    """Multihead attention pooling with a learned probe (reference …siglip.py:1128-1154)."""

    def __init__(self, config):
        super().__init__()
        self.probe = nn.Parameter(torch.randn(1, 1, config.hidden_size))
        self.attention = torch.nn.MultiheadAttention(config.hidden_size, config.num_attention_heads, batch_first=True)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.mlp = SiglipMLP(config)

    def forward(self, hidden_state):  # (B*T, N, D) -> (B*T, D)
        return _owner(self)._head(hidden_state)


# =====================================================================================  models
class TimesformerPreTrainedModel(PreTrainedModel):
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


class TimesformerMultiTaskingModelSigLIP(TimesformerPreTrainedModel):
    """Drop-in for the reference class of the same name (…siglip.py:1241-1354) and for its KV-cache
    twin (…timesformer_encoder.py:1255-1392): ``past_key_values / use_cache`` select the streaming path."""

    def __init__(self, config: StreamformerConfig):
        super().__init__(config)
        self.config = config
        self.embeddings = TimesformerEmbeddingsSigLIP(config)
        self.encoder = TimesformerEncoder(config)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.head = TimesformerSiglipMultiheadAttentionPoolingHead(config)
        if config.add_lora_spatial:
            self.add_lora_spatial()
        self._engines = {}
        ref = weakref.ref(self)
        for m in [self.embeddings, self.head, *self.encoder.layer]:
            object.__setattr__(m, "_sf_owner", ref)
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

    def frozen_spatial(self):
        for layer in self.encoder.layer:
            for p in layer.attention.attention.qkv.parameters():
                p.requires_grad = False
            for p in layer.attention.output.dense.parameters():
                p.requires_grad = False

    # ---- engine plumbing -----------------------------------------------------------------------
    def _compute_dtype(self) -> torch.dtype:
        pd = self.embeddings.position_embeddings.dtype
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
        dtype = self._compute_dtype()
        key = (device.index or 0, dtype)
        eng = self._engines.get(key)
        if eng is None:
            eng = _Engine(self, torch.device("cuda", device.index or 0), dtype)
            self._engines[key] = eng
        eng.sync_weights(self)
        return eng

    def _check_inputs(self, pixel_values: torch.Tensor) -> torch.Tensor:
        if pixel_values.dim() != 5:
            raise ValueError(f"pixel_values must be [B, T, C, H, W], got {tuple(pixel_values.shape)}")
        if pixel_values.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            pixel_values = pixel_values.float()
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "streamformer_b200 round 1 implements the forward pass only (SURVEY §8f rank 1: backward is next); "
                "call model.eval() or wrap the forward in torch.no_grad()")
        return pixel_values.contiguous()

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

    # ---- block-level API -----------------------------------------------------------------------
    def _embed(self, pixel_values: torch.Tensor, past_key_values: Optional[StreamformerKVCache] = None) -> torch.Tensor:
        pixel_values = self._check_inputs(pixel_values)
        eng = self._engine(pixel_values.device)
        B, T, _, H, W = pixel_values.shape
        eng.ensure_pos_table(self, H, W)
        S = (H // self.config.patch_size) * (W // self.config.patch_size)
        x = torch.empty(B, S * T, self.config.hidden_size, dtype=eng.dtype, device=pixel_values.device)
        ws = eng.get_workspace(B, T, H, W)
        off = past_key_values.get_seq_length() if past_key_values is not None else 0
        N.check(eng.lib.sf_embed_forward(eng.handle, torch.cuda.current_stream().cuda_stream, pixel_values.data_ptr(),
                                         sf_dtype(pixel_values.dtype), B, T, H, W, off, off + T, x.data_ptr(),
                                         ws.data_ptr(), ws.numel()), "sf_embed_forward")
        return x

    def _layer(self, index: int, hidden_states: torch.Tensor, num_frames: int, output_attentions: bool = False,
               past_key_value: Optional[StreamformerKVCache] = None):
        eng = self._engine(hidden_states.device)
        B, NT, D = hidden_states.shape
        S = NT // num_frames
        x = hidden_states.to(eng.dtype).contiguous()
        out = torch.empty_like(x)
        probs = torch.empty(B * num_frames, self.config.num_attention_heads, S, S, dtype=torch.float32,
                            device=x.device) if output_attentions else None
        P = self.config.patch_size
        ws = eng.get_workspace(B, num_frames, P, P * S)  # any geometry with S patches per frame
        N.check(eng.lib.sf_layer_forward(eng.handle, torch.cuda.current_stream().cuda_stream, index, x.data_ptr(),
                                         out.data_ptr(), B, num_frames, S,
                                         past_key_value.handle if past_key_value is not None else None,
                                         probs.data_ptr() if probs is not None else None, ws.data_ptr(), ws.numel()),
                "sf_layer_forward")
        return (out, probs.to(eng.dtype)) if output_attentions else (out,)

    def _head(self, hidden_state: torch.Tensor) -> torch.Tensor:
        eng = self._engine(hidden_state.device)
        frames, S, D = hidden_state.shape
        x = hidden_state.to(eng.dtype).contiguous()
        out = torch.empty(frames, D, dtype=eng.dtype, device=x.device)
        P = self.config.patch_size
        ws = eng.get_workspace(frames, 1, P, P * S)
        N.check(eng.lib.sf_head_forward(eng.handle, torch.cuda.current_stream().cuda_stream, x.data_ptr(), frames, S,
                                        out.data_ptr(), ws.data_ptr(), ws.numel()), "sf_head_forward")
        return out

    # ---- the hot path --------------------------------------------------------------------------
    def forward(
        self,
        pixel_values: torch.Tensor,  # (B, T, 3, H, W)
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

        pixel_values = self._check_inputs(pixel_values)
        eng = self._engine(pixel_values.device)
        B, T, _, H, W = pixel_values.shape
        P = cfg.patch_size
        S = (H // P) * (W // P)
        D = cfg.hidden_size
        dev = pixel_values.device
        eng.ensure_pos_table(self, H, W)

        streaming = use_cache or past_key_values is not None
        if streaming:
            if past_key_values is None:
                past_key_values = self.new_kv_cache(B, image_size=(H, W), device=dev)
            if not isinstance(past_key_values, StreamformerKVCache):
                raise TypeError("past_key_values must be a StreamformerKVCache (model.new_kv_cache(...)); "
                                "transformers.DynamicCache objects are not supported by the native runtime")
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
                                              sf_dtype(pixel_values.dtype), B, T, H, W, last_hidden.data_ptr(),
                                              pooled.data_ptr(), hs_ptrs, ws.data_ptr(), ws.numel()),
                    "sf_forward_stream")
        else:
            at_ptrs = N.ptr_array([t.data_ptr() for t in atts]) if atts is not None else None
            N.check(eng.lib.sf_forward(eng.handle, stream, pixel_values.data_ptr(), sf_dtype(pixel_values.dtype), B, T,
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
