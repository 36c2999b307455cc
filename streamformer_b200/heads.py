"""Task heads that consume the encoder's (all-gathered) ``pooler_output`` — the step right after the hot path
(SURVEY.md §8 f2).  Mirrors the reference's zero-shot classification head and SigLIP loss
(models/modeling_timesformer_siglip.py:1640-1726, 193-297, 2324-2351): same class names, arguments and
return values; the arithmetic (L2 norm, logits GEMM, log-sigmoid loss, and their gradients) runs in ONE
sm_100a launch behind ``sf_op_siglip_head`` — no eager fallback.

The text side (SigLIP text tower, tokenizer) is outside §8: label / caption embeddings are handed in as tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import _native as N
from . import ops
from .ops import sf_dtype

__all__ = ["siglip_head", "TimesformerVideoClassificationHead", "SigLipLoss", "gathered_classification_loss"]


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _rows(t: torch.Tensor) -> torch.Tensor:
    """[rows, D] view with unit inner stride and 16-byte aligned rows (a strided slice such as pooler_output[:, -1]
    is used in place when it qualifies)."""
    if t.dim() != 2:
        raise ValueError(f"expected a [rows, D] tensor, got {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


class _SiglipHeadFn(torch.autograd.Function):
    """loss, logits = head(image, text): the forward launch also leaves d loss / d logits behind; the backward is
    one GEMM (dlogits . text^) and the normalisation's Jacobian, both native."""

    @staticmethod
    def forward(ctx, image, text, logit_scale, logit_bias, targets, diag_offset, normalize_image, normalize_text, loss_div,
                want_logits):
        if not image.is_cuda:
            raise N.NativeError("streamformer_b200 heads run on CUDA (sm_100a) only; there is no CPU fallback")
        lib = N.load()
        dt = image.dtype
        image, text = _rows(image), _rows(text.to(dt))
        B, D = image.shape
        L = text.shape[0]
        dev = image.device
        scale = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        bias = logit_bias.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous() if logit_bias is not None else None
        logits = torch.empty(B, L, dtype=torch.float32, device=dev) if want_logits else None
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        needs_grad = any(ctx.needs_input_grad[:4])
        Lp = (L + 7) // 8 * 8
        dlogits = torch.zeros(B, Lp, dtype=dt, device=dev) if needs_grad else None
        dparams = torch.zeros(2, dtype=torch.float32, device=dev) if needs_grad else None
        tg = targets.to(device=dev, dtype=torch.int64).contiguous() if targets is not None else None
        N.check(lib.sf_op_siglip_head(_stream(image), sf_dtype(dt), image.data_ptr(), image.stride(0), text.data_ptr(),
                                      text.stride(0), B, L, D, scale.data_ptr(), bias.data_ptr() if bias is not None else None,
                                      int(normalize_image), int(normalize_text), tg.data_ptr() if tg is not None else None,
                                      int(diag_offset), float(loss_div), logits.data_ptr() if logits is not None else None, L,
                                      loss.data_ptr(), dlogits.data_ptr() if dlogits is not None else None, Lp,
                                      dparams.data_ptr() if dparams is not None else None), "sf_op_siglip_head")
        ctx.save_for_backward(image, text, scale, dlogits, dparams)
        ctx.cfg = (bool(normalize_image), bool(normalize_text), logit_bias is not None)
        ctx.mark_non_differentiable(*([logits] if logits is not None else []))
        return loss.reshape(()), logits

    @staticmethod
    def backward(ctx, g_loss, _g_logits):
        image, text, scale, dlogits, dparams = ctx.saved_tensors
        norm_image, norm_text, has_bias = ctx.cfg
        if norm_text and ctx.needs_input_grad[1]:
            raise NotImplementedError("gradients w.r.t. normalised text features are outside SURVEY §8 (the text tower is not on the path)")
        lib = N.load()
        dt = image.dtype
        B, D = image.shape
        L, Lp = text.shape[0], dlogits.shape[1]
        g_image = None
        if ctx.needs_input_grad[0]:
            # d image^ = exp(scale) * dlogits[B, L] . text^[L, D]  ->  gemm(a = dlogits, w = text^T [D, L])
            tn = text.float()
            if norm_text:
                tn = tn / tn.norm(dim=-1, keepdim=True)            # data preparation of a constant operand
            wt = torch.zeros(D, Lp, dtype=dt, device=image.device)
            wt[:, :L] = tn.t().to(dt)
            dxhat = ops.gemm(dlogits, wt)
            if norm_image:
                g_image = torch.empty(B, D, dtype=dt, device=image.device)
                N.check(lib.sf_op_l2norm_backward(_stream(image), sf_dtype(dt), image.data_ptr(), image.stride(0), dxhat.data_ptr(),
                                                  dxhat.stride(0), scale.data_ptr(), g_image.data_ptr(), D, B, D), "sf_op_l2norm_backward")
            else:
                g_image = (dxhat.float() * scale.exp()).to(dt)
            g_image = g_image * g_loss.to(dt)
        g_scale = (dparams[0] * g_loss.to(dparams.device)).reshape(()) if ctx.needs_input_grad[2] else None
        g_bias = (dparams[1] * g_loss.to(dparams.device)).reshape(()) if (has_bias and ctx.needs_input_grad[3]) else None
        return g_image, None, g_scale, g_bias, None, None, None, None, None, None


def siglip_head(image: torch.Tensor, text: torch.Tensor, logit_scale: torch.Tensor, logit_bias: Optional[torch.Tensor] = None,
                targets: Optional[torch.Tensor] = None, diag_offset: int = 0, normalize_image: bool = True,
                normalize_text: bool = False, loss_div: Optional[float] = None, want_logits: bool = True
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """``loss, logits_per_image`` with logits = exp(logit_scale) * <image^, text^> + logit_bias and
    loss = sum(-logsigmoid(label * logits)) / loss_div (default: the number of image rows).
    Labels: +1 at ``targets[i]`` (classification) or at column ``i + diag_offset`` (contrastive), -1 elsewhere;
    ``diag_offset < 0`` with no targets = negatives only."""
    loss_div = float(image.shape[0]) if loss_div is None else float(loss_div)
    scale = logit_scale if logit_scale.dim() == 0 else logit_scale.reshape(())
    bias = None if logit_bias is None else (logit_bias if logit_bias.dim() == 0 else logit_bias.reshape(()))
    return _SiglipHeadFn.apply(image, text, scale, bias, targets, diag_offset, normalize_image, normalize_text, loss_div, want_logits)


class TimesformerVideoClassificationHead(nn.Module):
    """Zero-shot video classification head (reference …siglip.py:1640-1726).  ``label_embeddings`` [L, D] are the
    text-tower embeddings of the class prompts (already normalised and averaged over the templates, :1665-1672);
    the reference computes them in ``prepare_multi_task`` with the SigLIP text model, which is outside this
    repo's scope — hand them in with ``set_label_embeddings``."""

    def __init__(self, config=None, label2id=None):
        super().__init__()
        self.config = config
        self.label2id = label2id
        self.logit_scale = nn.Parameter(torch.tensor(2.3026))       # SigLIP init: log(10)
        self.logit_bias = nn.Parameter(torch.tensor(-10.0))
        self.register_buffer("label_embeddings", None, persistent=False)

    def prepare_multi_task(self, text_encoder=None, text_tokenizer=None, logit_scale=None, logit_bias=None, vision_model=None,
                           label_embeddings: Optional[torch.Tensor] = None):
        dev = self.logit_scale.device
        if logit_scale is not None:
            self.logit_scale = nn.Parameter(logit_scale.detach().clone().reshape(()).to(dev))
        if logit_bias is not None:
            self.logit_bias = nn.Parameter(logit_bias.detach().clone().reshape(()).to(dev))
        if label_embeddings is None:
            raise NotImplementedError("the SigLIP text tower is outside this repo's scope (SURVEY §8): pass label_embeddings=[L, D]")
        self.set_label_embeddings(label_embeddings)

    def set_label_embeddings(self, emb: torch.Tensor) -> None:
        self.label_embeddings = emb.detach()

    def forward(self, task_head_input, task_specific_input: Optional[dict] = None):
        image_embeds = task_head_input.pooler_output[:, -1, :]                     # last frame, …siglip.py:1708
        targets = task_specific_input["label"]
        return siglip_head(image_embeds, self.label_embeddings.to(image_embeds.device), self.logit_scale, self.logit_bias,
                           targets=targets, normalize_image=True, normalize_text=False)


class SigLipLoss(nn.Module):
    """Sigmoid contrastive loss (reference …siglip.py:193-297).  The reference exchanges text features round the
    ring of ranks and adds negative-only terms; here the text features are all-gathered once and every rank
    scores its images against all captions with the positives on the diagonal block of its rank — the same sum."""

    def __init__(self, cache_labels=False, rank=0, world_size=1, bidir=True, use_horovod=False, group=None):
        super().__init__()
        assert not use_horovod
        self.rank, self.world_size, self.group = rank, world_size, group

    def forward(self, image_features, text_features, logit_scale, logit_bias, output_dict=False):
        """``logit_scale`` is the EXPONENTIATED scale, as the reference passes it (…siglip.py:2341-2343)."""
        B = image_features.shape[0]
        if self.world_size > 1:
            gathered = torch.empty(self.world_size * B, text_features.shape[1], dtype=text_features.dtype, device=text_features.device)
            dist.all_gather_into_tensor(gathered, text_features.contiguous(), group=self.group)
            text_features = gathered
        # features arrive normalised (…siglip.py:2335-2336): no second normalisation
        loss, _ = siglip_head(image_features, text_features, torch.log(logit_scale), logit_bias, targets=None,
                              diag_offset=self.rank * B, normalize_image=False, normalize_text=False, loss_div=B, want_logits=False)
        return loss


def gathered_classification_loss(pooler_output: torch.Tensor, head: TimesformerVideoClassificationHead, labels: torch.Tensor,
                                 group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The multitask step's use of the one exchange on the path: all-gather the last-frame
    ``pooler_output`` [B_local, D] and the labels over the ranks and evaluate the classification head on the
    GLOBAL batch (identical on every rank).  Inference / metric path: the gather is not differentiated."""
    feats = pooler_output[:, -1, :].contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        w = dist.get_world_size(group)
        all_f = torch.empty(w * feats.shape[0], feats.shape[1], dtype=feats.dtype, device=feats.device)
        all_l = torch.empty(w * labels.shape[0], dtype=labels.dtype, device=labels.device)
        dist.all_gather_into_tensor(all_f, feats, group=group)
        dist.all_gather_into_tensor(all_l, labels.contiguous(), group=group)
        feats, labels = all_f, all_l
    return siglip_head(feats, head.label_embeddings.to(feats.device), head.logit_scale, head.logit_bias, targets=labels)
