"""Task heads on the gathered pooler_output.  CPU: the oracle restatement is pinned to the REAL
reference classes (SigLipLoss, the classification head's forward body) imported from the mount / the staged copy.
GPU: the one-launch sm_100a head (sf_op_siglip_head) against the oracle, forward and gradients."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import streamformer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_module():
    sys.path.insert(0, ROOT)
    from baseline import stage_reference as SR
    if not SR.stage(quiet=True):
        pytest.skip("reference neither mounted nor staged")
    SR.import_reference()
    import importlib
    return importlib.import_module("models.modeling_timesformer_siglip")


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def test_oracle_siglip_loss_matches_reference_class():
    ref = _reference_module()
    rng = np.random.RandomState(0)
    img, txt = _unit(rng.randn(12, 64)).astype(np.float32), _unit(rng.randn(12, 64)).astype(np.float32)
    want = ref.SigLipLoss()(torch.from_numpy(img), torch.from_numpy(txt), torch.tensor(math.exp(2.0)), torch.tensor(-3.0))
    got = O.siglip_loss(img, txt, math.exp(2.0), -3.0)
    assert abs(float(want) - got) <= 1e-4 * abs(got)
    neg = ref.SigLipLoss()._loss(torch.from_numpy(img), torch.from_numpy(txt), torch.tensor(math.exp(2.0)), torch.tensor(-3.0), negative_only=True)
    assert abs(float(neg) - O.siglip_loss(img, txt, math.exp(2.0), -3.0, negative_only=True)) <= 1e-4 * abs(float(neg))


def test_oracle_classification_head_matches_reference_forward():
    """The reference head's forward body (…siglip.py:1704-1726) run on a stub `self` (its __init__ needs the SigLIP
    text tower, which cannot be downloaded here)."""
    ref = _reference_module()
    rng = np.random.RandomState(1)
    pooled = rng.randn(6, 4, 64).astype(np.float32)
    emb = _unit(rng.randn(17, 64)).astype(np.float32)
    labels = rng.randint(0, 17, size=6)

    class Stub:
        label_embeddings = torch.from_numpy(emb)
        logit_scale = torch.tensor(2.3)
        logit_bias = torch.tensor(-4.0)

    class Out:
        pooler_output = torch.from_numpy(pooled)

    loss, logits = ref.TimesformerVideoClassificationHead.forward(Stub(), Out(), {"label": torch.from_numpy(labels)})
    got_loss, got_logits = O.classification_head(pooled, emb, 2.3, -4.0, labels)
    assert abs(float(loss) - got_loss) <= 1e-4 * abs(got_loss)
    np.testing.assert_allclose(got_logits, logits.numpy(), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ GPU
def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max()
    assert err <= atol + rtol * np.abs(b).max(), f"{what}: max abs err {err:.4g} (ref max {np.abs(b).max():.4g})"


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(24, 400), (7, 174), (256, 400), (16, 8)])
def test_classification_head_matches_oracle(B, L):
    """bf16 features in, fp32 arithmetic inside: against the fp64 oracle on the same bf16-rounded inputs the
    logits agree to 2e-3 relative (mma fp32 accumulation order) and the loss to 1e-3."""
    from streamformer_b200.heads import TimesformerVideoClassificationHead
    rng = np.random.RandomState(B * 1000 + L)
    pooled = torch.from_numpy(rng.randn(B, 3, 768).astype(np.float32)).to("cuda", torch.bfloat16)
    emb = torch.from_numpy(_unit(rng.randn(L, 768)).astype(np.float32)).to("cuda", torch.bfloat16)
    labels = torch.from_numpy(rng.randint(0, L, size=B)).cuda()
    head = TimesformerVideoClassificationHead().cuda()
    head.prepare_multi_task(logit_scale=torch.tensor(2.5), logit_bias=torch.tensor(-6.0), label_embeddings=emb)

    class Out:
        pooler_output = pooled

    with torch.no_grad():
        loss, logits = head(Out(), {"label": labels})
    want_loss, want_logits = O.classification_head(pooled.float().cpu().numpy(), emb.float().cpu().numpy(), 2.5, -6.0,
                                                   labels.cpu().numpy())
    assert logits.shape == (B, L) and logits.dtype == torch.float32
    _close(logits.cpu().numpy(), want_logits, 2e-3, 1e-3, "logits")
    assert abs(float(loss) - want_loss) <= 1e-3 * abs(want_loss), (float(loss), want_loss)


@pytest.mark.gpu
def test_siglip_loss_rank_blocks_equal_the_reference_ring_sum():
    """world_size = 3 emulated in one process: rank r scores its images against ALL captions with the positives on
    its diagonal block (diag_offset = r * B); the oracle restates the reference's ring exchange (…siglip.py:245-297)."""
    from streamformer_b200.heads import siglip_head
    rng = np.random.RandomState(5)
    W, B, Dd = 3, 8, 768
    imgs = [_unit(rng.randn(B, Dd)).astype(np.float32) for _ in range(W)]
    txts = [_unit(rng.randn(B, Dd)).astype(np.float32) for _ in range(W)]
    t_img = [torch.from_numpy(x).to("cuda", torch.bfloat16) for x in imgs]
    t_txt = [torch.from_numpy(x).to("cuda", torch.bfloat16) for x in txts]
    want = O.siglip_loss_world([x.float().cpu().numpy() for x in t_img], [x.float().cpu().numpy() for x in t_txt], math.exp(2.0), -5.0)
    all_txt = torch.cat(t_txt, 0)
    for r in range(W):
        with torch.no_grad():
            loss, _ = siglip_head(t_img[r], all_txt, torch.tensor(2.0, device="cuda"), torch.tensor(-5.0, device="cuda"),
                                  diag_offset=r * B, normalize_image=False, loss_div=B, want_logits=False)
        assert abs(float(loss) - want[r]) <= 1e-3 * abs(want[r]), (r, float(loss), want[r])


@pytest.mark.gpu
def test_classification_head_gradients_match_torch_autograd_of_the_reference_formula():
    from streamformer_b200.heads import siglip_head
    rng = np.random.RandomState(6)
    B, L, Dd = 20, 50, 768
    x0 = torch.from_numpy(rng.randn(B, Dd).astype(np.float32)).to("cuda", torch.bfloat16)
    emb = torch.from_numpy(_unit(rng.randn(L, Dd)).astype(np.float32)).to("cuda", torch.bfloat16)
    labels = torch.from_numpy(rng.randint(0, L, size=B)).cuda()
    x = x0.clone().requires_grad_(True)
    s = torch.tensor(2.2, device="cuda", requires_grad=True)
    b = torch.tensor(-3.0, device="cuda", requires_grad=True)
    loss, _ = siglip_head(x, emb, s, b, targets=labels)
    (loss * 1.5).backward()
    # the reference formula (…siglip.py:1704-1726) in fp32 under torch autograd
    xr = x0.float().clone().requires_grad_(True)
    sr = torch.tensor(2.2, device="cuda", requires_grad=True)
    br = torch.tensor(-3.0, device="cuda", requires_grad=True)
    img = xr / xr.norm(p=2, dim=-1, keepdim=True)
    logits = (emb.float() @ img.t() * sr.exp() + br).t()
    tl = -torch.ones_like(logits)
    tl[range(B), labels] = 1
    ref = -torch.nn.functional.logsigmoid(tl * logits).sum() / B
    (ref * 1.5).backward()
    assert abs(float(loss) - float(ref)) <= 1e-3 * abs(float(ref))
    g, gr = x.grad.float(), xr.grad
    rel = float((g - gr).norm() / gr.norm())
    assert rel <= 2e-2, f"d image rel err {rel:.4g}"        # bf16 dlogits / bf16 output
    assert abs(float(s.grad) - float(sr.grad)) <= 2e-2 * abs(float(sr.grad)) + 1e-4
    assert abs(float(b.grad) - float(br.grad)) <= 2e-2 * abs(float(br.grad)) + 1e-4
