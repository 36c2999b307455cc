"""Generate the golden fixtures in this directory by running the REAL reference (imported read-only
from /root/reference; it cannot travel to the GPU box) on the oracle's seeded weights and inputs.

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py full_l12   # one case

Each fixture stores the case description (enough for oracle.make_weights / make_pixels to rebuild
the exact inputs anywhere) and the reference's fp32 outputs, sub-sampled where they are large.
The reference has no tests or golden vectors of its own for this path: these files
ARE the pin that ties oracle/streamformer_oracle.py to the reference's behaviour.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("STREAMFORMER_REF", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import streamformer_oracle as O  # noqa: E402

# case name -> description.  "twin" = the KV-cache copy under downstream/VideoQA (always has LoRA).
CASES = {
    # BASELINE.json config 1: README quick-start shape, full depth
    "full_l12": dict(model="root", B=1, T=16, layers=12, causal=True, lora=False, style="reference", seed=0),
    "stress_l2_lora": dict(model="root", B=2, T=16, layers=2, causal=True, lora=True, style="stress", seed=1),
    "oad_t6": dict(model="root", B=1, T=6, layers=2, causal=True, lora=False, style="stress", seed=2),
    "t24_nearest": dict(model="root", B=1, T=24, layers=2, causal=True, lora=False, style="stress", seed=3),
    "bidirectional": dict(model="root", B=1, T=16, layers=2, causal=False, lora=False, style="stress", seed=4),
    "t1_image": dict(model="root", B=2, T=1, layers=2, causal=True, lora=False, style="stress", seed=5),
    "long_t128_l1": dict(model="root", B=1, T=128, layers=1, causal=True, lora=False, style="stress", seed=6),
    # streaming: the twin run as 8+8 frames and as 16 x 1 frames, plus its one-shot forward
    "twin_stream": dict(model="twin", B=1, T=16, layers=2, causal=True, lora=True, style="stress", seed=7,
                        chunks=[[8, 8], [1] * 16]),
    # BASELINE.json config 3 shape (64 appends of one frame, B > 1): the twin built with num_frames=64 so the
    # reference can stream 64 frames at all (it raises at frame num_frames+1, KV:343-348), one-shot and 64 x 1
    "twin_stream64": dict(model="twin", B=2, T=64, layers=2, causal=True, lora=True, style="stress", seed=8,
                          num_frames=64, chunks=[[1] * 64]),
    # the root copy one-shot at T = num_frames = 64 (what the cfg3 stream must equal)
    "root_t64_nf64": dict(model="root", B=1, T=64, layers=2, causal=True, lora=False, style="stress", seed=9,
                          num_frames=64),
    # non-square input: bicubic-antialias position table in the reference's (w0, h0) order (R:380-411),
    # im2col with H != W, spatial attention over 392 tokens
    "nonsquare_224x448": dict(model="root", B=1, T=4, layers=2, causal=True, lora=False, style="stress", seed=10,
                              H=224, W=448),
    # lower resolution (down-sampling branch of the antialias filter), 49 tokens per frame
    "lowres_112": dict(model="root", B=2, T=3, layers=2, causal=True, lora=False, style="stress", seed=11,
                       H=112, W=112),
}

SUB_TOK, SUB_DIM = 14, 8  # last_hidden_state[..., ::14, ::8]


def import_root():
    sys.path.insert(0, REF)
    from models import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # type: ignore
    return StreamformerConfig, TimesformerMultiTaskingModelSigLIP


def import_twin():
    # the twin imports llava.utils.rank0_print; give it a stub instead of the whole LLaVA tree
    llava = types.ModuleType("llava")
    llava_utils = types.ModuleType("llava.utils")
    llava_utils.rank0_print = lambda *a, **k: None
    sys.modules.setdefault("llava", llava)
    sys.modules.setdefault("llava.utils", llava_utils)
    path = os.path.join(REF, "downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py")
    spec = importlib.util.spec_from_file_location("ref_timesformer_encoder", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_timesformer_encoder"] = mod
    spec.loader.exec_module(mod)
    return mod


def oracle_cfg(case) -> O.OracleConfig:
    return O.OracleConfig(num_hidden_layers=case["layers"], enable_causal_temporal=case["causal"],
                          add_lora_spatial=case["lora"], num_frames=case.get("num_frames", 16))


def load_into(model: torch.nn.Module, weights) -> None:
    own = set(model.state_dict().keys())
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items() if k in own or not k.endswith(".mask")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("mask" in m for m in missing), missing


def sub(x: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(x[..., ::SUB_TOK, ::SUB_DIM])


def run_case(name: str) -> None:
    case = CASES[name]
    ocfg = oracle_cfg(case)
    weights = O.make_weights(ocfg, seed=case["seed"], style=case["style"])
    pixels = O.make_pixels(case["B"], case["T"], ocfg, seed=case["seed"], H=case.get("H"), W=case.get("W"))
    out = {"case": json.dumps(case)}
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    if case["model"] == "root":
        Config, Model = import_root()
        cfg = Config(num_hidden_layers=case["layers"], enable_causal_temporal=case["causal"],
                     add_lora_spatial=case["lora"], num_frames=ocfg.num_frames)
        model = Model(cfg).eval()
        load_into(model, weights)
        r = model(torch.from_numpy(pixels), output_hidden_states=True, output_attentions=True)
        out["pooler_output"] = r.pooler_output.numpy()
        out["last_hidden_state_sub"] = sub(r.last_hidden_state.numpy())
        out["hidden_state_1_sub"] = np.ascontiguousarray(r.hidden_states[1].numpy()[:, ::97, ::SUB_DIM])
        out["embedding_sub"] = np.ascontiguousarray(r.hidden_states[0].numpy()[:, ::97, ::SUB_DIM])
        out["attention_0_sub"] = np.ascontiguousarray(r.attentions[0].numpy()[::3, ::5, ::13, :])
    else:
        mod = import_twin()
        cfg = mod.StreamformerConfig(num_hidden_layers=case["layers"], enable_causal_temporal=case["causal"],
                                    num_frames=ocfg.num_frames)
        model = mod.TimesformerMultiTaskingModelSigLIP(cfg).eval()
        load_into(model, weights)
        px = torch.from_numpy(pixels)
        one = model(px)
        out["last_hidden_state_sub"] = sub(one.last_hidden_state.numpy())
        for ci, chunks in enumerate(case["chunks"]):
            from transformers import DynamicCache
            cache = DynamicCache()
            pos = 0
            parts = []
            for n in chunks:
                r = model(px[:, pos:pos + n], use_cache=True, past_key_values=cache)
                cache = r.past_key_values
                parts.append(r.last_hidden_state.numpy())
                pos += n
            out[f"stream_{ci}_last_hidden_state_sub"] = sub(np.concatenate(parts, axis=1))
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
