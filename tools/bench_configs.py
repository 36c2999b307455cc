"""Throughput of the BASELINE.json configs that are not the bench.py headline (one JSON line each):
    cfg3  streaming KV-cache path: 64 appends of T=1 at B=4 (per-step latency, frames/s)
    cfg4  per-GPU share of the batch-sharded step: B=32, T=16 forward (no gather on 1 GPU)
    cfg5  long clip: B=2, T=128
    python tools/bench_configs.py [--only cfg3,cfg5] [--layers 12]
CUDA events on the launching stream, warm-up first; inputs resident on the device in bf16.
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200 import _native as N  # noqa: E402
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP  # noqa: E402


def make_model(layers, num_frames=16):
    cfg = StreamformerConfig(num_hidden_layers=layers, enable_causal_temporal=True, num_frames=num_frames)
    torch.manual_seed(0)
    m = TimesformerMultiTaskingModelSigLIP(cfg)
    with torch.no_grad():
        for layer in m.encoder.layer:
            layer.temporal_attention_gating.uniform_(-1, 1)
        m.embeddings.time_embeddings.normal_(0, 0.02)
    return m.to("cuda", torch.bfloat16).eval()


def forward_case(name, layers, B, T, warm=3, reps=8):
    model = make_model(layers)
    xs = [torch.randn(B, T, 3, 224, 224, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
    with torch.no_grad():
        for i in range(warm):
            model(xs[i % 2])
        torch.cuda.synchronize()
        l0 = N.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            model(xs[i % 2])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"config": name, "B": B, "T": T, "layers": layers, "ms_per_step": round(ms, 3),
                      "frames_per_s": round(B * T / ms * 1e3, 1), "gpu_launches_per_step": (N.launch_count() - l0) // reps,
                      "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
    del model, xs
    torch.cuda.empty_cache()


def streaming_case(layers, B=4, steps=64, rounds=3):
    model = make_model(layers, num_frames=steps)
    frames = [torch.randn(B, 1, 3, 224, 224, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    cache = model.new_kv_cache(B, max_frames=steps)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    totals, per_step = [], None
    with torch.no_grad():
        for r in range(rounds + 1):          # round 0 = warm-up
            cache.reset()
            torch.cuda.synchronize()
            l0 = N.launch_count()
            ev[0].record()
            for s in range(steps):
                model(frames[s % 4], past_key_values=cache, use_cache=True)
                ev[s + 1].record()
            torch.cuda.synchronize()
            launches = N.launch_count() - l0
            if r:
                totals.append(ev[0].elapsed_time(ev[steps]))
                per_step = [ev[s].elapsed_time(ev[s + 1]) for s in range(steps)]
    # per-kernel-class GPU time of one stream (events around every launch; direct launches, no graph)
    with torch.no_grad():
        cache.reset()
        torch.cuda.synchronize()
        N.profile(1)
        for s in range(steps):
            model(frames[s % 4], past_key_values=cache, use_cache=True)
        prof = N.profile_collect()
        N.profile(0)
    kernel_ms = {k: round(v["ms"] / steps, 4) for k, v in prof.items() if v["launches"]}
    tot = statistics.median(totals)
    print(json.dumps({"config": "cfg3 streaming", "B": B, "steps": steps, "layers": layers,
                      "total_ms": round(tot, 2), "frames_per_s": round(B * steps / tot * 1e3, 1),
                      "ms_per_step_mean": round(tot / steps, 4),
                      "ms_step_1_16_32_64": [round(per_step[i], 4) for i in (0, 15, 31, steps - 1)],
                      "gpu_launches_per_step": launches // steps, "graph_steps": cache.graph_launches,
                      "kernel_ms_per_step": kernel_ms,
                      "kv_cache_gb": round(2 * layers * B * 196 * 12 * steps * 64 * 2 / 2**30, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg3,cfg4,cfg5")
    ap.add_argument("--layers", type=int, default=12)
    a = ap.parse_args()
    want = set(a.only.split(","))
    if "cfg3" in want:
        streaming_case(a.layers)
    if "cfg5" in want:
        forward_case("cfg5 long clip", a.layers, 2, 128)
    if "cfg4" in want:
        forward_case("cfg4 per-GPU shard", a.layers, 32, 16, warm=2, reps=4)


if __name__ == "__main__":
    main()
