#!/usr/bin/env python
"""bench.py — StreamFormer encoder frames/sec on B200 (BASELINE.json metric) in the driver's contract.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, sm_100a)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference arm: CPU oracle port
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N > 1: one rank per GPU, NCCL

A step = one encoder forward (patch embed -> 12 divided space-time blocks -> post-LN -> SigLIP
pooling head) over one batch of synthetic clips [B,16,3,224,224] per GPU (BASELINE.json configs[1]:
B=8, T=16, 224x224, bf16), followed for N>1 by the all-gather of pooler_output.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encoder frames/sec at [B,16,3,224,224]"
UNIT = "frames/s"
# dram__bytes_read.sum + dram__bytes_write.sum summed over the GEMM launches of one cfg2 step
# (per layer: 108.3 + 85.8 + 108.9 + 82.8 + 147.4 + 224.3 MB from profiles/r1_ncu_layer.md, x 12 layers;
# embed and head GEMMs estimated at their algorithmic bytes, 0.25 GB)
NCU_GEMM_TRAFFIC_BYTES = 12 * (108.3 + 85.8 + 108.9 + 82.8 + 147.4 + 224.3) * 1e6 + 0.25e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU (configs[1]: 8)")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fold", action="store_true", help="run temporal out-proj and temporal_dense un-folded")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_clip_seconds(layers: int, frames: int, repeats: int = 1):
    """Time the numpy oracle (port of the reference forward) on one clip with all host threads."""
    from oracle import streamformer_oracle as O
    cfg = O.OracleConfig(num_hidden_layers=layers)
    w = O.make_weights(cfg, seed=0)
    px = O.make_pixels(1, frames, cfg, seed=0)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.forward(w, cfg, px)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference_arm(args):
    """Reference arm for this tier: the reference's CPU implementation of the path.  The reference is
    Python/PyTorch and cannot travel to the GPU box, so this times the oracle port (numpy, all host
    threads) on a bounded sample of the same workload: one clip [1,T,3,224,224] per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget_s = 240.0
    from oracle import streamformer_oracle as O
    cfg = O.OracleConfig(num_hidden_layers=args.layers)
    w = O.make_weights(cfg, seed=0)
    px = O.make_pixels(1, args.frames, cfg, seed=0)
    t_start = time.perf_counter()
    warm = 0
    for _ in range(args.warmup):
        t0 = time.perf_counter()
        O.forward(w, cfg, px)
        warm += 1
        per = time.perf_counter() - t0
        if (time.perf_counter() - t_start) + per * (args.steps + 1) > budget_s:
            break  # keep the whole run within a few minutes
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.forward(w, cfg, px)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    total = sum(times)
    fps = args.frames * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CPU oracle port of the reference forward, one clip [1,{args.frames},3,224,224] per step, "
                               f"{args.layers} layers, fp32 numpy/BLAS on {cores} host threads"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} x 1 clip of {args.frames} frames"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 10 ms through NVML."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from oracle import streamformer_oracle as O  # only for flops accounting + cpu_baseline leg
    from streamformer_b200 import _native as N
    from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    B, T, K, W = args.batch, args.frames, args.steps, max(args.warmup, 3)
    cfg = StreamformerConfig(num_hidden_layers=args.layers, enable_causal_temporal=True,
                             fold_temporal_proj=not args.no_fold)
    torch.manual_seed(0)
    model = TimesformerMultiTaskingModelSigLIP(cfg)
    with torch.no_grad():  # non-trivial gates / time embeddings
        for layer in model.encoder.layer:
            layer.temporal_attention_gating.uniform_(-1, 1)
        model.embeddings.time_embeddings.normal_(0, 0.02)
    model = model.to(dev, dtype).eval()

    S, D = 196, cfg.hidden_size
    NBUF = 4  # rotate inputs: 4 x 38.5 MB (bf16) > L2, and the per-step working set (~0.7 GB) >> 126 MB L2
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_f32 = [torch.randn(B, T, 3, 224, 224, generator=g).pin_memory() for _ in range(2)]
    dev_in = [torch.randn(B, T, 3, 224, 224, device=dev, dtype=dtype) for _ in range(NBUF)]
    gathered = torch.empty(world * B, T, D, device=dev, dtype=dtype) if world > 1 else None

    def step(i):
        out = model(dev_in[i % NBUF])
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.pooler_output)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(W):
            step(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = N.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(K):
            step(i)
        e1.record()
        barrier()
        launches = N.launch_count() - l0
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop()

        # ---- e2e: through the public API with HOST inputs (fp32 pinned, as the reference's loader
        # hands them over), H2D + forward + D2H of pooler_output every step, double-buffered.
        copy_stream = torch.cuda.Stream(dev)
        stage = [torch.empty(B, T, 3, 224, 224, device=dev, dtype=torch.float32) for _ in range(2)]
        host_out = [torch.empty(B, T, D, dtype=dtype).pin_memory() for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(n):
            main = torch.cuda.current_stream(dev)
            for j in range(2):
                consumed[j].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[0])
                stage[0].copy_(host_f32[0], non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(n):
                cur, nxt = i % 2, (i + 1) % 2
                if i + 1 < n:
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(consumed[nxt])
                        stage[nxt].copy_(host_f32[(i + 1) % 2], non_blocking=True)
                        ready[nxt].record(copy_stream)
                main.wait_event(ready[cur])
                out = model(stage[cur])
                consumed[cur].record(main)
                if world > 1:
                    dist.all_gather_into_tensor(gathered, out.pooler_output)
                host_out[cur].copy_(out.pooler_output, non_blocking=True)

        e2e_run(3)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(K)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)

        # ---- per-kernel-class durations, measured in situ (CUDA events around every launch of a
        # full step, on the launching stream) in a separate instrumented pass
        prof = phases = None
        if rank == 0:
            PK = min(K, 5)
            torch.cuda.synchronize()
            N.profile(True)
            for i in range(PK):
                model(dev_in[i % NBUF])
            prof = N.profile_collect()
            N.profile(False)
            for v in prof.values():
                for k in ("ms", "flops", "bytes"):
                    v[k] /= PK
                v["launches"] //= PK
            # per-phase pass: one event pair around each layer's space-time attention block / MLP,
            # kernels inside run back to back exactly as in the timed region
            N.profile(2)
            for i in range(PK):
                model(dev_in[i % NBUF])
            phases = N.profile_collect_phases()
            N.profile(0)

    # max over ranks
    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = load_peaks()
        frames_per_step = world * B * T
        value = frames_per_step * K / (ms_total / 1e3)
        e2e_value = frames_per_step * K / (ms_e2e / 1e3)
        ocfg = O.OracleConfig(num_hidden_layers=args.layers)
        algo_flops_step = O.flops_per_clip(ocfg, T) * B  # per GPU
        gemm = prof["gemm"]
        gemm_tflops = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
        step_kernel_ms = sum(v["ms"] for v in prof.values())
        roofline = {
            "bound": "tensor", "kernel": "gemm_tcgen05_kernel",
            "achieved": gemm_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
            "frac": gemm_tflops / peaks["sustained"],
            "peak_kind": "bf16_tflops_sustained, " + peaks["source"] + " — kernel timed inside a long step",
            "frac_of_burst": gemm_tflops / peaks["burst"],
            "flops_per_step_executed": gemm["flops"], "gemm_ms_per_step": gemm["ms"], "gemm_launches_per_step": gemm["launches"],
            "gemm_share_of_kernel_time": gemm["ms"] / step_kernel_ms if step_kernel_ms else None,
            "traffic": NCU_GEMM_TRAFFIC_BYTES if (B, T, args.layers) == (8, 16, 12) else None,
            "traffic_note": "dram read+write of the 77 GEMM launches of one cfg2 step, from the ncu --set full capture "
                            "of one layer (profiles/r1_ncu_layer.md) x 12 + embed/head; algorithmic bytes "
                            f"{gemm['bytes'] / 1e9:.2f} GB",
            "achieved_unit_note": "executed FLOPs of all GEMM launches of a step / their summed durations",

            "step_algorithmic_tflops": algo_flops_step * world / (ms_total / K * 1e-3) / 1e12 / world,
            "step_frac_of_sustained": algo_flops_step / (ms_total / K * 1e-3) / 1e12 / peaks["sustained"],
            "kernel_ms_per_step": {k: round(v["ms"], 4) for k, v in prof.items() if v["launches"]},
        }
        # BASELINE.json's second metric: tensor-pipe fraction of the space-time attention block of one
        # layer (temporal QKV + causal attention + out-proj.temporal_dense + gate, spatial QKV +
        # attention + out-proj; SURVEY 8d: 35.34 GFLOP per clip and layer algorithmic, un-folded)
        ab = phases["attention_block"]
        ab_ms = ab["ms"] / (args.layers * PK)      # per layer (a layer contributes two timed spans in this mode)
        ab_algo = O.attention_block_flops_per_clip(ocfg, T) * B
        ab_exec = ab_algo - (2.0 * B * T * S * D * D if not args.no_fold else 0.0)
        attention_block = {
            "ms_per_layer": ab_ms, "layers_timed": args.layers * PK,
            "algorithmic_gflop_per_layer": ab_algo / 1e9, "executed_gflop_per_layer": ab_exec / 1e9,
            "algorithmic_tflops": ab_algo / (ab_ms * 1e-3) / 1e12, "executed_tflops": ab_exec / (ab_ms * 1e-3) / 1e12,
            "frac_of_burst_peak_algorithmic": ab_algo / (ab_ms * 1e-3) / 1e12 / peaks["burst"],
            "frac_of_burst_peak_executed": ab_exec / (ab_ms * 1e-3) / 1e12 / peaks["burst"],
            "frac_of_sustained_peak_executed": ab_exec / (ab_ms * 1e-3) / 1e12 / peaks["sustained"],
            "mlp_ms_per_layer": phases["mlp"]["ms"] / (args.layers * PK),
            "embed_ms": phases["embed"]["ms"] / max(phases["embed"]["count"], 1),
            "head_ms": phases["head"]["ms"] / max(phases["head"]["count"], 1),
            "how": "CUDA events around the attention block of every layer (sf_profile mode 2: temporal QKV .. spatial "
                   "attention, then the spatial out-projection), kernels back to back with PDL",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ts = cpu_oracle_clip_seconds(args.layers, T, repeats=1)
            cpu = {"value": T / ts[0], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"1 clip [1,{T},3,224,224], {args.layers} layers, numpy fp32 oracle, {ts[0]:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"configs[1]: single-GPU encoder forward B={B} T={T} 224x224 {args.dtype}, "
                                   f"{args.layers} layers, per GPU; N>1 adds all_gather(pooler_output)",
                       "global_batch": world * B, "frames": T, "parallelism": f"dp{world}",
                       "l2": f"inputs rotate over {NBUF} device buffers; per-step working set (weights 257 MB + "
                             "activations ~0.7 GB) exceeds the 126 MB L2, no explicit flush",
                       "fold_temporal_proj": not args.no_fold, "weights": "random init (no network for checkpoints)"},
            "roofline": roofline,
            "attention_block": attention_block,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * T * 3 * 224 * 224 * 4,
                    "d2h_bytes_per_step": B * T * D * 2, "ms_per_step": ms_e2e / K,
                    "note": "fp32 pinned host clips -> H2D (copy stream, double-buffered) -> model(pixel_values) -> "
                            "pooler_output D2H, every step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
