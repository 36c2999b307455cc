#!/usr/bin/env python
"""bench.py — StreamFormer encoder frames/sec on B200 (BASELINE.json metric) in the driver's contract.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, sm_100a)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference arm: the reference's own PyTorch
                                                                   # classes on the host cores (baseline/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N > 1: one rank per GPU, NCCL

A step = one encoder forward (patch embed -> 12 divided space-time blocks -> post-LN -> SigLIP pooling head)
over one batch of synthetic clips [B,16,3,224,224] per GPU (BASELINE.json configs[1]: B=8, T=16, 224x224,
bf16), followed for N>1 by the all-gather of pooler_output.  Rank 0 prints ONE JSON line.
The other BASELINE configs (cfg3 streaming, cfg4 per-GPU shard of the global-256 step, cfg5 long clip) are
measured in the same run and reported under "configs".
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
# geometry of the reference model (StreamformerConfig defaults)
D, HEADS, MLP, PATCHES, KP = 768, 12, 3072, 196, 3 * 16 * 16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU (configs[1]: 8)")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 sub-measurements")
    ap.add_argument("--no-fold", action="store_true", help="run temporal out-proj and temporal_dense un-folded")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------- work accounting
def flops_per_clip(T: int, layers: int) -> float:
    """Algorithmic FLOPs of one clip (BASELINE.md §4 / SURVEY §8d: 2*M*N*K, full non-causal attention count,
    no LoRA, un-folded reference graph): 790.5 G at T=16, 12 layers."""
    M = T * PATCHES
    per_layer = 2 * M * D * 3 * D * 2 + 2 * M * D * D * 3 + 2 * 2 * M * T * D + 2 * 2 * M * PATCHES * D + 2 * 2 * M * D * MLP
    head = 2 * M * D * 2 * D + 2 * 2 * T * PATCHES * D + 2 * T * D * D + 2 * 2 * T * D * MLP
    return float(2 * M * KP * D + layers * per_layer + head)


def attention_block_flops_per_clip(T: int) -> float:
    """One layer's space-time attention block incl. its projections, excl. the MLP (35.34 G at T=16)."""
    M = T * PATCHES
    return float(2 * M * D * 3 * D * 2 + 2 * M * D * D * 3 + 2 * 2 * M * T * D + 2 * 2 * M * PATCHES * D)


# ------------------------------------------------------------------------------------- CPU arms
def host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core regardless and reports the count."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def reference_model(layers: int):
    """The UNMODIFIED reference classes from baseline/_ref (staged by baseline/stage_reference.py), random init
    with non-trivial gates / time embeddings, eval, fp32.  None when the staged copy is absent."""
    import torch
    try:
        from baseline.stage_reference import import_reference, stage
        if not stage(quiet=True):
            return None
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            Config, Model = import_reference()
            torch.manual_seed(0)
            model = Model(Config(num_hidden_layers=layers, enable_causal_temporal=True)).eval()
        with torch.no_grad():
            for layer in model.encoder.layer:
                layer.temporal_attention_gating.uniform_(-1, 1)
            model.embeddings.time_embeddings.normal_(0, 0.02)
        return model
    except Exception as e:  # noqa: BLE001
        print(f"[bench] reference import failed: {e!r}", file=sys.stderr)
        return None


def cpu_forward_fn(layers: int, frames: int):
    """(callable running one clip [1,T,3,224,224] on the host, kind, threads)."""
    import torch
    threads = host_threads()
    model = reference_model(layers)
    if model is not None:
        px = torch.randn(1, frames, 3, 224, 224)

        def run():
            with torch.no_grad():
                return model(px).pooler_output
        return run, "reference", threads
    from oracle import streamformer_oracle as O   # fallback: the numpy port (cpu_baseline leg only)
    cfg = O.OracleConfig(num_hidden_layers=layers)
    w = O.make_weights(cfg, seed=0)
    px = O.make_pixels(1, frames, cfg, seed=0)
    return (lambda: O.forward(w, cfg, px)), "port", threads


def run_reference_arm(args):
    """Reference arm for this tier: the reference's own CPU implementation of the path — its PyTorch classes,
    imported unmodified from baseline/_ref — on the box's host cores, one clip [1,T,3,224,224] per step
    (a bounded sample of the workload), all host threads.  Rank 0 only; other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, kind, threads = cpu_forward_fn(args.layers, args.frames)
    budget_s = 240.0
    t_start = time.perf_counter()
    warm = 0
    for _ in range(max(args.warmup, 1)):
        t0 = time.perf_counter()
        run()
        warm += 1
        per = time.perf_counter() - t0
        if (time.perf_counter() - t_start) + per * (args.steps + 1) > budget_s:
            break  # keep the whole run within a few minutes
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    total = sum(times)
    fps = args.frames * len(times) / total
    what = ("the reference's TimesformerMultiTaskingModelSigLIP (unmodified, baseline/_ref), PyTorch eager fp32"
            if kind == "reference" else "numpy port of the reference forward (baseline/_ref not staged)")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{what}, one clip [1,{args.frames},3,224,224] per step, {args.layers} layers, on {threads} "
                               f"host threads (rank 0 only; the CPU arm does not scale with --gpus)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{len(times)} x 1 clip of {args.frames} frames, median {statistics.median(times):.2f} s/clip"},
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
def make_model(torch, layers, dev, dtype, num_frames=16, fold=True):
    from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP
    cfg = StreamformerConfig(num_hidden_layers=layers, enable_causal_temporal=True, fold_temporal_proj=fold,
                             num_frames=num_frames)
    torch.manual_seed(0)
    model = TimesformerMultiTaskingModelSigLIP(cfg)
    with torch.no_grad():  # non-trivial gates / time embeddings
        for layer in model.encoder.layer:
            layer.temporal_attention_gating.uniform_(-1, 1)
        model.embeddings.time_embeddings.normal_(0, 0.02)
    return model.to(dev, dtype).eval()


def timed_steps(torch, fn, warm, reps):
    """Per-step CUDA-event times [ms] of `reps` calls of fn(i) after `warm` untimed ones."""
    for i in range(warm):
        fn(i)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(reps):
        fn(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]


def bench_cfg3(torch, N, args, dev, dtype, peaks):
    """BASELINE configs[2]: streaming KV-cache path, 64 appends of one frame at B=4 (replicas only: the
    state is per stream, so this runs on one GPU)."""
    B, steps = 4, 64
    model = make_model(torch, args.layers, dev, dtype, num_frames=steps)
    frames = [torch.randn(B, 1, 3, 224, 224, device=dev, dtype=dtype) for _ in range(4)]
    cache = model.new_kv_cache(B, max_frames=steps)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    totals, per_step, launches, host_us = [], None, 0, None
    with torch.no_grad():
        for r in range(4):          # round 0 = warm-up (captures the graph)
            cache.reset()
            torch.cuda.synchronize()
            l0 = N.launch_count()
            t0 = time.perf_counter()
            ev[0].record()
            for s in range(steps):
                model(frames[s % 4], past_key_values=cache, use_cache=True)
                ev[s + 1].record()
            host = time.perf_counter() - t0
            torch.cuda.synchronize()
            launches = N.launch_count() - l0
            if r:
                totals.append(ev[0].elapsed_time(ev[steps]))
                per_step = [ev[s].elapsed_time(ev[s + 1]) for s in range(steps)]
                host_us = host / steps * 1e6
    tot = statistics.median(totals)
    ms = tot / steps
    # floor of a step: max(GEMM FLOPs at tensor peak, bytes that must move at HBM peak): the weights are read
    # once per step (M = 784 rows cannot amortise them) + the K/V history of every layer, growing with the step
    gemm_flops = flops_per_clip(1, args.layers) * B
    weight_bytes = args.layers * (2 * 3 * D * D + 2 * D * D + 2 * D * MLP) * 2 + (2 * D * D + D * D + 2 * D * MLP + KP * D) * 2
    kv_bytes_mean = args.layers * 2 * B * PATCHES * HEADS * 64 * 2 * (steps + 1) / 2
    floor_ms = max(gemm_flops / (peaks["burst"] * 1e12), (weight_bytes + kv_bytes_mean) / (peaks["hbm"] * 1e9)) * 1e3
    return {"workload": f"configs[2]: streaming KV cache, {steps} x (B={B}, T=1) appends, {args.layers} layers, CUDA-graph replay",
            "frames_per_s": B * steps / tot * 1e3, "ms_per_step": ms,
            "ms_step_1_16_32_64": [round(per_step[i], 4) for i in (0, 15, 31, steps - 1)],
            "host_us_per_step": host_us, "gpu_launches_per_step": launches // steps, "graph_steps": cache.graph_launches,
            "kv_cache_gb": 2 * args.layers * B * PATCHES * HEADS * steps * 64 * 2 / 2**30,
            "roofline": {"bound": "max(tensor, hbm)", "floor_ms_per_step": floor_ms, "frac": floor_ms / ms,
                         "gemm_ms_at_burst_peak": gemm_flops / (peaks["burst"] * 1e12) * 1e3,
                         "weight_plus_mean_kv_bytes": weight_bytes + kv_bytes_mean,
                         "note": "floor = max(step GEMM FLOPs / burst bf16 peak, (packed weights + mean K/V history) / HBM peak)"}}


def bench_forward(torch, dist, N, args, dev, dtype, peaks, name, B, T, world, warm, reps):
    """One-shot forward of B clips x T frames per GPU (+ all_gather of pooler_output at N > 1)."""
    model = make_model(torch, args.layers, dev, dtype)
    xs = [torch.randn(B, T, 3, 224, 224, device=dev, dtype=dtype) for _ in range(2)]
    gathered = torch.empty(world * B, T, D, device=dev, dtype=dtype) if world > 1 else None
    with torch.no_grad():
        def fn(i):
            out = model(xs[i % 2])
            if world > 1:
                dist.all_gather_into_tensor(gathered, out.pooler_output)
        if world > 1:
            dist.barrier()
        ts = timed_steps(torch, fn, warm, reps)
    t = torch.tensor([sum(ts)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / reps
    fl = flops_per_clip(T, args.layers) * B
    res = {"workload": name, "global_batch": world * B, "frames": T, "ms_per_step": ms,
           "ms_per_step_median_rank0": statistics.median(ts), "frames_per_s": world * B * T / ms * 1e3,
           "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peaks["burst"], "unit": "TFLOP/s",
                        "frac": fl / (ms * 1e-3) / 1e12 / peaks["burst"],
                        "note": "algorithmic FLOPs of the step (un-folded reference graph) per GPU / step time, vs burst bf16 peak"},
           "mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    del model, xs
    torch.cuda.empty_cache()
    return res


def bench_train_step(torch, dist, N, args, dev, dtype, peaks, B, T, world, warm, reps):
    """BASELINE configs[3] as a real pre-train step: forward -> all_gather(last-frame pooler_output, labels) ->
    classification head (SigLIP sigmoid loss over the global batch) -> native backward through the encoder ->
    (N > 1) all-reduce of the flattened gradients.  No optimizer update in the timed region (the reference's
    optimizer is DeepSpeed's, outside the path); weights therefore stay bound."""
    from streamformer_b200.heads import TimesformerVideoClassificationHead, siglip_head
    model = make_model(torch, args.layers, dev, dtype).train()
    head = TimesformerVideoClassificationHead().to(dev)
    torch.manual_seed(1)
    head.set_label_embeddings(torch.nn.functional.normalize(torch.randn(400, D, device=dev), dim=-1).to(dtype))
    xs = [torch.randn(B, T, 3, 224, 224, device=dev, dtype=dtype) for _ in range(2)]
    labels = torch.randint(0, 400, (B,), device=dev)
    params = [p for p in model.parameters() if p.requires_grad]

    def fn(i):
        for p in params:
            p.grad = None
        out = model(xs[i % 2])
        feats = out.pooler_output[:, -1, :]
        if world > 1:
            # the loss needs the global batch: gather features and labels; each rank back-propagates the gradient
            # of ITS rows of the global loss
            all_f = torch.empty(world * B, D, device=dev, dtype=dtype)
            all_l = torch.empty(world * B, dtype=labels.dtype, device=dev)
            dist.all_gather_into_tensor(all_f, feats.detach().contiguous())
            dist.all_gather_into_tensor(all_l, labels)
        loss, _ = siglip_head(feats, head.label_embeddings, head.logit_scale, head.logit_bias, targets=labels,
                              loss_div=float(world * B), want_logits=False)
        loss.backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
        return loss

    if world > 1:
        dist.barrier()
    ts = timed_steps(torch, fn, warm, reps)
    t = torch.tensor([sum(ts)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / reps
    fl = 3.0 * flops_per_clip(T, args.layers) * B      # forward + 2x for dgrad / wgrad (recompute not counted)
    res = {"workload": f"configs[3]: pre-train step, {B} clips/GPU x {world} GPU(s) (global B={B * world}), T={T}: forward + gather + "
                       "classification-head loss + native backward" + (" + gradient all-reduce" if world > 1 else "") + "; no optimizer update",
           "global_batch": world * B, "ms_per_step": ms, "frames_per_s": world * B * T / ms * 1e3,
           "backward_included": True,
           "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peaks["burst"], "unit": "TFLOP/s",
                        "frac": fl / (ms * 1e-3) / 1e12 / peaks["burst"],
                        "note": "3 x algorithmic forward FLOPs (forward, dgrad, wgrad; the layer-wise recompute is extra, executed work) / step time"},
           "mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    del model, xs
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    from streamformer_b200 import _native as N

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
    model = make_model(torch, args.layers, dev, dtype, fold=not args.no_fold)
    peaks = load_peaks()

    S = PATCHES
    NBUF = 4  # rotate inputs: 4 x 38.5 MB (bf16) > L2, and the per-step working set (~0.7 GB) >> 126 MB L2
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    torch.manual_seed(4321 + rank)          # the ranks work on DIFFERENT clips (make_model seeded every rank identically)
    # e2e inputs: uint8 frames as a video decoder hands them over ([B,T,H,W,3]); the loader's
    # ClipToTensor + Normalize(0.5, 0.5) runs inside the patch-embedding front end on the GPU
    host_u8 = [torch.randint(0, 256, (B, T, 224, 224, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
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

    gather_check = None
    with torch.no_grad():
        for i in range(W):
            step(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = N.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        barrier()
        ev[0].record()
        for i in range(K):
            step(i)
            ev[i + 1].record()
        barrier()
        launches = N.launch_count() - l0
        ms_total = ev[0].elapsed_time(ev[K])
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
        clocks = sampler.stop()

        # ---- multi-rank correctness of the one exchange step (SURVEY §4 vi): every rank's slice of the
        # gathered tensor is its local pooler_output, and all ranks hold the same gathered tensor
        if world > 1:
            out = step(0)
            torch.cuda.synchronize()
            ok = torch.equal(gathered[rank * B:(rank + 1) * B], out.pooler_output)
            for r in range(world):
                ok = ok and bool(torch.isfinite(gathered[r * B:(r + 1) * B].float()).all())
            digest = gathered.float().double().sum().reshape(1)
            digests = [torch.empty_like(digest) for _ in range(world)]
            dist.all_gather(digests, digest)
            same = all(float(d) == float(digests[0]) for d in digests)
            distinct = len({float(gathered[r * B:(r + 1) * B].float().double().sum()) for r in range(world)}) == world
            flag = torch.tensor([int(ok and same and distinct)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            gather_check = {"ok": bool(int(flag)), "what": "gathered[r*B:(r+1)*B] == local pooler_output on every rank, finite, "
                            "identical checksum on all ranks, ranks' slices differ (different seeds)"}
            if not gather_check["ok"]:
                raise SystemExit(f"rank {rank}: all_gather(pooler_output) check failed")

        # ---- e2e: through the public API with HOST inputs: pinned uint8 frames -> H2D (copy stream,
        # double-buffered) -> model(pixel_values) -> pooler_output D2H, every step
        copy_stream = torch.cuda.Stream(dev)
        stage = [torch.empty(B, T, 224, 224, 3, device=dev, dtype=torch.uint8) for _ in range(2)]
        host_out = [torch.empty(B, T, D, dtype=dtype).pin_memory() for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(n):
            main = torch.cuda.current_stream(dev)
            for j in range(2):
                consumed[j].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[0])
                stage[0].copy_(host_u8[0], non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(n):
                cur, nxt = i % 2, (i + 1) % 2
                if i + 1 < n:
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(consumed[nxt])
                        stage[nxt].copy_(host_u8[(i + 1) % 2], non_blocking=True)
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
        PK = min(K, 5)
        if rank == 0:
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

    del dev_in, stage
    torch.cuda.empty_cache()
    configs = {}
    if not args.no_configs and (B, T) == (8, 16):
        # BASELINE.md §5 protocol for configs[1]: >= 10 warm-up, >= 50 timed, median of per-step event times
        with torch.no_grad():
            xs = [torch.randn(B, T, 3, 224, 224, device=dev, dtype=dtype) for _ in range(NBUF)]
            ts = timed_steps(torch, lambda i: model(xs[i % NBUF]), 10, 50) if rank == 0 else None
            del xs
        if rank == 0:
            med = statistics.median(ts)
            configs["cfg2_median"] = {"workload": "configs[1]: B=8 T=16, 10 warm-up + 50 timed forwards, median per-step CUDA-event time (rank 0, no gather)",
                                      "ms_per_step_median": med, "frames_per_s": B * T / med * 1e3,
                                      "frac_of_burst": flops_per_clip(T, args.layers) * B / (med * 1e-3) / 1e12 / peaks["burst"]}
        del model
        torch.cuda.empty_cache()
        # configs[3]: per-GPU shard of the global-256 step (32 clips / GPU; at N=8 the global batch is 256)
        configs["cfg4_shard"] = bench_forward(torch, dist, N, args, dev, dtype, peaks,
                                              f"configs[3]: batch-sharded forward, 32 clips/GPU x {world} GPU(s) (global B={32 * world}), T=16, "
                                              "forward + all_gather(pooler_output); backward not included in this line", 32, 16, world, 2, 6)
        configs["cfg4_train_step"] = bench_train_step(torch, dist, N, args, dev, dtype, peaks, 32, 16, world, 1, 3)
        if world == 1:
            configs["cfg5_long_clip"] = bench_forward(torch, dist, N, args, dev, dtype, peaks,
                                                      "configs[4]: long clip B=2 T=128 224x224 bf16, 1 GPU", 2, 128, 1, 3, 10)
            configs["cfg3_streaming"] = bench_cfg3(torch, N, args, dev, dtype, peaks)

    if rank == 0:
        frames_per_step = world * B * T
        value = frames_per_step * K / (ms_total / 1e3)
        e2e_value = frames_per_step * K / (ms_e2e / 1e3)
        algo_flops_step = flops_per_clip(T, args.layers) * B  # per GPU
        gemm = prof["gemm"]
        gemm_tflops = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
        step_kernel_ms = sum(v["ms"] for v in prof.values())
        # the timed region is K steps of a few ms: a burst, not the multi-second / power-limited regime the
        # sustained peak was measured under -> burst peak unless the region ran for >= 2 s
        long_run = ms_total >= 2000.0
        peak = peaks["sustained"] if long_run else peaks["burst"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")
        if os.path.exists(tpath) and (B, T, args.layers) == (8, 16, 12):
            traffic = json.load(open(tpath))
        roofline = {
            "bound": "tensor", "kernel": "gemm_tcgen05_kernel",
            "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s", "frac": gemm_tflops / peak,
            "peak_kind": ("bf16_tflops_sustained" if long_run else "bf16_tflops (burst)") + ", " + peaks["source"]
                         + f"; timed region {ms_total / 1e3:.2f} s",
            "frac_of_burst": gemm_tflops / peaks["burst"], "frac_of_sustained": gemm_tflops / peaks["sustained"],
            "flops_per_step_executed": gemm["flops"], "gemm_ms_per_step": gemm["ms"], "gemm_launches_per_step": gemm["launches"],
            "gemm_share_of_kernel_time": gemm["ms"] / step_kernel_ms if step_kernel_ms else None,
            "traffic": traffic["dram_bytes_per_step"] if traffic else None,
            "traffic_note": (traffic["how"] if traffic else "no ncu --set full capture of this build committed") +
                            f"; algorithmic bytes {gemm['bytes'] / 1e9:.2f} GB",
            "achieved_unit_note": "executed FLOPs of all GEMM launches of a step / their summed durations (CUDA events around "
                                  "every launch, on the launching stream)",
            "step_algorithmic_tflops": algo_flops_step / (ms_total / K * 1e-3) / 1e12,
            "step_frac_of_burst": algo_flops_step / (ms_total / K * 1e-3) / 1e12 / peaks["burst"],
            "kernel_ms_per_step": {k: round(v["ms"], 4) for k, v in prof.items() if v["launches"]},
        }
        # BASELINE.json's second metric: tensor-pipe fraction of the space-time attention block of one
        # layer (temporal QKV + causal attention + out-proj.temporal_dense + gate, spatial QKV +
        # attention + out-proj; SURVEY 8d: 35.34 GFLOP per clip and layer algorithmic, un-folded)
        ab = phases["attention_block"]
        ab_ms = ab["ms"] / (args.layers * PK)      # per layer (a layer contributes two timed spans in this mode)
        ab_algo = attention_block_flops_per_clip(T) * B
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
            run, kind, threads = cpu_forward_fn(args.layers, T)
            run()                                    # warm-up (lazy init, page-in)
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            cpu = {"value": T / dt, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"1 clip [1,{T},3,224,224], {args.layers} layers, fp32, {dt:.1f} s after one warm-up clip "
                             + ("(the reference's own PyTorch classes from baseline/_ref)" if kind == "reference" else "(numpy port)")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "ms_per_step_median": statistics.median(step_ms), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"configs[1]: single-GPU encoder forward B={B} T={T} 224x224 {args.dtype}, "
                                   f"{args.layers} layers, per GPU; N>1 adds all_gather(pooler_output)",
                       "global_batch": world * B, "frames": T, "parallelism": f"dp{world}",
                       "l2": f"inputs rotate over {NBUF} device buffers; per-step working set (weights 257 MB + "
                             "activations ~0.7 GB) exceeds the 126 MB L2, no explicit flush",
                       "fold_temporal_proj": not args.no_fold, "weights": "random init (no network for checkpoints)"},
            "roofline": roofline,
            "attention_block": attention_block,
            "configs": configs,
            "gather_check": gather_check,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * T * 3 * 224 * 224,
                    "d2h_bytes_per_step": B * T * D * 2, "ms_per_step": ms_e2e / K,
                    "note": "pinned uint8 host frames [B,T,224,224,3] (decoder layout) -> H2D (copy stream, double-buffered) -> "
                            "model(pixel_values): normalise(0.5,0.5) + patch embed on the GPU -> pooler_output D2H, every step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
