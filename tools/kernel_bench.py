"""Per-shape timings of the hot kernels through the C ABI (CUDA events, rotating buffers > L2).
    python tools/kernel_bench.py [--B 8] [--T 16]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200 import ops  # noqa: E402
from streamformer_b200 import _native as N  # noqa: E402


REPS, WARM = 20, 3


def timeit(fn, nrot, reps=None, warm=None):
    reps = REPS if reps is None else reps
    warm = WARM if warm is None else warm
    for i in range(warm):
        fn(i % nrot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--T", type=int, default=16)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    global REPS, WARM
    REPS, WARM = a.reps, a.warm
    dev, dt = "cuda", torch.bfloat16
    S, D, I = 196, 768, 3072
    M = a.B * a.T * S
    nrot = int(os.environ.get("SF_KB_NROT", "4"))   # input sets the timed loop rotates over (1: L2-resident inputs)
    res = {}

    def gemm_case(name, Nn, K, **kw):
        As = [torch.randn(M, K, device=dev, dtype=dt) for _ in range(nrot)]
        Ws = [torch.randn(Nn, K, device=dev, dtype=dt) * 0.05 for _ in range(nrot)]
        outs = [torch.empty(M, Nn, device=dev, dtype=dt) for _ in range(nrot)]
        bias = torch.randn(Nn, device=dev)
        extra = {}
        if kw.get("residual"):
            extra["residual_list"] = [torch.randn(M, Nn, device=dev, dtype=dt) for _ in range(nrot)]
        gate = torch.tensor([0.5], device=dev) if kw.get("gate") else None
        ln_stats = ops.rowstats(As[0]).repeat(6, 1, 1).contiguous() if kw.get("ln") else None
        ln_colsum = torch.randn(Nn, device=dev) if kw.get("ln") else None
        stats_out = torch.empty(ops.gemm_stats_parts(M, Nn), M, 2, device=dev) if kw.get("stats") else None

        def fn(i):
            ops.gemm(As[i], Ws[i], bias=bias, act=kw.get("act", 0),
                     residual=extra["residual_list"][i] if "residual_list" in extra else None, gate=gate,
                     row_map=kw.get("row_map", 0), T=a.T, S=S, out=outs[i], ln_stats=ln_stats, ln_colsum=ln_colsum,
                     ln_eps=1e-6, stats_out=stats_out)
        ms = timeit(fn, nrot)
        tf = 2.0 * M * Nn * K / (ms * 1e-3) / 1e12
        res[name] = {"ms": round(ms, 4), "tflops": round(tf, 1)}
        print(f"{name:28s} M={M} N={Nn} K={K}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)

    if a.only and "epi" in a.only:
        gemm_case("qkv (folded LN)", 3 * D, D, ln=True)
        gemm_case("proj (bias)", D, D)
        gemm_case("fc1 (folded LN+gelu)", I, D, act=1, ln=True)
    if not a.only or "gemm" in a.only:
        gemm_case("qkv (bias)", 3 * D, D)
        gemm_case("qkv (folded LN)", 3 * D, D, ln=True)
        gemm_case("proj (bias)", D, D)
        gemm_case("proj (bias+res)", D, D, residual=True)
        gemm_case("proj (bias+res+gate+stats)", D, D, residual=True, gate=True, stats=True)
        gemm_case("fc1 (bias)", I, D)
        gemm_case("fc1 (bias+gelu)", I, D, act=1)
        gemm_case("fc1 (folded LN+gelu)", I, D, act=1, ln=True)
        gemm_case("fc2 (bias+res)", D, I, residual=True)
        gemm_case("fc2 (bias+res+stats)", D, I, residual=True, stats=True)
        gemm_case("head kv (bias)", 2 * D, D)

    if a.only and "k64" in a.only:
        # K = 64 (one K block): the main loop is negligible, so these time the EPILOGUE + store path
        gemm_case("k64 N=2304 (bias)", 3 * D, 64)
        gemm_case("k64 N=2304 (folded LN)", 3 * D, 64, ln=True)
        gemm_case("k64 N=768 (bias)", D, 64)
        gemm_case("k64 N=768 (bias+res+gate+stats)", D, 64, residual=True, gate=True, stats=True)
        gemm_case("k64 N=3072 (bias)", I, 64)
        gemm_case("k64 N=3072 (folded LN+gelu)", I, 64, act=1, ln=True)
        gemm_case("k256 N=2304 (bias)", 3 * D, 256)
        gemm_case("k1536 N=2304 (bias)", 3 * D, 1536)

    if a.only and "cublas" in a.only:
        # library reference on the same shapes (context only: cuBLAS is not on the product path)
        for name, Nn, K in (("qkv", 3 * D, D), ("proj", D, D), ("fc1", I, D), ("fc2", D, I)):
            As = [torch.randn(M, K, device=dev, dtype=dt) for _ in range(nrot)]
            Ws = [torch.randn(Nn, K, device=dev, dtype=dt) * 0.05 for _ in range(nrot)]
            outs = [torch.empty(M, Nn, device=dev, dtype=dt) for _ in range(nrot)]
            bias = torch.randn(Nn, device=dev, dtype=dt)
            ms = timeit(lambda i: torch.matmul(As[i], Ws[i].t(), out=outs[i]), nrot)
            tf = 2.0 * M * Nn * K / (ms * 1e-3) / 1e12
            ms2 = timeit(lambda i: torch.addmm(bias, As[i], Ws[i].t(), out=outs[i]), nrot)
            tf2 = 2.0 * M * Nn * K / (ms2 * 1e-3) / 1e12
            res["cublas " + name] = {"ms": round(ms, 4), "tflops": round(tf, 1), "addmm_tflops": round(tf2, 1)}
            print(f"cuBLAS {name:8s} M={M} N={Nn} K={K}: {ms*1e3:8.1f} us {tf:7.1f} TFLOP/s | addmm(bias) {ms2*1e3:8.1f} us {tf2:7.1f}", flush=True)

    if not a.only or "ln" in a.only:
        xs = [torch.randn(M, D, device=dev, dtype=dt) for _ in range(nrot)]
        gmm, bta = torch.randn(D, device=dev), torch.randn(D, device=dev)
        ms = timeit(lambda i: ops.layernorm(xs[i], gmm, bta, 1e-6), nrot)
        gbs = 4.0 * M * D / (ms * 1e-3) / 1e9
        res["layernorm"] = {"ms": round(ms, 4), "gbs": round(gbs, 1)}
        print(f"layernorm M={M}: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s", flush=True)
        ms = timeit(lambda i: ops.layernorm(xs[i], gmm, bta, 1e-6, row_map=2, T=a.T, S=S), nrot)
        print(f"layernorm (BNT->BTN) M={M}: {ms*1e3:8.1f} us  {4.0*M*D/(ms*1e-3)/1e9:7.1f} GB/s", flush=True)

    if not a.only or "attn" in a.only:
        qs = [torch.randn(M, 3 * D, device=dev, dtype=dt) for _ in range(nrot)]
        ms = timeit(lambda i: ops.spatial_attention(qs[i], a.B * a.T, 12, S, 0.125, T_inner=a.T), nrot)
        fl = 4.0 * a.B * a.T * 12 * S * S * 64
        res["spatial_attention"] = {"ms": round(ms, 4), "tflops": round(fl / (ms * 1e-3) / 1e12, 1),
                                    "gbs": round(8.0 * M * D / (ms * 1e-3) / 1e9, 1)}
        print(f"spatial attention: {ms*1e3:8.1f} us  {fl/(ms*1e-3)/1e12:6.1f} TFLOP/s  {8.0*M*D/(ms*1e-3)/1e9:7.1f} GB/s", flush=True)
        ms1 = timeit(lambda i: ops.spatial_attention(qs[i], a.B * a.T, 12, S, 0.125, T_inner=1), nrot)
        print(f"spatial attention (frames contiguous, T_inner=1): {ms1*1e3:8.1f} us", flush=True)
        ms = timeit(lambda i: ops.temporal_attention(qs[i], a.B * S, 12, a.T, True, 0.125), nrot)
        res["temporal_attention"] = {"ms": round(ms, 4), "gbs": round(8.0 * M * D / (ms * 1e-3) / 1e9, 1)}
        print(f"temporal attention: {ms*1e3:8.1f} us  {8.0*M*D/(ms*1e-3)/1e9:7.1f} GB/s", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
