"""Streaming decode kernel alone (BASELINE configs[2] shapes: 784 sites x 12 heads, cache capacity 64) at several
history lengths; 12 independent (qkv, cache) sets rotate so that nothing is L2-resident (12 x 154 MB)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamformer_b200 import ops

sites, H, cap = 784, 12, 64
dev, dt = "cuda", torch.bfloat16
nrot = 8
qs = [torch.randn(sites, 3 * H * 64, device=dev, dtype=dt) for _ in range(nrot)]
kcs = [torch.randn(sites, H, cap, 64, device=dev, dtype=dt) for _ in range(nrot)]
vcs = [torch.randn(sites, H, cap, 64, device=dev, dtype=dt) for _ in range(nrot)]
for seen in (0, 7, 15, 16, 33, 47, 63):
    for i in range(nrot):
        ops.temporal_decode(qs[i], kcs[i], vcs[i], sites, H, seen, 0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for i in range(reps):
        ops.temporal_decode(qs[i % nrot], kcs[i % nrot], vcs[i % nrot], sites, H, seen, 0.125)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    mb = sites * H * (seen * 2 * 128 + 5 * 128) / 1e6
    print(f"seen={seen:2d}: {us:7.1f} us  {mb:6.1f} MB  {mb / us * 1e-3 * 1e3:6.0f} GB/s", flush=True)
