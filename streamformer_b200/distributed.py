"""Batch-sharded data parallelism for the encoder.

Clips are independent through the whole encoder, so the path shards on the batch axis with no
data-path collective: rank r of W runs ``model(pixel_values[shard])`` on its own GPU (weights
replicated, 257 MB bf16).  The one exchange step is an all-gather of ``pooler_output`` when a
multitask loss needs the full batch (the reference instead sends text features round a P2P ring,
models/modeling_timesformer_siglip.py:92-146, 244-295; DDP / init code: utils.py:372-445).
One process per GPU, ``torch.distributed`` over NCCL/NVLink; the gloo backend is supported so the
host logic can be tested on CPU with world_size > 1.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_clips", "gather_pooler_output", "sharded_forward"]


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank's clips; the first ``global_batch % world`` ranks take one extra."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    if global_batch < 0:
        raise ValueError("global_batch must be >= 0")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _rank_world(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_clips(pixel_values: torch.Tensor, group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """This rank's contiguous slice of a [B_global, T, C, H, W] batch (a view, no copy)."""
    rank, world = _rank_world(group)
    s, e = shard_bounds(pixel_values.shape[0], rank, world)
    return pixel_values[s:e]


def gather_pooler_output(pooled: torch.Tensor, global_batch: Optional[int] = None,
                         group: Optional["dist.ProcessGroup"] = None,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather of ``pooler_output`` [B_local, T, D] -> [B_global, T, D], in rank order.

    Equal shards (the bench / training case) use one ``all_gather_into_tensor`` (a single
    ncclAllGather over NVLink, 786 KB per rank at global B=256 — latency-bound).  Ragged shards
    (``global_batch % world != 0``) are padded to the largest shard and trimmed after the gather.
    ``global_batch`` may be omitted when every rank holds the same number of clips.
    """
    rank, world = _rank_world(group)
    if world == 1:
        return pooled
    pooled = pooled.contiguous()
    b_local = pooled.shape[0]
    if global_batch is None:
        global_batch = b_local * world
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    counts = [e - s for s, e in sizes]
    if counts[rank] != b_local:
        raise ValueError(f"rank {rank} holds {b_local} clips, expected {counts[rank]} of {global_batch}")
    b_max = max(counts)
    tail = tuple(pooled.shape[1:])
    if min(counts) == b_max:
        if out is None:
            out = torch.empty((global_batch, *tail), dtype=pooled.dtype, device=pooled.device)
        dist.all_gather_into_tensor(out, pooled, group=group)
        return out
    padded = pooled
    if b_local < b_max:
        padded = torch.zeros((b_max, *tail), dtype=pooled.dtype, device=pooled.device)
        padded[:b_local] = pooled
    buf = torch.empty((world * b_max, *tail), dtype=pooled.dtype, device=pooled.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * b_max:r * b_max + counts[r]] for r in range(world)]
    res = torch.cat(parts, dim=0)
    if out is not None:
        out.copy_(res)
        return out
    return res


def sharded_forward(model, pixel_values_global: torch.Tensor, group: Optional["dist.ProcessGroup"] = None,
                    gather: bool = True, **forward_kwargs):
    """Run ``model`` on this rank's clips of a global batch and (optionally) gather the pooled output.

    Returns ``(local_output, pooler_output_global_or_None)``.  ``last_hidden_state`` stays local:
    only per-sample dense heads consume it.
    """
    if forward_kwargs.get("return_dict") is False:
        # the tuple form is (last_hidden_state[, hidden_states][, attentions]) — it has no pooled output
        raise ValueError("sharded_forward needs return_dict=True: the tuple output carries no pooler_output")
    local = shard_clips(pixel_values_global, group)
    out = model(local, **forward_kwargs)
    if not hasattr(out, "pooler_output") or out.pooler_output is None:
        raise ValueError("the model returned no pooler_output to gather (return_dict=False in its config?)")
    pooled = out.pooler_output
    gathered = gather_pooler_output(pooled, pixel_values_global.shape[0], group) if gather else None
    return out, gathered
