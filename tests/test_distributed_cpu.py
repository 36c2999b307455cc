"""Host-side logic of the batch-sharded path on CPU: world_size-2 (and 3) gloo
process groups.  The encoder itself is replaced by a per-clip deterministic stand-in (the real one
needs a B200; clips are independent, which is all the sharding relies on): what is tested is shard
bounds, rank order of the gather, ragged shards and equality with one process over the whole batch."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streamformer_b200.distributed import gather_pooler_output, shard_bounds, shard_clips, sharded_forward


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_encoder(pixel_values):
    # per-clip, per-frame statistic [B, T, 4]: any cross-clip mixing or reordering would change it
    B, T = pixel_values.shape[:2]
    flat = pixel_values.reshape(B, T, -1)
    pooled = torch.stack([flat.mean(-1), flat.amax(-1), flat.amin(-1), flat.square().mean(-1)], dim=-1)
    return SimpleNamespace(pooler_output=pooled, last_hidden_state=flat)


def _worker(rank, world, port, global_batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        pixels = torch.randn(global_batch, 3, 3, 8, 8, generator=g)          # identical on every rank
        local = shard_clips(pixels)
        s, e = shard_bounds(global_batch, rank, world)
        assert local.shape[0] == e - s and torch.equal(local, pixels[s:e])
        out, gathered = sharded_forward(_fake_encoder, pixels)
        want = _fake_encoder(pixels).pooler_output
        assert gathered.shape == want.shape
        assert torch.equal(gathered, want), "gathered pooler_output differs from the one-process result"
        assert torch.equal(out.pooler_output, want[s:e])
        # pre-allocated output buffer variant (what bench.py uses)
        buf = torch.empty_like(want)
        got = gather_pooler_output(out.pooler_output, global_batch, out=buf)
        assert got.data_ptr() == buf.data_ptr() and torch.equal(buf, want)
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover - reported to the parent
        q.put((rank, f"{type(ex).__name__}: {ex}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,global_batch", [(2, 8), (2, 5), (3, 7)])
def test_sharded_forward_equals_single_process(world, global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_shard_bounds_cover_batch_exactly():
    for gb in (0, 1, 5, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_single_process_is_identity():
    x = torch.randn(4, 2, 3, 8, 8)
    assert shard_clips(x).data_ptr() == x.data_ptr()
    p = torch.randn(4, 2, 16)
    assert gather_pooler_output(p) is p
