"""Host logic of the batch-sharded DP path on CPU with the gloo backend, world_size 2."""
import io
import os
import socket
import sys
from contextlib import redirect_stdout

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tensorized_rnn_b200.dist import allreduce_gradients, shard_batch, shard_bounds  # noqa: E402


def test_shard_bounds_cover_batch_exactly():
    for batch in (1, 2, 7, 640, 1024, 16384):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(16384, 8, 3) == (6144, 8192)        # cfg4: 2048 utterances per GPU
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tensorized_rnn_b200 as tr
    torch.manual_seed(3)                         # identical replicas on every rank
    with redirect_stdout(io.StringIO()):
        m = tr.TTGRU(12, 24, 2, torch.device("cpu"), n_cores=2, tt_rank=2)
    params = list(m.parameters())
    # stand-in gradients (no CPU compute path exists): rank-dependent, deterministic
    for i, p in enumerate(params):
        g = torch.Generator().manual_seed(100 * rank + i)
        p.grad = torch.randn(p.shape, generator=g)
    params[1].grad = None                       # a parameter without gradient contributes zeros
    n = allreduce_gradients(params)
    x = torch.arange(10 * 3 * 2, dtype=torch.float32).view(10, 3, 2)
    shard = shard_batch(x)
    torch.save({"n": n, "grads": [p.grad.clone() for p in params], "shard": shard.clone()},
               os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_single_flat_allreduce_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    import tensorized_rnn_b200 as tr
    torch.manual_seed(3)
    with redirect_stdout(io.StringIO()):
        m = tr.TTGRU(12, 24, 2, torch.device("cpu"), n_cores=2, tt_rank=2)
    params = list(m.parameters())
    assert res[0]["n"] == sum(p.numel() for p in params) == m.param_count()
    for i, p in enumerate(params):
        want = torch.zeros(p.shape)
        for r in range(world):
            if i == 1:
                continue
            want += torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * r + i))
        for r in range(world):
            assert torch.allclose(res[r]["grads"][i], want, atol=1e-6), (i, r)
    full = torch.arange(10 * 3 * 2, dtype=torch.float32).view(10, 3, 2)
    assert torch.equal(torch.cat([res[0]["shard"], res[1]["shard"]]), full)


def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensorized_rnn_b200.ge2e import _AllGatherRows
    # each rank owns 3 speakers x 2 utterances x 4 features of a global batch of 6 speakers
    x = (torch.arange(3 * 2 * 4, dtype=torch.float32).view(3, 2, 4) + 100 * rank).requires_grad_(True)
    full = _AllGatherRows.apply(x, None)
    weight = torch.arange(full.numel(), dtype=torch.float32).view_as(full)
    (full * weight).sum().backward()             # every rank evaluates the same global loss
    torch.save({"full": full.detach().clone(), "grad": x.grad.clone()}, os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_ge2e_embedding_all_gather_world2(tmp_path):
    """GE2E couples every speaker of the global batch: embeddings are all-gathered, every rank evaluates the global loss
    and keeps the gradient rows of its own speakers (tensorized_rnn_b200/ge2e.py)."""
    world = 2
    port = _free_port()
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "g%d.pt" % r)) for r in range(world)]
    base = torch.arange(3 * 2 * 4, dtype=torch.float32).view(3, 2, 4)
    want_full = torch.cat([base, base + 100])
    weight = torch.arange(want_full.numel(), dtype=torch.float32).view_as(want_full)
    for r in range(world):
        assert torch.equal(res[r]["full"], want_full)
        assert torch.equal(res[r]["grad"], weight[3 * r:3 * r + 3])
