"""world_size-2 gloo test of the multi-GPU host logic (batch sharding + single-bucket gradient all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maskflownet_b200 import dist as mdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = mdist.init_from_env("gloo")
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Conv2d(4, 2, 3, padding=1))
    bucket = mdist.GradBucket(model.parameters())
    data = torch.arange(8 * 3 * 5 * 5, dtype=torch.float32).reshape(8, 3, 5, 5) / 100.0
    b, e = mdist.shard_batch(8, r, w)
    bucket.zero_()
    loss = model(data[b:e]).square().sum()          # per-sample losses are SUMMED (reference: loss.backward() per device)
    loss.backward()
    bucket.allreduce_(global_batch=8)
    t = mdist.max_over_ranks(float(r + 1), "cpu")
    q.put((r, bucket.flat.clone(), t))
    dist.destroy_process_group()


def test_gloo_two_ranks_match_single_process():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Conv2d(4, 2, 3, padding=1))
    data = torch.arange(8 * 3 * 5 * 5, dtype=torch.float32).reshape(8, 3, 5, 5) / 100.0
    model(data).square().sum().backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()]) / 8.0
    for r, flat, t in res:
        assert torch.allclose(flat, ref, rtol=1e-5, atol=1e-6)
        assert t == 2.0
    with pytest.raises(ValueError):
        mdist.shard_batch(7, 0, 2)
    assert mdist.shard_batch(32, 3, 8) == (12, 16)


def test_grad_bucket_detects_broken_aliasing_and_async_scale():
    """zero_grad(set_to_none=True) drops the views into the bucket -- allreduce_ must refuse instead of reducing a
    stale buffer; rebind_() repairs it; the async handle applies the 1/batch scale in wait() (single process: no work)."""
    torch.manual_seed(1)
    model = torch.nn.Linear(4, 3)
    bucket = mdist.GradBucket(model.parameters())
    model(torch.ones(2, 4)).sum().backward()
    g = bucket.flat.clone()
    assert g.abs().sum() > 0
    assert bucket.allreduce_(global_batch=2, async_op=True) is None and torch.allclose(bucket.flat, g / 2)
    model.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        bucket.allreduce_(global_batch=2)
    bucket.rebind_()
    bucket.zero_()
    model(torch.ones(2, 4)).sum().backward()
    bucket.allreduce_(global_batch=1)
    assert torch.allclose(bucket.flat, g)
    h = mdist.GradBucket._Handle(None, bucket.flat, 0.5)
    h.wait()                       # nothing in flight: no-op
    assert torch.allclose(bucket.flat, g)
