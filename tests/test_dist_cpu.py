"""world_size-2 gloo tests of the multi-GPU host logic: batch sharding + single-bucket gradient all-reduce, and the whole
pipeline.PipelineFlownet.train_batch step (network/pipeline.py:89-115) on two ranks against one process with the full batch."""
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
    q.put((r, bucket.flat.clone().numpy(), t))          # numpy: a tensor would travel as a shared-memory handle
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
        assert torch.allclose(torch.from_numpy(flat), ref, rtol=1e-5, atol=1e-6)
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


# ---- PipelineFlownet.train_batch on two ranks == one process with the whole batch ----------------------------------------
class _TinyNet(torch.nn.Module):
    """Stand-in with MaskFlownetS's output contract ([flow6..flow2], [mask2], None): the CUDA model cannot run here."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(6, 3, 3, padding=1)

    def forward(self, a, b):
        import torch.nn.functional as tF
        y = self.conv(torch.cat([a, b], dim=1))
        return [tF.avg_pool2d(y[:, :2], s) * 20.0 for s in (64, 32, 16, 8, 4)], [torch.sigmoid(tF.avg_pool2d(y[:, 2:3], 4))], None


def _geo_stub(i1, i2, fl, mk):
    n, _, H, W = i1.shape
    return i1.float() / 255, i2.float() / 255, fl.clone(), (mk.float() / 255).expand(n, 1, H, W).contiguous()


def _color_stub(a, b):
    return a, b


def _make_pipeline(setattr_=setattr):
    """setattr_: plain setattr in the spawned workers (their process ends with the test), monkeypatch.setattr in the parent."""
    from maskflownet_b200 import network, ops, pipeline
    from oracle import torch_ref
    setattr_(network, "MaskFlownetS", _TinyNet)
    setattr_(ops, "upsample", lambda x, f, scale=1.0: torch_ref.upsample(x, f) * scale)
    torch.manual_seed(0)
    return pipeline.PipelineFlownet(device="cpu")


def _train_data():
    g = torch.Generator().manual_seed(3)
    img1 = torch.randint(0, 256, (4, 3, 64, 128), generator=g, dtype=torch.uint8)
    img2 = torch.randint(0, 256, (4, 3, 64, 128), generator=g, dtype=torch.uint8)
    return img1, img2, torch.randn(4, 2, 64, 128, generator=g) * 2


def _pipeline_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = mdist.init_from_env("gloo")
    pipe = _make_pipeline()
    img1, img2, label = _train_data()
    b, e = mdist.shard_batch(4, r, w)
    for _ in range(2):                                    # two optimizer steps: the ranks must stay in lock-step
        out = pipe.train_batch(img1[b:e], img2[b:e], label[b:e], _geo_stub, _color_stub)
    # numpy, not a tensor: torch.multiprocessing would pass a shared-memory handle that dies with this process
    q.put((r, torch.cat([p.detach().flatten() for p in pipe.network.parameters()]).numpy(), out["epe"]))
    dist.destroy_process_group()


def test_gloo_pipeline_train_batch_two_ranks_match_single_process(monkeypatch):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pipe = _make_pipeline(monkeypatch.setattr)
    img1, img2, label = _train_data()
    for _ in range(2):
        pipe.train_batch(img1, img2, label, _geo_stub, _color_stub)
    ref = torch.cat([p.detach().flatten() for p in pipe.network.parameters()])
    for r, flat, epe in res:
        flat = torch.from_numpy(flat)
        assert torch.allclose(flat, ref, rtol=1e-5, atol=1e-6), (r, (flat - ref).abs().max())
        assert epe > 0
    assert (res[0][1] == res[1][1]).all()                 # identical replicas after the all-reduce
