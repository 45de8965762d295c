"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding + gradient
all-reduce give the full-batch-mean gradient on every rank (SURVEY 8e parity check), with and
without the overlapped hooks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, overlap, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cpg_b200 import ddp
    ddp.BIG = 64
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 4))
    x = torch.randn(8, 16)
    y = torch.randn(8, 4)
    # full-batch reference gradient
    ref = [p.detach().clone() for p in model.parameters()]
    loss = ((model(x) - y) ** 2).mean()
    loss.backward()
    full = [p.grad.clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    red = ddp.GradAllReducer(model, world, overlap=overlap)
    xs, ys = ddp.shard_batch(x, rank, world), ddp.shard_batch(y, rank, world)
    ((model(xs) - ys) ** 2).mean().backward()
    red.reduce()
    err = max((p.grad - f).abs().max().item() for p, f in zip(model.parameters(), full))
    masks = {'a': torch.ones(4, 4, dtype=torch.uint8)}
    ddp.assert_masks_identical(masks)
    bad = False
    try:
        ddp.assert_masks_identical({'a': torch.full((4, 4), rank, dtype=torch.uint8)})
    except RuntimeError:
        bad = True
    ret[rank] = (err, bad)
    dist.destroy_process_group()


def _run(overlap, port):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, overlap, ret), nprocs=2, join=True)
    for r in range(2):
        err, bad = ret[r]
        assert err < 1e-6, err
        assert bad


def test_ddp_overlap_gloo():
    _run(True, 29611)


def test_ddp_plain_gloo():
    _run(False, 29612)


def test_shard_batch_errors():
    from cpg_b200.ddp import shard_batch
    import pytest
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(5, 3), 0, 2)
    assert shard_batch(torch.arange(8).reshape(8, 1), 1, 2).flatten().tolist() == [4, 5, 6, 7]
