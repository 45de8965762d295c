"""cpg_b200.optim.SGD / Adam against torch.optim on the same GPU -- the optimizers the reference
constructs at CPG_cifar100_main_normal.py:339-346 -- bit for bit over several steps, including the momentum drift of
weights whose gradient is exactly zero, CUDA-graph capture, checkpoint round trips and the packed mask
words the Adam kernel emits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

SHAPES = [(64, 3, 3, 3), (128, 64, 3, 3), (4096, 512), (4096,), (7,), (1,), (5016, 627), (100, 4096), (33, 5, 3, 3)]


def _params(seed, shapes=SHAPES):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.1).to(DEV)) for s in shapes]


def _grads(step, params, zero_frac=0.5):
    g = torch.Generator().manual_seed(1000 + step)
    out = []
    for p in params:
        gr = torch.randn(*p.shape, generator=g) * 0.01
        gr[torch.rand(*p.shape, generator=g) < zero_frac] = 0.0         # the masked gradients are mostly exact zeros
        out.append(gr.to(DEV))
    return out


def _bits(t):
    return t.detach().view(torch.int32)


def test_sgd_nesterov_bit_exact_vs_torch():
    from cpg_b200.optim import SGD
    pa, pb = _params(0), _params(0)
    ours = SGD(pa, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True)
    ref = torch.optim.SGD(pb, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True)      # the reference's call
    for step in range(6):
        if step == 3:                        # the reference's schedule writes param_group['lr']
            for o in (ours, ref):
                for gr in o.param_groups:
                    gr['lr'] = 1e-3
        gs = _grads(step, pa, zero_frac=1.0 if step == 4 else 0.5)       # step 4: all-zero gradients, weights still move
        for p, q, g in zip(pa, pb, gs):
            p.grad = g.clone(); q.grad = g.clone()
        before = [p.detach().clone() for p in pa]
        ours.step(); ref.step()
        for i, (p, q) in enumerate(zip(pa, pb)):
            assert torch.equal(_bits(p), _bits(q)), (step, i, (p - q).abs().max().item())
            assert torch.equal(_bits(ours.state[p]['momentum_buffer']), _bits(ref.state[q]['momentum_buffer'])), (step, i)
        if step == 4:
            assert all(not torch.equal(b, p.detach()) for b, p in zip(before, pa) if p.numel() > 8)   # momentum drift


def test_adam_bit_exact_vs_torch_and_packed_words():
    from cpg_b200 import _lib
    from cpg_b200.optim import Adam
    lib = _lib.load()
    shapes = [(4096, 512), (64, 64, 3, 3), (100, 37), (5,)]
    g0 = torch.Generator().manual_seed(5)
    pa = [torch.nn.Parameter((torch.rand(*s, generator=g0) * 0.01).to(DEV)) for s in shapes]      # piggymasks live near 5e-3
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    tm = [torch.randint(0, 4, s, generator=g0).to(torch.uint8).to(DEV) for s in shapes]
    pack = {p: (torch.zeros((p.numel() + 31) // 32, dtype=torch.int64, device=DEV), t if i != 1 else None)
            for i, (p, t) in enumerate(zip(pa, tm))}
    ours = Adam(pa, lr=5e-4, pack=pack, pack_inference_idx=2)
    ref = torch.optim.Adam(pb, lr=5e-4)                                                       # the reference's call
    worst = 0
    for step in range(8):
        gs = _grads(step, pa, zero_frac=0.6)
        for p, q, g in zip(pa, pb, gs):
            p.grad = g.clone(); q.grad = g.clone()
        ours.step(); ref.step()
        for i, (p, q) in enumerate(zip(pa, pb)):
            d = (_bits(p).long() - _bits(q).long()).abs().max().item()
            worst = max(worst, d)
            assert torch.equal(_bits(ours.state[p]['exp_avg']), _bits(ref.state[q]['exp_avg'])), (step, i)
            assert torch.equal(_bits(ours.state[p]['exp_avg_sq']), _bits(ref.state[q]['exp_avg_sq'])), (step, i)
            assert d == 0, (step, i, d)
        # the words the kernel emitted == cpgb_pack_mask on the updated parameter
        for i, p in enumerate(pa):
            want = torch.zeros_like(pack[p][0])
            _lib.check(lib.cpgb_pack_mask(_lib.ptr(p.detach()), _lib.ptr(pack[p][1]), p.numel(), 5e-3, 2, _lib.ptr(want),
                                          _lib.stream_ptr()), 'pack')
            assert torch.equal(pack[p][0], want), (step, i)
    assert int(ours.state_dict()['state'][0]['step']) == 8 == int(ref.state_dict()['state'][0]['step'])


def test_optimizers_in_a_cuda_graph_and_checkpoint_round_trip():
    from cpg_b200.optim import SGD, Adam
    pa, pb = _params(3, SHAPES[:4]), _params(3, SHAPES[:4])
    qa, qb = _params(4, SHAPES[1:3]), _params(4, SHAPES[1:3])
    ours = [SGD(pa, lr=1e-2, momentum=0.9, nesterov=True, lr_tensor=True), Adam(qa, lr=5e-4, lr_tensor=True)]
    ref = [torch.optim.SGD(pb, lr=1e-2, momentum=0.9, nesterov=True), torch.optim.Adam(qb, lr=5e-4)]
    for p in pa + qa:
        p.grad = torch.zeros_like(p)
    for o in ours:
        o.sync_lr()

    def feed(step):
        for plist, rlist in ((pa, pb), (qa, qb)):
            for p, q, g in zip(plist, rlist, _grads(step, plist)):
                p.grad.copy_(g); q.grad = g.clone()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for step in range(2):                    # eager warm-up steps (state buffers, device counters)
            feed(step)
            for o in ours + ref:
                o.step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    feed(2)
    with torch.cuda.graph(graph):
        for o in ours:
            o.step()
    graph.replay()                               # capture does not execute: this is step 2
    for o in ref:
        o.step()
    for step in range(3, 6):
        if step == 4:                            # schedule change between replays: device copies of lr
            for o in ours + ref:
                for gr in o.param_groups:
                    gr['lr'] *= 0.1
            for o in ours:
                o.sync_lr()
        feed(step)
        graph.replay()
        for o in ref:
            o.step()
    for p, q in zip(pa + qa, pb + qb):
        assert torch.equal(_bits(p), _bits(q))
    # checkpoint round trip in torch's own format, both directions
    sd = ours[1].state_dict()
    assert int(sd['state'][0]['step']) == 6
    fresh = torch.optim.Adam(qb, lr=5e-5)
    fresh.load_state_dict(sd)                    # ours -> torch
    ours2 = Adam(qa, lr=5e-5)
    ours2.load_state_dict(ref[1].state_dict())   # torch -> ours
    feed(7)
    fresh.step(); ours2.step()
    for p, q in zip(qa, qb):
        assert torch.equal(_bits(p), _bits(q))


def test_optimizers_reject_what_they_do_not_implement():
    from cpg_b200 import _lib
    from cpg_b200.optim import SGD, Adam
    p = torch.nn.Parameter(torch.zeros(4, device=DEV))
    with pytest.raises(ValueError):
        SGD([p], lr=0.1, momentum=0.9, nesterov=False)
    with pytest.raises(ValueError):
        SGD([p], lr=0.1, momentum=0.9, weight_decay=1e-4)
    with pytest.raises(ValueError):
        Adam([p], amsgrad=True)
    c = torch.nn.Parameter(torch.zeros(4))
    c.grad = torch.zeros(4)
    with pytest.raises(_lib.CpgbError):
        SGD([c], lr=0.1).step()


def test_adam_emitted_mask_words_feed_the_in_tile_masked_layer():
    """cpg_b200.optim.Adam.emit_packed_masks: the FC layer's next forward pass takes the Binarizer bits the Adam kernel
    wrote (no pack launch), gives the same output bit for bit, and ignores them once anything else touched the
    piggymask."""
    import torch.nn as nn
    import cpg_b200.layers as nl
    from cpg_b200 import _lib
    from cpg_b200.optim import Adam
    lib = _lib.load()
    torch.manual_seed(0)
    lin = nl.SharableLinear(512, 1024).to(DEV)
    with torch.no_grad():                                  # the reference layer leaves its parameters uninitialised
        lin.weight.normal_(0, 0.05)
        lin.bias.normal_(0, 0.1)
    lin.piggymask = nn.Parameter((torch.rand(1024, 512) * 0.01).to(DEV))
    opt = Adam([lin.piggymask], lr=5e-4)
    assert opt.emit_packed_masks(lin) == 1
    x = torch.randn(128, 512, device=DEV)
    lin(x).square().mean().backward()
    opt.step()
    assert lin.piggymask._cpgb_bits is not None
    before = lib.cpgb_launch_count()
    y1 = lin(x)
    n1 = lib.cpgb_launch_count() - before
    assert lin.piggymask._cpgb_bits is None                 # consumed
    before = lib.cpgb_launch_count()
    y2 = lin(x)                                             # packs by itself
    n2 = lib.cpgb_launch_count() - before
    assert n2 == n1 + 1, (n1, n2)
    assert torch.equal(y1, y2), (y1 - y2).abs().max().item()
    # stale words: a torch op on the piggymask after the optimizer step bumps its version
    lin(x).square().mean().backward()
    opt.step()
    with torch.no_grad():
        lin.piggymask.mul_(-1.0)                            # every bit flips to 0
    y3 = lin(x)
    assert torch.equal(y3, lin.bias.detach().expand_as(y3) if lin.bias is not None else torch.zeros_like(y3))
