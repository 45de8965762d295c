"""Bookkeeping of the deferred wgrad-stream join (cpg_b200.functional._defer / join_side_stream) with
stand-ins for the CUDA pieces: one engine callback per backward pass, tensors released at the join, a
stale entry of a backward pass that died before its callback is joined by the next pass."""
import torch

import cpg_b200.functional as Fn


class _FakeStream:
    def __init__(self):
        self.waited = 0

    def wait_stream(self, other):
        self.waited += 1


class _NullCtx:
    def __init__(self, *a):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def test_defer_bookkeeping(monkeypatch):
    cur = _FakeStream()
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: cur)
    monkeypatch.setattr(torch.cuda, 'device', _NullCtx)
    monkeypatch.setattr(Fn, '_dev_index', lambda d: 0)
    monkeypatch.setitem(Fn._SIDE_STREAMS, 0, _FakeStream())
    Fn._PENDING.clear()
    held = []

    class Node(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a):
            return a * 2

        @staticmethod
        def backward(ctx, g):
            Fn._defer('cuda:0', (g, None))
            held.append(len(Fn._PENDING[0][1]))
            assert Fn.pending_side_stream('cuda:0') is Fn._SIDE_STREAMS[0]
            return g * 2

    x = torch.ones(2, requires_grad=True)
    Node.apply(Node.apply(x)).sum().backward()
    assert held == [1, 2]                      # both nodes of the pass share one entry (None is dropped)
    assert not Fn._PENDING and cur.waited == 1 # one join, by the engine callback
    assert Fn.pending_side_stream('cuda:0') is None
    Fn._PENDING[0] = (12345, [torch.ones(1)])  # a pass that never reached its callback
    Node.apply(x).sum().backward()
    assert not Fn._PENDING and cur.waited == 3 # stale entry joined first, then this pass's own join
    Fn.join_side_stream()                      # idempotent
    assert cur.waited == 3
