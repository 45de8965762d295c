"""Batch-sharded data parallelism for the masked-conv path.

The reference's only multi-GPU mechanism is single-process nn.DataParallel (it re-broadcasts
every parameter each forward and reduces gradients to GPU 0, CPG_cifar100_main_normal.py:199).
Here: one process per GPU, weights/masks resident and identical on every rank, each rank runs
its batch shard, and the gradients are averaged with NCCL all-reduce over NVLink.  Masks are
rank-invariant, so masking commutes with the reduction: it does not matter whether the fused
wgrad epilogue (weight decay + grad mask) ran before the all-reduce -- avg_r((g_r*b + wd*W)[T==cur])
== (avg_r(g_r)*b + wd*W)[T==cur].

Large gradients (the sharable layers: 134 MB for VGG16) are all-reduced asynchronously from
post-accumulate-grad hooks, i.e. while the rest of the backward pass is still running (FC2's
67 MB bucket goes first); small ones (BN, biases, heads) are flattened into one bucket at the
end.  Prune steps need no collective: every rank holds the same W and T.
"""
import torch
import torch.distributed as dist

from .functional import pending_side_stream

BIG = 1 << 18   # elements; tensors at least this large get their own overlapped all-reduce


def tune_env(world):
    """Environment defaults for the overlapped all-reduce; call before ``init_process_group`` and before
    the first cpg_b200 kernel.  The all-reduce of the 134 MB gradient runs next to dgrad / wgrad, whose
    grids are planned as exact waves of the SM count: NCCL's default CTA count evicts enough of those
    CTAs to cost more than the collective itself.  Measured on 2 x B200 (bench.py, batch 128 per GPU):
    default 2.33 ms/step; NCCL_MAX_CTAS=16 + grids planned for 8 fewer SMs 2.25 ms; 8 CTAs or fewer make
    the collective the critical path (2.53 ms, 3.31 ms at 4).  Existing settings win."""
    import os
    if world > 1:
        os.environ.setdefault('NCCL_MAX_CTAS', '16')
        os.environ.setdefault('CPGB_SM_MARGIN', '8')


def shard_batch(batch, rank, world):
    """rank r gets X[r*B/G:(r+1)*B/G] (SURVEY 8e 'Partitioning')."""
    n = batch.shape[0]
    if n % world != 0:
        raise ValueError(f'global batch {n} is not divisible by world size {world}')
    per = n // world
    return batch[rank * per:(rank + 1) * per]


class GradAllReducer:
    def __init__(self, model, world=None, overlap=True, group=None):
        self.world = world if world is not None else dist.get_world_size(group)
        self.group = group
        # NCCL averages inside the collective; gloo (CPU tests) only sums
        self.avg = self.world > 1 and dist.get_backend(group) == 'nccl'
        self.op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.pending = []
        self.overlap = overlap
        self._hooks = []
        if overlap and self.world > 1:
            for p in self.params:
                if p.numel() >= BIG:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._hook))

    def _hook(self, p):
        if p.grad is None:
            return
        # the weight gradient may still be in flight on the wgrad side stream (cpg_b200.functional defers the
        # join to the end of the backward pass): order the collective after that stream, not after the main one
        side = pending_side_stream(p.grad.device) if p.grad.is_cuda else None
        if side is not None:
            with torch.cuda.stream(side):
                work = dist.all_reduce(p.grad, op=self.op, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(p.grad, op=self.op, group=self.group, async_op=True)
        self.pending.append((p.grad, work))

    def reduce(self):
        """Call after backward(): waits for the overlapped reductions, reduces everything else
        in one flat bucket and divides by the world size (gradient of the global-batch mean)."""
        if self.world <= 1:
            return
        done = set()
        for g, work in self.pending:
            work.wait()
            if not self.avg:
                g.div_(self.world)
            done.add(g.data_ptr())
        self.pending = []
        rest = [p.grad for p in self.params if p.grad is not None and p.grad.data_ptr() not in done]
        if rest:
            # one flat bucket for the ~75 small tensors (BN, biases, heads, the first conv layers); packed
            # and unpacked with multi-tensor kernels instead of one launch per tensor
            flat = torch.cat([g.reshape(-1) for g in rest])
            dist.all_reduce(flat, op=self.op, group=self.group)
            if not self.avg:
                flat.div_(self.world)
            views, off = [], 0
            for g in rest:
                n = g.numel()
                views.append(flat[off:off + n].view_as(g))
                off += n
            torch._foreach_copy_(rest, views)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def assert_masks_identical(masks, group=None):
    """Debug check of SURVEY 8e: T must be bit-identical on all ranks after every prune event."""
    for name in sorted(masks):
        m = masks[name]
        s = m.to(torch.int64).sum() * 31 + (m.to(torch.int64) * torch.arange(m.numel(), device=m.device)
                                            .reshape(m.shape) % 65521).sum()
        lo, hi = s.clone(), s.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        if int(lo) != int(hi):
            raise RuntimeError(f'task mask {name} differs across ranks')
