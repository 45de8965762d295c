"""Batch-sharded data parallelism for the masked-conv path.

The reference's only multi-GPU mechanism is single-process nn.DataParallel (it re-broadcasts
every parameter each forward and reduces gradients to GPU 0, CPG_cifar100_main_normal.py:199).
Here: one process per GPU, weights/masks resident and identical on every rank, each rank runs
its batch shard, and the gradients are averaged with NCCL all-reduce over NVLink.  Masks are
rank-invariant, so masking commutes with the reduction: it does not matter whether the fused
wgrad epilogue (weight decay + grad mask) ran before the all-reduce -- avg_r((g_r*b + wd*W)[T==cur])
== (avg_r(g_r)*b + wd*W)[T==cur].

Gradient buckets.  The weight gradients of the sharable layers (134 MB for VGG16) are written by the wgrad
epilogues STRAIGHT INTO a few large, pre-allocated flat buffers, laid out in backward order (the last layer
first): ``.weight.grad`` becomes a view of its slot, so nothing is concatenated, copied or issued per parameter.
A bucket's all-reduce is launched the moment the wgrad of its last layer has been issued, ordered after the
wgrad side stream, and overlaps the rest of the backward pass (FC2's 67 MB bucket goes first).  With a piggymask
(task >= 2) the epilogue writes the MERGED gradient m = dW + dP -- disjoint supports after the pruner's masking --
so one buffer of n floats travels instead of two; ``reduce()`` splits it back into ``.weight.grad`` /
``.piggymask.grad`` (cpgb_split_merged_grad).  Everything else (BN, biases, heads: < 1 MB) rides in one small flat
bucket at the end.  Prune steps need no collective: every rank holds the same W and T.
"""
import torch
import torch.distributed as dist

from . import _lib
from .functional import join_epilogue_stream, pending_side_stream, _side_stream

BIG = 1 << 18            # elements; other (non-sharable) tensors at least this large get their own overlapped all-reduce
# a bucket is closed once it holds this much (CPGB_BUCKET_MB overrides: smaller buckets shorten the exposed tail of the
# last all-reduce, larger ones amortise launch latency)
BUCKET_BYTES = int(float(__import__('os').environ.get('CPGB_BUCKET_MB', '48')) * (1 << 20))


def tune_env(world):
    """Environment defaults for the overlapped all-reduce; call before ``init_process_group`` and before
    the first cpg_b200 kernel.  The all-reduce of the 134 MB gradient runs next to dgrad / wgrad, whose
    grids are planned as exact waves of the SM count: NCCL's default CTA count evicts enough of those
    CTAs to cost more than the collective itself, too few CTAs make the collective the critical path.  Measured
    with bench.py (batch 128 per GPU, ms per step): 2 x B200 -- default 2.33, NCCL_MAX_CTAS=16 + grids planned for 8
    fewer SMs 2.25, 8 CTAs 2.53, 4 CTAs 3.31 (round 1); 4 x B200 -- 16 CTAs 1.753, 32 CTAs 1.720; 8 x B200 -- 16 CTAs
    1.975, 32 CTAs 1.780 (round 2: the ring / NVLS schedule has more steps per byte, the CTA budget has to grow with
    it).  Hence 16 CTAs up to two ranks, 32 beyond.  Existing settings win."""
    import os
    if world > 1:
        os.environ.setdefault('NCCL_MAX_CTAS', '16' if world <= 2 else '32')
        os.environ.setdefault('CPGB_SM_MARGIN', '8')


def shard_batch(batch, rank, world):
    """rank r gets X[r*B/G:(r+1)*B/G] (SURVEY 8e 'Partitioning')."""
    n = batch.shape[0]
    if n % world != 0:
        raise ValueError(f'global batch {n} is not divisible by world size {world}')
    per = n // world
    return batch[rank * per:(rank + 1) * per]


class GradSlot:
    """Where a sharable layer's wgrad epilogue writes: `n` floats at `offset` of bucket `bucket`."""
    __slots__ = ('reducer', 'bucket', 'offset', 'n', 'state', 'fuse')

    def __init__(self, reducer, bucket, offset, n):
        self.reducer, self.bucket, self.offset, self.n = reducer, bucket, offset, n
        self.state = 0          # 0: unused this step, 1: holds dW, 2: holds the merged dW + dP
        self.fuse = None

    def view(self, like):
        return self.reducer.flat[self.bucket][self.offset:self.offset + self.n].view(like.shape)


class GradAllReducer:
    def __init__(self, model, world=None, overlap=True, group=None, bucket_bytes=None, slots=True):
        self.world = world if world is not None else dist.get_world_size(group)
        self.group = group
        # NCCL averages inside the collective; gloo (CPU tests) only sums
        self.avg = self.world > 1 and dist.get_backend(group) == 'nccl'
        self.op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.pending = []
        self.overlap = overlap
        self._hooks = []
        self.flat, self._bucket_layers, self._bucket_left, self._slots = [], [], [], []
        self._slot_params = set()
        # slots=False: no flat buckets -- every gradient stays the tensor autograd produced (large ones get their own
        # overlapped all-reduce, the rest one flat bucket).  For models that apply one sharable layer at several call
        # sites of a forward pass: two wgrad epilogues must not write the same slot.
        if self.world > 1 and slots:
            self._make_slots(model, bucket_bytes or BUCKET_BYTES)
        if overlap and self.world > 1:
            for p in self.params:
                if p.numel() >= BIG and id(p) not in self._slot_params:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._hook))

    # ------------------------------------------------------------------ buckets of the sharable layers
    def _make_slots(self, model, bucket_bytes):
        from . import layers as nl
        mods = [m for m in model.modules() if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))
                and m.weight.is_cuda and m.weight.requires_grad and m.weight.dtype == torch.float32]
        mods.reverse()                                   # backward order: the last layer's gradient is ready first
        buckets, cur, cur_elems = [], [], 0
        for m in mods:
            cur.append(m)
            cur_elems += (m.weight.numel() + 63) // 64 * 64          # 256-byte aligned slots
            if cur_elems * 4 >= bucket_bytes:
                buckets.append(cur)
                cur, cur_elems = [], 0
        if cur:
            buckets.append(cur)
        for b, layer_list in enumerate(buckets):
            total = sum((m.weight.numel() + 63) // 64 * 64 for m in layer_list)
            dev = layer_list[0].weight.device
            self.flat.append(torch.zeros(total, dtype=torch.float32, device=dev))
            off = 0
            for m in layer_list:
                slot = GradSlot(self, b, off, m.weight.numel())
                m._cpg_grad_slot = slot
                self._slots.append((m, slot))
                self._slot_params.add(id(m.weight))
                if m.piggymask is not None:
                    self._slot_params.add(id(m.piggymask))
                off += (m.weight.numel() + 63) // 64 * 64
            self._bucket_layers.append(layer_list)
            self._bucket_left.append(len(layer_list))

    def layer_done(self, slot, device):
        """Called by the backward pass right after the wgrad of a slotted layer was issued (on the side stream
        when deferred).  Launches the bucket's all-reduce once every layer of the bucket has reported."""
        b = slot.bucket
        if self._bucket_left[b] > 0:
            self._bucket_left[b] -= 1
            if self._bucket_left[b] == 0 and self.overlap:
                self._launch_bucket(b, device)

    def _launch_bucket(self, b, device):
        if self._bucket_left[b] < 0:
            return
        self._bucket_left[b] = -1                        # launched
        buf = self.flat[b]
        with torch.cuda.device(device):
            side = _side_stream(device)
            # after every wgrad issued so far: those on the side stream by stream order, those that joined the main
            # stream immediately through this wait
            side.wait_stream(torch.cuda.current_stream())
            join_epilogue_stream(device, side)           # ... and after their epilogues (third stream)
            with torch.cuda.stream(side):
                work = dist.all_reduce(buf, op=self.op, group=self.group, async_op=True)
        self.pending.append((buf, work))

    # ------------------------------------------------------------------ other large parameters
    def _hook(self, p):
        if p.grad is None:
            return
        # the weight gradient may still be in flight on the wgrad side stream (cpg_b200.functional defers the
        # join to the end of the backward pass): order the collective after that stream, not after the main one
        side = pending_side_stream(p.grad.device) if p.grad.is_cuda else None
        if side is not None:
            # ... and after the main stream as well: this gradient may have been produced there (a stock
            # nn.Linear / nn.Embedding parameter, or a layer whose wgrad joined immediately) while an earlier
            # layer's wgrad is still deferred
            side.wait_stream(torch.cuda.current_stream(p.grad.device))
            join_epilogue_stream(p.grad.device, side)
            with torch.cuda.stream(side):
                work = dist.all_reduce(p.grad, op=self.op, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(p.grad, op=self.op, group=self.group, async_op=True)
        self.pending.append((p.grad, work))

    # ------------------------------------------------------------------ end of the backward pass
    def reduce(self):
        """Call after backward(): launches what is still unlaunched, waits for the overlapped reductions, splits the
        merged buffers, reduces everything else in one flat bucket and divides by the world size (gradient of the
        global-batch mean)."""
        if self.world <= 1:
            return
        done = set()
        used = [(m, s) for m, s in self._slots if s.state != 0]
        if used:
            dev = used[0][0].weight.device
            for b in range(len(self.flat)):
                if self._bucket_left[b] != -1 and any(s.bucket == b for _, s in used):
                    self._bucket_left[b] = 0
                    self._launch_bucket(b, dev)
        for g, work in self.pending:
            work.wait()
            if not self.avg:
                g.div_(self.world)
            done.add(g.data_ptr())
        self.pending = []
        lib = _lib.load() if used else None
        for m, s in used:
            view = s.view(m.weight)
            done.add(view.data_ptr())
            gW = m.weight.grad
            if s.state == 2:
                # m = dW + dP back into the pair the optimizers see: dW in place in its slot, dP into the tensor the
                # backward pass handed to autograd as .piggymask.grad (left unwritten there)
                gP = m.piggymask.grad if m.piggymask is not None else None
                with torch.cuda.device(view.device):
                    _lib.check(lib.cpgb_split_merged_grad(_lib.ptr(view), _lib.ptr(s.fuse.tmask), s.n, s.fuse.cur,
                                                          _lib.ptr(view), _lib.ptr(gP), _lib.stream_ptr()),
                               'cpgb_split_merged_grad')
                if gP is not None:
                    done.add(gP.data_ptr())
            if gW is not None and gW.data_ptr() != view.data_ptr():
                gW.copy_(view)           # autograd cloned instead of adopting the slot view (someone held a reference)
                done.add(gW.data_ptr())
            s.state, s.fuse = 0, None
        for b in range(len(self.flat)):
            self._bucket_left[b] = len(self._bucket_layers[b])
        rest = [p.grad for p in self.params if p.grad is not None and p.grad.data_ptr() not in done]
        if rest:
            # one flat bucket for the ~75 small tensors (BN, biases, heads); packed and unpacked with multi-tensor
            # kernels instead of one launch per tensor
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
        for m, _ in self._slots:
            m._cpg_grad_slot = None
        self._slots = []


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
