"""Data-parallel plumbing: structures are independent graphs, so ranks own whole structures
and the only exchange is the gradient all-reduce (reference: DistributedSampler + DDP,
train_oc20v2_parallel.py:335-346,431-436).

  shard_structures : edge-balanced assignment of structures to ranks (the reference's plain sampler
                     balances counts; attention cost is proportional to edges).
  split_batch      : the sub-batch dict of one rank (reference collate schema).
  GradientAllReducer : bucketed flat all-reduce of `.grad` after backward for models DDP cannot wrap
                     (the GATA family leaves some parameters without gradient); grads that
                     are None are sent as zeros so every rank issues the same collectives.
  OverlappedGradientAllReducer : the same exchange issued bucket by bucket FROM INSIDE the backward pass on a side
                     stream (what DDP's reducer does, train_oc20v2_parallel.py:431-436), capturable into the CUDA graph
                     of the step; the gradients then LIVE in the flat buckets (no copy back).
"""
import torch
import torch.distributed as dist


def shard_structures(natoms, world_size, max_neighbors=20):
    """Greedy longest-processing-time partition by estimated edge count n_atoms * min(max_nb, n_atoms - 1).
    Returns a list (one entry per rank) of sorted structure indices; every structure appears exactly once."""
    natoms = [int(n) for n in natoms]
    cost = [n * max(1, min(max_neighbors, n - 1)) for n in natoms]
    order = sorted(range(len(natoms)), key=lambda i: (-cost[i], i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += cost[i]
    return [sorted(s) for s in out]


def split_batch(data, structures):
    """Sub-batch with the given structures (in that order); per-atom and per-structure tensors are sliced,
    `batch` is renumbered from 0."""
    natoms = data["natoms"]
    B = int(natoms.shape[0])
    starts = torch.zeros(B + 1, dtype=torch.long)
    starts[1:] = torch.cumsum(natoms.cpu(), 0)
    atom_idx = torch.cat([torch.arange(int(starts[s]), int(starts[s + 1])) for s in structures]) if structures \
        else torch.zeros(0, dtype=torch.long)
    sidx = torch.tensor(structures, dtype=torch.long)
    n_total = int(starts[-1])
    out = {}
    for k, v in data.items():
        if not torch.is_tensor(v):
            out[k] = v
        elif k == "batch":
            out[k] = torch.repeat_interleave(torch.arange(len(structures)), natoms.cpu()[sidx]).to(v.device)
        elif v.shape[:1] == (n_total,) and k != "natoms":
            out[k] = v[atom_idx.to(v.device)]
        elif v.shape[:1] == (B,):
            out[k] = v[sidx.to(v.device)]
        else:
            out[k] = v
    return out


class GradientAllReducer:
    """Average gradients over the process group in persistent flat buckets of ~`bucket_mb` MB.

    Per bucket: ONE multi-tensor copy of the gradients into the flat buffer, one all-reduce (AVG on NCCL; SUM + scale
    on backends without AVG), one multi-tensor copy back -- a handful of launches per step instead of one per parameter
    (82.5 M parameters live in ~800 tensors; per-tensor copies cost the host ~4 ms per step)."""

    def __init__(self, parameters, bucket_mb=25, group=None, sync_presence=False):
        """sync_presence: also exchange which parameters have a gradient on ANY rank and give the ranks that lack one
        the averaged gradient, so that every replica applies the same update when gradients are present on some ranks
        only (empty sub-batch, a branch not exercised on a rank).  Costs one small all-reduce and one host read-back
        per step, so it is off by default: the reference models either give every parameter a gradient on every rank
        or leave the same parameters without one everywhere (GATA family)."""
        self.params = [p for p in parameters if p.requires_grad]
        self.group = group
        self.sync_presence = sync_presence
        self.buckets, cur, size = [], [], 0
        limit = bucket_mb * (1 << 20)
        for p in self.params:
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= limit:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat = {}          # bucket index -> (flat buffer, per-parameter views)

    def _buffers(self, i, bucket):
        hit = self._flat.get(i)
        if hit is None or hit[0].device != bucket[0].device:
            flat = torch.empty(sum(p.numel() for p in bucket), dtype=bucket[0].dtype, device=bucket[0].device)
            views, off = [], 0
            for p in bucket:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            hit = self._flat[i] = (flat, views)
        return hit

    def reduce(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        avg = dist.get_backend(self.group) == "nccl"
        pending = []
        for i, bucket in enumerate(self.buckets):
            flat, views = self._buffers(i, bucket)
            have = [(v, p.grad) for v, p in zip(views, bucket) if p.grad is not None]
            if len(have) < len(bucket):
                flat.zero_()                                  # parameters without gradient are sent as zeros
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
            pending.append((dist.all_reduce(flat, op=op, group=self.group, async_op=True), flat, have, i))
        # A parameter may have a gradient on some ranks only (empty sub-batch, a branch not exercised on a rank): every
        # rank must then apply the same averaged update or the replicas drift apart.  One small all-reduce of
        # the per-parameter presence flags tells each rank which gradients exist anywhere.
        anywhere = None
        if self.sync_presence:
            present = torch.tensor([0.0 if p.grad is None else 1.0 for p in self.params], device=self.params[0].device)
            dist.all_reduce(present, op=dist.ReduceOp.MAX, group=self.group)
            anywhere = present.bool().tolist()
            index = {id(p): k for k, p in enumerate(self.params)}
        for work, flat, have, i in pending:
            work.wait()
            if not avg:
                flat.div_(world)
            if have:
                torch._foreach_copy_([g for _, g in have], [v for v, _ in have])
            if anywhere is not None:
                _, views = self._flat[i]
                for v, p in zip(views, self.buckets[i]):
                    if p.grad is None and anywhere[index[id(p)]]:
                        p.grad = v.clone()


class OverlappedGradientAllReducer:
    """Gradient averaging that overlaps the backward pass, and that a CUDA-graph capture of the step can contain.

    Buckets are filled in the order gradients become ready (reverse registration order ~ backward order).  A
    post-accumulate-grad hook on every parameter counts its bucket down; when a bucket is complete the hook forks a side
    stream off the stream the backward pass runs on, copies the bucket's gradients into its persistent flat buffer (one
    multi-tensor copy) and all-reduces the buffer there, while the backward pass carries on.  `finish()` flushes the
    buckets that never completed (parameters without gradient are sent as zeros, so every rank issues the same
    collectives in the same order), joins the side stream and re-points `.grad` at the flat views: the optimizer reads
    the averaged gradients where NCCL left them -- no copy back, and the FusedAdamW pointer table stays valid for ever.

        sync = OverlappedGradientAllReducer(model.parameters())
        sync.begin(); loss.backward(); sync.finish(); opt.step()

    Inside `torch.cuda.graph` capture the fork / copy / all-reduce / join are captured with the step (NCCL collectives are
    capturable); nothing is exchanged at capture time, every replay exchanges once.  Hooks are inert outside
    begin()/finish(), so warm-up passes do not communicate (ranks may capture at different steps).
    Parameters that did not get a gradient in the previous pass (the GATA family leaves some without) are
    not waited for in the next one, so their bucket still goes out from inside the backward pass.
    Requirement (as for DDP): every rank runs the same model, so buckets complete in the same order everywhere."""

    def __init__(self, parameters, bucket_mb=32, group=None, overlap=True):
        """overlap=False: same flat-bucket exchange (captured with the step, no copy back), but every bucket is sent from
        finish(), after the backward pass -- for A/B measurements of what the concurrency costs the compute kernels (an
        NCCL kernel occupies SMs for as long as its collective runs; the persistent GEMM kernels size their grid to all
        148)."""
        self.params = [p for p in parameters if p.requires_grad]
        self.group = group
        self.overlap = overlap
        limit = bucket_mb * (1 << 20)
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= limit:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat, self.view, self.bucket_of = [], {}, {}
        for b, bucket in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in bucket), dtype=bucket[0].dtype, device=bucket[0].device)
            off = 0
            for p in bucket:
                self.view[id(p)] = flat[off:off + p.numel()].view_as(p)
                self.bucket_of[id(p)] = b
                off += p.numel()
            self.flat.append(flat)
        self.cuda = self.params[0].is_cuda
        self.side = torch.cuda.Stream(device=self.params[0].device) if self.cuda else None
        self.active = dist.is_initialized() and dist.get_world_size(group) > 1
        self.avg = self.active and dist.get_backend(group) == "nccl"
        self.world = dist.get_world_size(group) if self.active else 1
        self.armed = False
        self.silent = set()          # id(p) of the parameters that got no gradient in the previous pass
        self.launched = 0            # buckets sent since begin() (diagnostics / tests)
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        if self.active and self.cuda:      # create the communicator now: it cannot be created inside a graph capture
            dist.all_reduce(torch.zeros(1, device=self.params[0].device), group=group)
            torch.cuda.synchronize()

    def begin(self):
        """Arm the hooks: call right before the backward pass (inside the capture when the step is captured)."""
        self.left = [sum(id(p) not in self.silent for p in b) for b in self.buckets]
        self.fired = set()
        self.sent = [False] * len(self.buckets)
        self.launched = 0
        self.armed = True

    def _on_grad(self, p):
        if not self.armed:
            return
        b = self.bucket_of[id(p)]
        self.fired.add(id(p))
        if id(p) in self.silent:         # a gradient that was not there last time: its bucket goes (again) at finish()
            self.sent[b] = False
            self.left[b] = -1
            return
        self.left[b] -= 1
        if self.left[b] == 0 and self.overlap:
            self._send(b)

    def _send(self, b):
        bucket, flat = self.buckets[b], self.flat[b]
        have = [p for p in bucket if id(p) in self.fired and p.grad is not None]
        if self.cuda:
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)              # fork: everything the backward pass enqueued so far
            ctx = torch.cuda.stream(self.side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            if len(have) < len(bucket):
                flat.zero_()
            if have:
                torch._foreach_copy_([self.view[id(p)] for p in have], [p.grad for p in have])
            if self.active:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, group=self.group)
                if not self.avg:
                    flat.div_(self.world)
        self.sent[b] = True
        self.launched += 1

    def finish(self):
        """After the backward pass: send what is left, join, and let `.grad` be the averaged flat views."""
        assert self.armed, "finish() without begin()"
        self.armed = False
        for b in range(len(self.buckets)):
            if not self.sent[b]:
                self._send(b)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)
        for p in self.params:
            if id(p) in self.fired and p.grad is not None:
                p.grad = self.view[id(p)]
        self.silent = {id(p) for p in self.params if id(p) not in self.fired}

    def reduce(self):
        """Drop-in for GradientAllReducer.reduce (no overlap): exchange the gradients that exist now."""
        self.begin()
        for p in self.params:
            if p.grad is not None:
                self._on_grad(p)
        self.finish()

    def detach(self):
        for h in self.handles:
            h.remove()
        self.handles = []
