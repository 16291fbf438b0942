"""Data-parallel plumbing: structures are independent graphs, so ranks own whole structures
and the only exchange is the gradient all-reduce (reference: DistributedSampler + DDP,
train_oc20v2_parallel.py:335-346,431-436).

  shard_structures : edge-balanced assignment of structures to ranks (the reference's plain sampler
                     balances counts; attention cost is proportional to edges).
  split_batch      : the sub-batch dict of one rank (reference collate schema).
  GradientAllReducer : bucketed flat all-reduce of `.grad` after backward for models DDP cannot wrap
                     (the GATA family leaves some parameters without gradient); grads that
                     are None are sent as zeros so every rank issues the same collectives.
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
    """Average gradients over the process group in flat buckets of ~`bucket_mb` MB."""

    def __init__(self, parameters, bucket_mb=25, group=None):
        self.params = [p for p in parameters if p.requires_grad]
        self.group = group
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

    def reduce(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        pending = []
        for bucket in self.buckets:
            flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
            pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat, bucket))
        for work, flat, bucket in pending:
            work.wait()
            flat.div_(world)
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is not None:
                    p.grad.copy_(flat[off:off + n].view_as(p))
                off += n
