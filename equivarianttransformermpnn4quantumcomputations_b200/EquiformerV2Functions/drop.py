"""Stochastic-depth / dropout layers (reference drop.py:16-149).  RNG stays in torch so that a
seeded run draws the same masks as the reference; the masks are cheap per-graph / per-node
scalars multiplied into the node tensor."""
import torch
import torch.nn as nn


def drop_path(x, drop_prob: float = 0.0, training: bool = False):
    if drop_prob == 0.0 or not training:
        return x
    keep = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    mask = (keep + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
    return x.div(keep) * mask


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)

    def extra_repr(self):
        return "drop_prob={}".format(self.drop_prob)


_NUM_GRAPHS = {"n": None}


def set_num_graphs(n):
    """Number of structures in the batch the next forward passes work on (None: derive it from `batch` as the
    reference does)."""
    _NUM_GRAPHS["n"] = None if n is None else int(n)


class GraphDropPath(nn.Module):
    """Per-graph stochastic depth (reference drop.py:49-68)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x, batch):
        # reference: batch.max() + 1 (drop.py:60) -- a device read-back on every call.  The model wrappers announce the
        # number of structures they were given (set_num_graphs), which is the same value without the synchronisation
        # (and keeps the block capturable in a CUDA graph); without the announcement the reference expression is used.
        num_graphs = _NUM_GRAPHS["n"] if _NUM_GRAPHS["n"] is not None else int(batch.max()) + 1
        if self.training and self.drop_prob and x.is_cuda and x.dtype == torch.float32 and batch.dtype == torch.long:
            # same draw as the reference (`torch.rand((num_graphs, 1, 1))` inside drop_path), then ONE kernel for
            # ones.div(keep) * floor(keep + u), the gather by `batch` and the product (bit-identical)
            from .. import ops
            u = torch.rand((num_graphs,) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)
            return ops.drop_path_scale(x, u, batch, 1.0 - self.drop_prob)
        ones = torch.ones((num_graphs,) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)
        return x * drop_path(ones, self.drop_prob, self.training)[batch]

    def extra_repr(self):
        return "drop_prob={}".format(self.drop_prob)


class EquivariantDropoutArraySphericalHarmonics(nn.Module):
    """Channel dropout shared by all (l, m) of a node (reference drop.py:119-149)."""

    def __init__(self, drop_prob, drop_graph=False):
        super().__init__()
        self.drop_prob = drop_prob
        self.drop = nn.Dropout(drop_prob, True)
        self.drop_graph = drop_graph

    def forward(self, x, batch=None):
        if not self.training or self.drop_prob == 0.0:
            return x
        assert x.dim() == 3
        if self.drop_graph:
            assert batch is not None
            mask = self.drop(torch.ones((batch.max() + 1, 1, x.shape[2]), dtype=x.dtype, device=x.device))
            return x * mask[batch]
        mask = self.drop(torch.ones((x.shape[0], 1, x.shape[2]), dtype=x.dtype, device=x.device))
        return x * mask

    def extra_repr(self):
        return "drop_prob={}, drop_graph={}".format(self.drop_prob, self.drop_graph)
