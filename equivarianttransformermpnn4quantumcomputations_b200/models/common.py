"""Pieces shared by the model wrappers."""
import math

import torch
import torch.nn as nn

from .. import ops
from ..EquiformerV2Functions.module_list import ModuleListInfo
from ..EquiformerV2Functions.so3 import SO3_Grid, SO3_LinearV2


class GaussianSmearing(nn.Module):
    """exp(coeff (d - mu_k)^2), mu = linspace(start, stop, n), coeff = -0.5 / (width * dmu)^2
    (reference equiformerv2_oc20.py:43-60; fairchem escaip smearing as used by equiformerv2_qm9.py:261-266).
    Buffer key `offset` as in the reference; evaluated by `eqv2_rbf_fwd/bwd`."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50, basis_width_scalar=1.0):
        super().__init__()
        self.start, self.stop, self.num_gaussians = start, stop, num_gaussians
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (basis_width_scalar * (offset[1] - offset[0])).item() ** 2
        self.register_buffer("offset", offset)

    @property
    def num_output(self):
        return self.num_gaussians

    def forward(self, dist):
        delta = (self.stop - self.start) / (self.num_gaussians - 1) if self.num_gaussians > 1 else None
        return ops.rbf(dist.reshape(-1), self.offset, self.coeff, start=self.start, delta=delta)


def build_so3_grid(lmax_list, grid_resolution):
    """SO3_grid[l][m] module table exactly as the reference registers it (equiformerv2_oc20.py:160-166)."""
    top = max(lmax_list)
    grid = ModuleListInfo("({}, {})".format(top, top))
    for l in range(top + 1):
        row = nn.ModuleList()
        for m in range(top + 1):
            row.append(SO3_Grid(l, m, resolution=grid_resolution, normalization="component"))
        grid.append(row)
    return grid


def init_linear(m, mode):
    """`_init_weights` of the reference wrappers: zero bias; weight ~ U/N(0, 1/sqrt(fan_in))."""
    if isinstance(m, (nn.Linear, SO3_LinearV2)):
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
        std = 1 / math.sqrt(m.in_features)
        if mode == "uniform":
            nn.init.uniform_(m.weight, -std, std)
        elif mode == "normal":
            nn.init.normal_(m.weight, 0, std)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


def segment_sum(values, batch, num_graphs):
    """Per-graph sum of per-atom scalars (reference: index_add_ over `batch`, equiformerv2_oc20.py:278-281)."""
    return ops.segment_sum_nodes(values, batch, num_graphs)
