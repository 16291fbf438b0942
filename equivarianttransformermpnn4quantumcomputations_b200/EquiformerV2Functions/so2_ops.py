"""SO(2) convolution modules with the reference's names, constructor arguments and parameter
keys (reference so2_ops.py: SO2_m_Convolution :11-61, SO2_Convolution :64-204).

Layout underneath: the kernels work on edge rows in m-primary order [E, Kr * C] (the reference's
`_m_primary` / `_l_primary` dense permutation einsums, so3.py:322-339, become index tables), and
all m-blocks of one convolution run as ONE grouped GEMM launch:
  * m = 0 : fc_m0 (bias, optional extra columns in front),
  * m > 0 : the (+m, -m) pair [x+ | x-] (2k) -> [out+ | out-] (2o) through the real 2x2 block form
            [[Wr, -Wi], [Wi, Wr]] of the complex multiply at so2_ops.py:53-61 -- the same FLOPs as the
            reference's [E,2,k] x [k,2o] product, with the +/- recombination folded into the GEMM.
"""
import copy
import math

import torch
import torch.nn as nn
from torch.nn import Linear

from .. import ops
from .radial_function import RadialFunction
from .so3 import SO3_Embedding


class SO2_m_Convolution(nn.Module):
    """Parameter holder for order m (`fc.weight` [2*o*n, n*c], no bias, scaled by 1/sqrt(2) at init)."""

    def __init__(self, m, sphere_channels, m_output_channels, lmax_list, mmax_list):
        super().__init__()
        self.m = m
        self.sphere_channels = sphere_channels
        self.m_output_channels = m_output_channels
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list
        self.num_resolutions = len(lmax_list)
        ncoef = sum(max(0, l - m + 1) for l, mm in zip(lmax_list, mmax_list) if mm >= m)
        assert ncoef > 0
        self.fc = Linear(ncoef * sphere_channels, 2 * m_output_channels * ncoef, bias=False)
        self.fc.weight.data.mul_(1 / math.sqrt(2))

    def block_weight(self):
        """[[Wr, -Wi], [Wi, Wr]]  ([2o', 2k]) -- differentiable w.r.t. fc.weight."""
        return ops.so2_block_weight(self.fc.weight)

    def forward(self, x_m):
        """API parity with the reference: x_m [E, 2, k] -> [E, 2, o']."""
        E = x_m.shape[0]
        Wb = self.block_weight()
        k2, o2 = Wb.shape[1], Wb.shape[0]
        y = ops.so2_conv(x_m.reshape(E, k2), None, ((0, k2, 0, o2),), [Wb])
        return y.view(E, 2, o2 // 2)


class SO2_Convolution(nn.Module):
    def __init__(self, sphere_channels, m_output_channels, lmax_list, mmax_list, mappingReduced,
                 internal_weights=True, edge_channels_list=None, extra_m0_output_channels=None):
        super().__init__()
        self.sphere_channels = sphere_channels
        self.m_output_channels = m_output_channels
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list
        self.mappingReduced = mappingReduced
        self.num_resolutions = len(lmax_list)
        if self.num_resolutions != 1:
            raise NotImplementedError("SO2_Convolution: a single (lmax, mmax) resolution is supported "
                                      "(every reference config uses one)")
        self.internal_weights = internal_weights
        self.edge_channels_list = copy.deepcopy(edge_channels_list)
        self.extra_m0_output_channels = extra_m0_output_channels

        n_m0 = (lmax_list[0] + 1) * sphere_channels
        out_m0 = m_output_channels * (lmax_list[0] + 1) + (extra_m0_output_channels or 0)
        self.fc_m0 = Linear(n_m0, out_m0)
        num_rad = n_m0
        self.so2_m_conv = nn.ModuleList()
        for m in range(1, max(mmax_list) + 1):
            self.so2_m_conv.append(SO2_m_Convolution(m, sphere_channels, m_output_channels, lmax_list, mmax_list))
            num_rad += self.so2_m_conv[-1].fc.in_features
        self.rad_func = None
        if not internal_weights:
            assert self.edge_channels_list is not None
            self.edge_channels_list.append(int(num_rad))
            self.rad_func = RadialFunction(self.edge_channels_list)

    # -- fused path -------------------------------------------------------------------------
    def layout(self):
        return ops.CoeffLayout.get(self.lmax_list[0], self.mmax_list[0])

    def radial_weights(self, x_edge):
        """[E, n_rad] per-edge modulation (so2_ops.py:145-146) or None."""
        return self.rad_func(x_edge) if self.rad_func is not None else None

    def groups_and_weights(self):
        """(column groups, GEMM weights) of the m = 0..mmax blocks in m-primary order (ops.so2_conv)."""
        groups = tuple(self.layout().conv_groups(self.sphere_channels, self.m_output_channels,
                                                 self.extra_m0_output_channels or 0))
        return groups, [self.fc_m0.weight] + [mc.block_weight() for mc in self.so2_m_conv]

    def conv_m_primary(self, A):
        """A: [E, Kr*c_in] m-primary rows with the radial modulation already applied (it is fused into
        the gather/rotate kernel).  Returns Y [E, extra + Kr*c_out] (m-primary)."""
        groups, weights = self.groups_and_weights()
        return ops.so2_conv(A, self.fc_m0.bias, groups, weights)

    # -- reference-shaped entry point -------------------------------------------------------
    def forward(self, x, x_edge):
        """x: SO3_Embedding with `.embedding` [E, Kr, c_in] in l-primary reduced order (as produced by
        `_rotate`); returns an l-primary SO3_Embedding (and the extra m=0 columns if configured)."""
        lay = self.layout()
        tabs = lay.dev(x.embedding.device)
        E = x.embedding.shape[0]
        A = x.embedding.index_select(1, tabs["to_m"])                    # m-primary
        rad = self.radial_weights(x_edge)
        if rad is not None:
            slot = tabs["rad_slot"].long()
            A = A * rad.view(E, lay.nslot, self.sphere_channels).index_select(1, slot)
        Y = self.conv_m_primary(A.reshape(E, -1))
        extra = self.extra_m0_output_channels or 0
        out = Y[:, extra:].reshape(E, lay.Kr, self.m_output_channels)
        to_l = torch.empty_like(tabs["to_m"])
        to_l[tabs["to_m"]] = torch.arange(lay.Kr, device=to_l.device)
        res = SO3_Embedding(0, x.lmax_list.copy(), self.m_output_channels, device=x.device, dtype=x.dtype)
        res.set_embedding(out.index_select(1, to_l))
        res.set_lmax_mmax(self.lmax_list.copy(), self.mmax_list.copy())
        if self.extra_m0_output_channels is not None:
            return res, Y[:, :extra]
        return res


class SO2_Linear(nn.Module):
    """Never instantiated by any reference model; kept as an import target."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("SO2_Linear is dead code in the reference")
