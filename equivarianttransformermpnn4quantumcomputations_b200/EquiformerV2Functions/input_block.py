"""EdgeDegreeEmbedding (reference input_block.py:8-131): radial MLP -> m = 0 coefficients ->
inverse Wigner rotation -> sum over incoming edges -> / rescale_factor.

The zero padding of the m != 0 rows, the `_l_primary` permutation, the dense inverse rotation and the
`index_add_` of the reference collapse into one launch of the deterministic `rotinv_reduce` kernel
reading only the first (lmax+1) m-primary rows."""
import copy

import torch
import torch.nn as nn

from .. import ops
from .radial_function import RadialFunction
from .so3 import SO3_Embedding


class EdgeDegreeEmbedding(nn.Module):
    def __init__(self, sphere_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, max_num_elements,
                 edge_channels_list, use_atom_edge_embedding, rescale_factor):
        super().__init__()
        self.sphere_channels = sphere_channels
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list
        self.num_resolutions = len(lmax_list)
        if self.num_resolutions != 1:
            raise NotImplementedError("EdgeDegreeEmbedding: a single (lmax, mmax) resolution is supported")
        self.SO3_rotation = SO3_rotation
        self.mappingReduced = mappingReduced
        self.m_0_num_coefficients = int(mappingReduced.m_size[0])
        self.m_all_num_coefficents = len(mappingReduced.l_harmonic)
        self.max_num_elements = max_num_elements
        self.edge_channels_list = copy.deepcopy(edge_channels_list)
        self.use_atom_edge_embedding = use_atom_edge_embedding
        if use_atom_edge_embedding:
            self.source_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            self.target_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            nn.init.uniform_(self.source_embedding.weight.data, -0.001, 0.001)
            nn.init.uniform_(self.target_embedding.weight.data, -0.001, 0.001)
            self.edge_channels_list[0] = self.edge_channels_list[0] + 2 * self.edge_channels_list[-1]
        else:
            self.source_embedding, self.target_embedding = None, None
        self.edge_channels_list.append(self.m_0_num_coefficients * sphere_channels)
        self.rad_func = RadialFunction(self.edge_channels_list)
        self.rescale_factor = rescale_factor

    def forward(self, atomic_numbers, edge_distance, edge_index):
        from .transformer_block import edge_scalar_features
        lmax, mmax = self.lmax_list[0], self.mmax_list[0]
        plan = ops.edge_plan(edge_index, atomic_numbers.shape[0])
        wig = self.SO3_rotation[0].wigner_packed
        if wig is None or wig.shape[0] != plan.E:
            raise RuntimeError("SO3_Rotation.set_wigner must be called with this graph's edge frames first")
        x_edge = edge_scalar_features(self, atomic_numbers, edge_distance, edge_index)
        m0 = self.rad_func(x_edge)                                            # [E, (lmax+1)*C]
        out = ops.rotinv_reduce(m0, None, plan, wig, lmax, mmax, self.m_0_num_coefficients, 0,
                                       1.0 / float(self.rescale_factor))
        res = SO3_Embedding(0, self.lmax_list.copy(), self.sphere_channels, device=out.device, dtype=out.dtype)
        res.set_embedding(out)
        res.set_lmax_mmax(self.lmax_list.copy(), self.lmax_list.copy())
        return res
