"""GATA variants of SO2EquivariantGraphAttention / TransBlockV2 (reference
NewFunctions/Gotennet_morethaninspired/transformer_block.py:50-372, :480-662); FeedForwardNetwork is the base one.

What the reference computes and then discards is not computed here: in this family the SO(2)
output of `so2_conv_1` is overwritten by the GATA value activation, so only the `extra_m0` columns of `fc_m0`
are evaluated; the `so2_conv_1.so2_m_conv.*` parameters exist (state_dict parity) and end a step with
`grad is None`, exactly as in the reference."""
import copy
import math

import torch
import torch.nn as nn

from ... import ops
from ...EquiformerV2Functions.drop import EquivariantDropoutArraySphericalHarmonics, GraphDropPath
from ...EquiformerV2Functions.layer_norm import get_normalization_layer
from ...EquiformerV2Functions.so2_ops import SO2_Convolution
from ...EquiformerV2Functions.so3 import SO3_Embedding, SO3_LinearV2
from ...EquiformerV2Functions.transformer_block import FeedForwardNetwork, edge_scalar_features  # noqa: F401
from .activation import GATAValueActivation, HTR, SmoothLeakyReLU


class SO2EquivariantGraphAttention(nn.Module):
    def __init__(self, sphere_channels, hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                 output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid, max_num_elements,
                 edge_channels_list, edge_channels, use_atom_edge_embedding=True, use_m_share_rad=False,
                 activation="scaled_silu", use_s2_act_attn=False, use_attn_renorm=True, use_gate_act=False,
                 use_sep_s2_act=True, alpha_drop=0.0, num_rbf=None):
        super().__init__()
        self.num_rbf = num_rbf
        self.sphere_channels = sphere_channels
        self.hidden_channels = hidden_channels
        self.num_heads = num_heads
        self.attn_alpha_channels = attn_alpha_channels
        self.attn_value_channels = attn_value_channels
        self.output_channels = output_channels
        self.lmax_list, self.mmax_list = lmax_list, mmax_list
        self.num_resolutions = len(lmax_list)
        self.lmax = max(lmax_list)
        self.edge_channels = edge_channels
        self.SO3_rotation, self.mappingReduced, self.SO3_grid = SO3_rotation, mappingReduced, SO3_grid
        self.max_num_elements = max_num_elements
        self.edge_channels_list = copy.deepcopy(edge_channels_list)
        self.use_atom_edge_embedding = use_atom_edge_embedding
        self.use_m_share_rad = use_m_share_rad
        if use_atom_edge_embedding:
            self.source_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            self.target_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            nn.init.uniform_(self.source_embedding.weight.data, -0.001, 0.001)
            nn.init.uniform_(self.target_embedding.weight.data, -0.001, 0.001)
            self.edge_channels_list[0] = self.edge_channels_list[0] + 2 * self.edge_channels_list[-1]
        else:
            self.source_embedding = self.target_embedding = None
        self.use_s2_act_attn, self.use_attn_renorm = use_s2_act_attn, use_attn_renorm
        self.use_gate_act, self.use_sep_s2_act = use_gate_act, use_sep_s2_act
        assert not self.use_s2_act_attn
        if len(lmax_list) != 1 or use_gate_act or not use_sep_s2_act or use_m_share_rad:
            raise NotImplementedError("GATA attention: single resolution, use_sep_s2_act=True, no gate activation / "
                                      "shared radial weights (the setting of configs 4-5) has kernels")
        S = 1 + 2 * self.lmax
        extra = num_heads * attn_alpha_channels + S * hidden_channels
        self.so2_conv_1 = SO2_Convolution(2 * sphere_channels, hidden_channels, lmax_list, mmax_list, mappingReduced,
                                          internal_weights=False, edge_channels_list=self.edge_channels_list,
                                          extra_m0_output_channels=extra)
        self.alpha_norm = nn.LayerNorm(attn_alpha_channels) if use_attn_renorm else nn.Identity()
        self.alpha_act = SmoothLeakyReLU()
        self.alpha_dot = nn.Parameter(torch.randn(num_heads, attn_alpha_channels))
        bound = 1.0 / math.sqrt(attn_alpha_channels)
        nn.init.uniform_(self.alpha_dot, -bound, bound)
        self.alpha_dropout = nn.Dropout(alpha_drop) if alpha_drop != 0.0 else None
        self.value_act = GATAValueActivation(sphere_channels=sphere_channels, hidden_channels=hidden_channels,
                                             edge_channels=edge_channels, lmax=self.lmax, mmax=max(mmax_list),
                                             num_rbf=num_rbf)
        self.so2_conv_2 = SO2_Convolution(hidden_channels, num_heads * attn_value_channels, lmax_list, mmax_list,
                                          mappingReduced, internal_weights=True, edge_channels_list=None,
                                          extra_m0_output_channels=None)
        self.proj = SO3_LinearV2(num_heads * attn_value_channels, output_channels, lmax=lmax_list[0])

    def forward(self, x, atomic_numbers, edge_distance, edge_index, t_ij, rl_ij, phi_r=None):
        lmax, mmax = self.lmax_list[0], self.mmax_list[0]
        lay = ops.CoeffLayout.get(lmax, mmax)
        emb = x.embedding
        plan = ops.edge_plan(edge_index, emb.shape[0])
        wig = self.SO3_rotation[0].wigner_packed
        if wig is None or wig.shape[0] != plan.E:
            raise RuntimeError("SO3_Rotation.set_wigner must be called with this graph's edge frames first")
        x_edge = edge_scalar_features(self, atomic_numbers, edge_distance, edge_index)
        rad = self.so2_conv_1.radial_weights(x_edge)
        A = ops.gather_rotate(emb, rad, plan, wig, lmax, mmax)               # [E, Kr*2C] m-primary
        # only the extra columns of fc_m0 are live in this family (reference :292-326)
        conv1 = self.so2_conv_1
        extra = conv1.extra_m0_output_channels
        n_m0 = (lmax + 1) * 2 * self.sphere_channels
        Y0 = ops.SliceMm.apply(A, conv1.fc_m0.bias[:extra].contiguous(), ((0, n_m0),), ((0, extra),), extra, True,
                               conv1.fc_m0.weight[:extra].contiguous())
        ha = self.num_heads * self.attn_alpha_channels
        ln_w = self.alpha_norm.weight if self.use_attn_renorm else None
        ln_b = self.alpha_norm.bias if self.use_attn_renorm else None
        alpha = ops.attn_alpha(Y0[:, :ha], ln_w, ln_b, self.alpha_dot, plan, self.num_heads, self.attn_alpha_channels)
        attn_output = alpha.mean(dim=1, keepdim=True) * Y0[:, ha:]
        x_dst = emb[edge_index[1]]                                           # un-rotated neighbour features
        msg = self.value_act(attn_output=attn_output, t_ij=t_ij, h_j=x_dst[:, 0, :], X_j=x_dst[:, 1:, :], rl_ij=rl_ij,
                             phi_r=phi_r if self.num_rbf is not None else None)
        tabs = lay.dev(emb.device)
        Zm = msg.index_select(1, tabs["to_m"]).reshape(plan.E, lay.Kr * self.hidden_channels)   # l- -> m-primary
        V = self.so2_conv_2.conv_m_primary(Zm)
        if self.alpha_dropout is not None:
            alpha = self.alpha_dropout(alpha)
        out = ops.rotinv_reduce(V, alpha, plan, wig, lmax, mmax, lay.Kr, self.num_heads, 1.0)
        res = SO3_Embedding(0, x.lmax_list.copy(), self.num_heads * self.attn_value_channels, device=x.device, dtype=x.dtype)
        res.set_embedding(out)
        res.set_lmax_mmax(self.lmax_list.copy(), self.lmax_list.copy())
        return self.proj(res)


class TransBlockV2(nn.Module):
    def __init__(self, sphere_channels, attn_hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                 ffn_hidden_channels, output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid,
                 max_num_elements, edge_channels_list, edge_channels, use_atom_edge_embedding=True,
                 use_m_share_rad=False, attn_activation="silu", use_s2_act_attn=False, use_attn_renorm=True,
                 ffn_activation="silu", use_gate_act=False, use_grid_mlp=False, use_sep_s2_act=True,
                 norm_type="rms_norm_sh", alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0, num_rbf=None):
        super().__init__()
        max_lmax = max(lmax_list)
        self.norm_1 = get_normalization_layer(norm_type, lmax=max_lmax, num_channels=sphere_channels)
        self.htr = HTR(sphere_channels=sphere_channels, edge_channels=edge_channels, lmax=max_lmax)
        self.ga = SO2EquivariantGraphAttention(
            sphere_channels=sphere_channels, hidden_channels=attn_hidden_channels, num_heads=num_heads,
            attn_alpha_channels=attn_alpha_channels, attn_value_channels=attn_value_channels,
            output_channels=sphere_channels, lmax_list=lmax_list, mmax_list=mmax_list, SO3_rotation=SO3_rotation,
            mappingReduced=mappingReduced, SO3_grid=SO3_grid, max_num_elements=max_num_elements,
            edge_channels_list=edge_channels_list, edge_channels=edge_channels,
            use_atom_edge_embedding=use_atom_edge_embedding, use_m_share_rad=use_m_share_rad,
            activation=attn_activation, use_s2_act_attn=use_s2_act_attn, use_attn_renorm=use_attn_renorm,
            use_gate_act=use_gate_act, use_sep_s2_act=use_sep_s2_act, alpha_drop=alpha_drop, num_rbf=num_rbf)
        self.drop_path = GraphDropPath(drop_path_rate) if drop_path_rate > 0.0 else None
        self.proj_drop = EquivariantDropoutArraySphericalHarmonics(proj_drop, drop_graph=False) if proj_drop > 0.0 else None
        self.norm_2 = get_normalization_layer(norm_type, lmax=max_lmax, num_channels=sphere_channels)
        self.ffn = FeedForwardNetwork(
            sphere_channels=sphere_channels, hidden_channels=ffn_hidden_channels, output_channels=output_channels,
            lmax_list=lmax_list, mmax_list=mmax_list, SO3_grid=SO3_grid, activation=ffn_activation,
            use_gate_act=use_gate_act, use_grid_mlp=use_grid_mlp, use_sep_s2_act=use_sep_s2_act)
        self.ffn_shortcut = (SO3_LinearV2(sphere_channels, output_channels, lmax=max_lmax)
                             if sphere_channels != output_channels else None)

    def _drop(self, t, batch):
        if self.drop_path is not None:
            t = self.drop_path(t, batch)
        if self.proj_drop is not None:
            t = self.proj_drop(t, batch)
        return t

    def forward(self, x, atomic_numbers, edge_distance, edge_index, batch, t_ij, rl_ij, phi_r=None):
        out = x
        X_all = x.embedding[:, 1:, :]
        t_ij = self.htr(t_ij, X_all[edge_index[0]], X_all[edge_index[1]], rl_ij)    # edge stream update (un-normed x)
        res = out.embedding
        out.embedding = self.norm_1(out.embedding)
        out = self.ga(out, atomic_numbers, edge_distance, edge_index, t_ij=t_ij, rl_ij=rl_ij, phi_r=phi_r)
        out.embedding = self._drop(out.embedding, batch) + res
        res = out.embedding
        out.embedding = self.norm_2(out.embedding)
        out = self.ffn(out)
        out.embedding = self._drop(out.embedding, batch)
        if self.ffn_shortcut is not None:
            sc = SO3_Embedding(0, out.lmax_list.copy(), self.ffn_shortcut.in_features, device=out.device, dtype=out.dtype)
            sc.set_embedding(res)
            sc.set_lmax_mmax(out.lmax_list.copy(), out.lmax_list.copy())
            res = self.ffn_shortcut(sc).embedding
        out.embedding = out.embedding + res
        return out, t_ij
