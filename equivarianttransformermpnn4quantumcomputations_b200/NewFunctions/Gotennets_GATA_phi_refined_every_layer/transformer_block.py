"""GATA block with the per-layer phi factor (reference Gotennets_GATA_phi_refined_every_layer/transformer_block.py):
identical to the `Gotennet_morethaninspired` fork plus the `num_rbf` constructor argument and the `phi_r` forward
argument."""
from ..Gotennet_morethaninspired import transformer_block as _b
from ..Gotennet_morethaninspired.transformer_block import FeedForwardNetwork  # noqa: F401


class SO2EquivariantGraphAttention(_b.SO2EquivariantGraphAttention):
    def __init__(self, sphere_channels, hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                 output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid, max_num_elements,
                 edge_channels_list, edge_channels, num_rbf, use_atom_edge_embedding=True, use_m_share_rad=False,
                 activation="scaled_silu", use_s2_act_attn=False, use_attn_renorm=True, use_gate_act=False,
                 use_sep_s2_act=True, alpha_drop=0.0):
        super().__init__(sphere_channels, hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                         output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid, max_num_elements,
                         edge_channels_list, edge_channels, use_atom_edge_embedding, use_m_share_rad, activation,
                         use_s2_act_attn, use_attn_renorm, use_gate_act, use_sep_s2_act, alpha_drop, num_rbf=num_rbf)

    def forward(self, x, atomic_numbers, edge_distance, edge_index, t_ij, rl_ij, phi_r):
        return super().forward(x, atomic_numbers, edge_distance, edge_index, t_ij, rl_ij, phi_r=phi_r)


class TransBlockV2(_b.TransBlockV2):
    def __init__(self, sphere_channels, attn_hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                 ffn_hidden_channels, output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid,
                 max_num_elements, edge_channels_list, edge_channels, num_rbf, use_atom_edge_embedding=True,
                 use_m_share_rad=False, attn_activation="silu", use_s2_act_attn=False, use_attn_renorm=True,
                 ffn_activation="silu", use_gate_act=False, use_grid_mlp=False, use_sep_s2_act=True,
                 norm_type="rms_norm_sh", alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0):
        super().__init__(sphere_channels, attn_hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                         ffn_hidden_channels, output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced,
                         SO3_grid, max_num_elements, edge_channels_list, edge_channels, use_atom_edge_embedding,
                         use_m_share_rad, attn_activation, use_s2_act_attn, use_attn_renorm, ffn_activation,
                         use_gate_act, use_grid_mlp, use_sep_s2_act, norm_type, alpha_drop, drop_path_rate, proj_drop,
                         num_rbf=num_rbf)

    def forward(self, x, atomic_numbers, edge_distance, edge_index, batch, t_ij, rl_ij, phi_r):
        return super().forward(x, atomic_numbers, edge_distance, edge_index, batch, t_ij, rl_ij, phi_r=phi_r)
