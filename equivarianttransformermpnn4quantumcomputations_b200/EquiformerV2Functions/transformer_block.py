"""SO2EquivariantGraphAttention, FeedForwardNetwork, TransBlockV2 with the reference's constructor
signatures, forward signatures and parameter keys (reference transformer_block.py:39-336, :339-453,
:456-634).

The graph-attention forward is a chain of eight kernel launches instead of ~60 eager ops:

  radial MLP (grouped GEMM + fused LN/SiLU)                      radial_function.py:29
  gather x[src]|x[dst] + Wigner rotate + m-primary + radial mod  transformer_block.py:250-275, so2_ops.py:142-175
  SO(2) conv 1 (one grouped GEMM, all m)                          so2_ops.py:150-185
  separable S2 activation (grid kept in registers)                transformer_block.py:289-294
  attention logits + dst-segment softmax                          transformer_block.py:311-315
  SO(2) conv 2                                                    transformer_block.py:305
  alpha-weight + inverse rotate + deterministic dst-sorted reduce transformer_block.py:321-331
  SO3_LinearV2 projection                                         transformer_block.py:334

Edge tensors stay in m-primary order between kernels, so the reference's dense `_m_primary` /
`_l_primary` permutation einsums disappear.
"""
import copy
import math

import torch
import torch.nn as nn

from .. import ops
from .activation import (GateActivation, S2Activation, ScaledSiLU, ScaledSmoothLeakyReLU, ScaledSwiGLU,
                         SeparableS2Activation, SmoothLeakyReLU, SwiGLU)
from .drop import EquivariantDropoutArraySphericalHarmonics, GraphDropPath
from .layer_norm import (EquivariantLayerNormArray, EquivariantLayerNormArraySphericalHarmonics,
                         EquivariantRMSNormArraySphericalHarmonics, get_normalization_layer)
from .radial_function import RadialFunction
from .so2_ops import SO2_Convolution, SO2_Linear
from .so3 import SO3_Embedding, SO3_Linear, SO3_LinearV2


def _single_resolution(lmax_list, mmax_list, who):
    if len(lmax_list) != 1:
        raise NotImplementedError(f"{who}: a single (lmax, mmax) resolution is supported "
                                  "(every reference config uses one)")
    return lmax_list[0], mmax_list[0]


def edge_scalar_features(module, atomic_numbers, edge_distance, edge_index):
    """x_edge = [rbf | source_embedding(Z_src) | target_embedding(Z_dst)]
    (transformer_block.py:241-248, input_block.py:93-100)."""
    if not module.use_atom_edge_embedding:
        return edge_distance
    src_w, dst_w = module.source_embedding.weight, module.target_embedding.weight
    if src_w.shape[0] != dst_w.shape[0] or module.source_embedding.padding_idx is not None:
        return torch.cat((edge_distance, module.source_embedding(atomic_numbers[edge_index[0]]),
                          module.target_embedding(atomic_numbers[edge_index[1]])), dim=1)
    # same lookups through the row-gather kernel: its backward is a deterministic segmented column sum over a CSR of
    # the element types that is built once per graph (F.embedding's backward sorts the indices on every call)
    plan = ops.edge_plan(edge_index, atomic_numbers.shape[0])
    zs, csr_s, zd, csr_d = plan.element_types(atomic_numbers, src_w.shape[0])
    src = ops.rbf_source_of(edge_distance)
    if src is not None:
        # first-order step and the rbf tensor came from this package's GaussianSmearing: hand the radial MLP a DESCRIPTION
        # of x_edge -- its first layer is evaluated from the raw distances (banded) and two table lookups (ops.rbf_linear)
        return ops.FusedEdgeFeatures(src, src_w, dst_w, zs, csr_s, zd, csr_d)
    return torch.cat((edge_distance, ops.embed_rows(src_w, zs, csr_s), ops.embed_rows(dst_w, zd, csr_d)), dim=1)


class SO2EquivariantGraphAttention(nn.Module):
    def __init__(self, sphere_channels, hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                 output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid, max_num_elements,
                 edge_channels_list, use_atom_edge_embedding=True, use_m_share_rad=False, activation="scaled_silu",
                 use_s2_act_attn=False, use_attn_renorm=True, use_gate_act=False, use_sep_s2_act=True,
                 alpha_drop=0.0):
        super().__init__()
        self.sphere_channels = sphere_channels
        self.hidden_channels = hidden_channels
        self.num_heads = num_heads
        self.attn_alpha_channels = attn_alpha_channels
        self.attn_value_channels = attn_value_channels
        self.output_channels = output_channels
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list
        self.num_resolutions = len(lmax_list)
        _single_resolution(lmax_list, mmax_list, "SO2EquivariantGraphAttention")
        self.SO3_rotation = SO3_rotation
        self.mappingReduced = mappingReduced
        self.SO3_grid = SO3_grid
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
            self.source_embedding, self.target_embedding = None, None
        self.use_s2_act_attn = use_s2_act_attn
        self.use_attn_renorm = use_attn_renorm
        self.use_gate_act = use_gate_act
        self.use_sep_s2_act = use_sep_s2_act
        assert not self.use_s2_act_attn          # reference transformer_block.py:136
        if use_gate_act or not use_sep_s2_act or use_m_share_rad:
            raise NotImplementedError("SO2EquivariantGraphAttention: only the separable-S2 attention path "
                                      "(use_gate_act=False, use_sep_s2_act=True, use_m_share_rad=False; the "
                                      "setting of every reference config) has kernels")
        extra = num_heads * attn_alpha_channels + hidden_channels
        self.so2_conv_1 = SO2_Convolution(2 * sphere_channels, hidden_channels, lmax_list, mmax_list, mappingReduced,
                                          internal_weights=False, edge_channels_list=self.edge_channels_list,
                                          extra_m0_output_channels=extra)
        self.alpha_norm = nn.LayerNorm(attn_alpha_channels) if use_attn_renorm else nn.Identity()
        self.alpha_act = SmoothLeakyReLU()
        self.alpha_dot = nn.Parameter(torch.randn(num_heads, attn_alpha_channels))
        bound = 1.0 / math.sqrt(attn_alpha_channels)
        nn.init.uniform_(self.alpha_dot, -bound, bound)
        self.alpha_dropout = nn.Dropout(alpha_drop) if alpha_drop != 0.0 else None
        self.s2_act = SeparableS2Activation(lmax=max(lmax_list), mmax=max(mmax_list))
        self.so2_conv_2 = SO2_Convolution(hidden_channels, num_heads * attn_value_channels, lmax_list, mmax_list,
                                          mappingReduced, internal_weights=True, edge_channels_list=None,
                                          extra_m0_output_channels=None)
        self.proj = SO3_LinearV2(num_heads * attn_value_channels, output_channels, lmax=lmax_list[0])

    def forward(self, x, atomic_numbers, edge_distance, edge_index):
        lmax, mmax = self.lmax_list[0], self.mmax_list[0]
        lay = ops.CoeffLayout.get(lmax, mmax)
        emb = x.embedding
        plan = ops.edge_plan(edge_index, emb.shape[0])
        wig = self.SO3_rotation[0].wigner_packed
        if wig is None or wig.shape[0] != plan.E:
            raise RuntimeError("SO3_Rotation.set_wigner must be called with this graph's edge frames first")

        x_edge = edge_scalar_features(self, atomic_numbers, edge_distance, edge_index)
        second_order = torch.is_grad_enabled() and edge_distance.requires_grad
        Y = None
        rad_func = self.so2_conv_1.rad_func
        if not second_order and rad_func is not None:
            h, last = rad_func.hidden_and_last(x_edge)                        # radial MLP up to its output layer
            if torch.is_tensor(h) and ops.fused_planes_available(emb, h, last.weight, lay.Kr * 2 * self.sphere_channels):
                # f16 engine, step differentiated once: the radial output layer, the gather / rotate / modulation and the
                # convolution run as one Function -- the rotated rows and the gradient of the radial weights exist only
                # as GEMM operand planes (ops.GatherRotateConvFn)
                groups, weights = self.so2_conv_1.groups_and_weights()
                Y = ops.gather_rotate_conv(emb, h, last.weight, last.bias, self.so2_conv_1.fc_m0.bias, plan, wig, lmax,
                                           mmax, groups, weights)
            else:
                rad = ops.linear(h, last.weight, last.bias) if torch.is_tensor(h) else rad_func(x_edge)
        else:
            rad = self.so2_conv_1.radial_weights(x_edge)                      # [E, n_rad]
        if Y is None:
            A = ops.gather_rotate(emb, rad, plan, wig, lmax, mmax)     # [E, Kr*2C]  m-primary
            Y = self.so2_conv_1.conv_m_primary(A)                             # [E, h*a + H + Kr*H]
        mats = self.SO3_grid[lmax][mmax].kernel_mats("m")
        ln_w = self.alpha_norm.weight if self.use_attn_renorm else None
        ln_b = self.alpha_norm.bias if self.use_attn_renorm else None
        if second_order:
            # positions are being differentiated (forces by autograd, train_MatPES_GATAWandB.py:72-77): use the
            # operators whose backward passes are themselves differentiable
            ha = self.num_heads * self.attn_alpha_channels
            extra = ha + self.hidden_channels
            alpha = ops.attn_alpha(Y[:, :ha], ln_w, ln_b, self.alpha_dot, plan, self.num_heads,
                                          self.attn_alpha_channels)
            Zm = ops.s2_act(Y[:, extra:].reshape(plan.E, lay.Kr, self.hidden_channels), Y[:, ha:extra], mats)
            Zm = Zm.reshape(plan.E, lay.Kr * self.hidden_channels)
        else:
            # f16 engine: the activation writes the second convolution's A operand as planes (no fp32 Z, no split pass)
            z_planes = ops.s2_planes_available(Y, mats, self.hidden_channels, lay.Kr * self.num_heads * self.attn_value_channels,
                                               self.num_heads * self.attn_value_channels, self.num_heads)
            Zm, alpha = ops.edge_act_alpha(Y, ln_w, ln_b, self.alpha_dot, plan, mats, self.num_heads,
                                                 self.attn_alpha_channels, self.hidden_channels, z_planes=z_planes)
        alpha_bound = 1.0
        if self.alpha_dropout is not None:
            if self.training:
                # The reference draws this mask on the OUTPUT LAYOUT OF ITS einsum('bik, ik -> bi') (transformer_block.py:
                # 314-318): [E, heads] with strides (1, E), i.e. head-major memory -- and torch's dropout assigns random
                # numbers in memory order.  Same layout here => same mask for the same seed (tests/test_drop_semantics.py).
                a = alpha.t().contiguous().t().reshape(alpha.shape[0], 1, self.num_heads, 1)
                alpha = self.alpha_dropout(a).reshape(alpha.shape[0], self.num_heads)
                alpha_bound = 1.0 / (1.0 - self.alpha_dropout.p)              # nn.Dropout rescales the kept weights
            else:
                alpha = self.alpha_dropout(alpha)
        if not second_order and ops.gemm_mode() in ("f16x3", "f16"):
            groups, weights = self.so2_conv_2.groups_and_weights()
            out = ops.conv_rotinv_reduce(Zm, alpha, self.so2_conv_2.fc_m0.bias, plan, wig, lmax, mmax, self.num_heads,
                                         alpha_bound, groups, weights)
        else:
            V = self.so2_conv_2.conv_m_primary(Zm)                            # [E, Kr*h*v]
            out = ops.rotinv_reduce(V, alpha, plan, wig, lmax, mmax, lay.Kr, self.num_heads, 1.0)
        msg = SO3_Embedding(0, x.lmax_list.copy(), self.num_heads * self.attn_value_channels,
                            device=x.device, dtype=x.dtype)
        msg.set_embedding(out)
        msg.set_lmax_mmax(self.lmax_list.copy(), self.lmax_list.copy())
        return self.proj(msg)


class FeedForwardNetwork(nn.Module):
    def __init__(self, sphere_channels, hidden_channels, output_channels, lmax_list, mmax_list, SO3_grid,
                 activation="scaled_silu", use_gate_act=False, use_grid_mlp=False, use_sep_s2_act=True):
        super().__init__()
        self.sphere_channels = sphere_channels
        self.hidden_channels = hidden_channels
        self.output_channels = output_channels
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list
        self.num_resolutions = len(lmax_list)
        self.sphere_channels_all = self.num_resolutions * sphere_channels
        self.SO3_grid = SO3_grid
        self.use_gate_act = use_gate_act
        self.use_grid_mlp = use_grid_mlp
        self.use_sep_s2_act = use_sep_s2_act
        self.max_lmax = max(lmax_list)
        if use_gate_act or use_grid_mlp or not use_sep_s2_act:
            raise NotImplementedError("FeedForwardNetwork: only the separable-S2 branch (use_gate_act=False, "
                                      "use_grid_mlp=False, use_sep_s2_act=True; every reference config) has kernels")
        self.so3_linear_1 = SO3_LinearV2(self.sphere_channels_all, hidden_channels, lmax=self.max_lmax)
        self.gating_linear = nn.Linear(self.sphere_channels_all, hidden_channels)
        self.s2_act = SeparableS2Activation(self.max_lmax, self.max_lmax)
        self.so3_linear_2 = SO3_LinearV2(hidden_channels, output_channels, lmax=self.max_lmax)

    def forward(self, input_embedding):
        emb = input_embedding.embedding
        gate = ops.linear(emb[:, 0, :], self.gating_linear.weight, self.gating_linear.bias)
        h = self.so3_linear_1(input_embedding)
        h.embedding = self.s2_act(gate, h.embedding, self.SO3_grid)
        return self.so3_linear_2(h)


class TransBlockV2(nn.Module):
    def __init__(self, sphere_channels, attn_hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                 ffn_hidden_channels, output_channels, lmax_list, mmax_list, SO3_rotation, mappingReduced, SO3_grid,
                 max_num_elements, edge_channels_list, use_atom_edge_embedding=True, use_m_share_rad=False,
                 attn_activation="silu", use_s2_act_attn=False, use_attn_renorm=True, ffn_activation="silu",
                 use_gate_act=False, use_grid_mlp=False, use_sep_s2_act=True, norm_type="rms_norm_sh",
                 alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0):
        super().__init__()
        max_lmax = max(lmax_list)
        self.norm_1 = get_normalization_layer(norm_type, lmax=max_lmax, num_channels=sphere_channels)
        self.ga = SO2EquivariantGraphAttention(
            sphere_channels=sphere_channels, hidden_channels=attn_hidden_channels, num_heads=num_heads,
            attn_alpha_channels=attn_alpha_channels, attn_value_channels=attn_value_channels,
            output_channels=sphere_channels, lmax_list=lmax_list, mmax_list=mmax_list, SO3_rotation=SO3_rotation,
            mappingReduced=mappingReduced, SO3_grid=SO3_grid, max_num_elements=max_num_elements,
            edge_channels_list=edge_channels_list, use_atom_edge_embedding=use_atom_edge_embedding,
            use_m_share_rad=use_m_share_rad, activation=attn_activation, use_s2_act_attn=use_s2_act_attn,
            use_attn_renorm=use_attn_renorm, use_gate_act=use_gate_act, use_sep_s2_act=use_sep_s2_act,
            alpha_drop=alpha_drop)
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

    def forward(self, x, atomic_numbers, edge_distance, edge_index, batch):
        out = x
        res = out.embedding
        out.embedding = self.norm_1(out.embedding)
        out = self.ga(out, atomic_numbers, edge_distance, edge_index)
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
        return out
