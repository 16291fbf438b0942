"""EquiformerV2_MatPES with GATA + HTR (reference models/equiformerv2_MatPES_GATAV2.py:62-479; BASELINE config 4):
edge-scalar stream t_ij (initialised from (h_i + h_j) * W_erp(rbf), refined by HTR in every block), edge spherical
harmonics rl_ij in the original frame (detached), GATA value activation inside the attention block.
Energy-only forward; forces by autograd in the training loop (see equiformerv2_MatPESv2.py in this package).
Same constructor arguments, forward(data) dict contract and state_dict keys as the reference."""
import torch
import torch.nn as nn

from .. import ops
from ..EquiformerV2Functions.input_block import EdgeDegreeEmbedding
from ..EquiformerV2Functions.layer_norm import get_normalization_layer
from ..EquiformerV2Functions.radial_function import RadialFunction
from ..EquiformerV2Functions.so3 import CoefficientMappingModule, SO3_Embedding, SO3_Rotation
from ..NewFunctions.Gotennet_morethaninspired.transformer_block import FeedForwardNetwork, TransBlockV2
from .common import GaussianSmearing, build_so3_grid, init_linear, segment_sum
from .equiformerv2_MatPESv2 import init_edge_rot_mat

_AVG_DEGREE_MATPES = 12.0


class EquiformerV2_MatPES(nn.Module):
    # hooks for the phi-at-every-iteration twin (equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata.py)
    _avg_degree = _AVG_DEGREE_MATPES
    _phi_every_layer = False
    _block_cls = TransBlockV2

    def __init__(self, use_pbc=True, regress_forces=True, regress_stress=False, otf_graph=True, max_neighbors=20,
                 max_radius=6.0, max_num_elements=100, num_layers=6, sphere_channels=128, attn_hidden_channels=128,
                 num_heads=8, attn_alpha_channels=32, attn_value_channels=16, ffn_hidden_channels=512,
                 norm_type="rms_norm_sh", lmax_list=None, mmax_list=None, grid_resolution=18, num_sphere_samples=128,
                 edge_channels=128, use_atom_edge_embedding=True, share_atom_edge_embedding=False,
                 use_m_share_rad=False, distance_function="gaussian", num_distance_basis=512,
                 attn_activation="scaled_silu", use_s2_act_attn=False, use_attn_renorm=True,
                 ffn_activation="scaled_silu", use_gate_act=False, use_grid_mlp=False, use_sep_s2_act=True,
                 alpha_drop=0.05, drop_path_rate=0.05, proj_drop=0.0, weight_init="normal"):
        super().__init__()
        lmax_list = [4] if lmax_list is None else lmax_list
        mmax_list = [2] if mmax_list is None else mmax_list
        for k, v in list(locals().items()):
            if k not in ("self", "__class__"):
                setattr(self, k, v)
        self.cutoff = max_radius
        self.device = "cpu"
        self.block_use_atom_edge_embedding = False if share_atom_edge_embedding else use_atom_edge_embedding
        self.num_resolutions = len(lmax_list)
        self.sphere_channels_all = self.num_resolutions * sphere_channels

        self.sphere_embedding = nn.Embedding(max_num_elements, self.sphere_channels_all)
        self.distance_expansion = GaussianSmearing(0.0, self.cutoff, 600, 2.0)
        self.edge_channels_list = [int(self.distance_expansion.num_output)] + [edge_channels] * 2
        if share_atom_edge_embedding and use_atom_edge_embedding:
            self.source_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            self.target_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            self.edge_channels_list[0] += 2 * self.edge_channels_list[-1]
        else:
            self.source_embedding = self.target_embedding = None
        num_rbf = int(self.distance_expansion.num_output)
        self.W_erp = nn.Linear(num_rbf, edge_channels)
        self.h_proj = nn.Linear(sphere_channels, edge_channels)
        self.SO3_rotation = nn.ModuleList([SO3_Rotation(l) for l in lmax_list])
        self.mappingReduced = CoefficientMappingModule(lmax_list, mmax_list)
        self.SO3_grid = build_so3_grid(lmax_list, grid_resolution)
        self.edge_degree_embedding = EdgeDegreeEmbedding(
            sphere_channels, lmax_list, mmax_list, self.SO3_rotation, self.mappingReduced, max_num_elements,
            self.edge_channels_list, self.block_use_atom_edge_embedding, rescale_factor=self._avg_degree)
        extra_kw = dict(num_rbf=num_rbf) if self._phi_every_layer else {}
        self.blocks = nn.ModuleList([
            self._block_cls(sphere_channels=sphere_channels, attn_hidden_channels=attn_hidden_channels, num_heads=num_heads,
                         attn_alpha_channels=attn_alpha_channels, attn_value_channels=attn_value_channels,
                         ffn_hidden_channels=ffn_hidden_channels, output_channels=sphere_channels, lmax_list=lmax_list,
                         mmax_list=mmax_list, SO3_rotation=self.SO3_rotation, mappingReduced=self.mappingReduced,
                         SO3_grid=self.SO3_grid, max_num_elements=max_num_elements,
                         edge_channels_list=self.edge_channels_list, edge_channels=edge_channels,
                         use_atom_edge_embedding=self.block_use_atom_edge_embedding, use_m_share_rad=use_m_share_rad,
                         attn_activation=attn_activation, use_s2_act_attn=use_s2_act_attn,
                         use_attn_renorm=use_attn_renorm, ffn_activation=ffn_activation, use_gate_act=use_gate_act,
                         use_grid_mlp=use_grid_mlp, use_sep_s2_act=use_sep_s2_act, norm_type=norm_type,
                         alpha_drop=alpha_drop, drop_path_rate=drop_path_rate, proj_drop=proj_drop, **extra_kw)
            for _ in range(num_layers)])
        self.norm = get_normalization_layer(norm_type, lmax=max(lmax_list), num_channels=sphere_channels)
        self.energy_block = FeedForwardNetwork(sphere_channels, ffn_hidden_channels, 1, lmax_list, mmax_list,
                                               self.SO3_grid, ffn_activation, use_gate_act, use_grid_mlp, use_sep_s2_act)
        self.apply(lambda m: init_linear(m, "normal" if weight_init == "normal" else "keep"))
        self.apply(self._uniform_init_rad_func_linear_weights)

    def _uniform_init_rad_func_linear_weights(self, m):
        if isinstance(m, RadialFunction):
            m.apply(lambda mm: init_linear(mm, "uniform") if isinstance(mm, nn.Linear) else None)

    @property
    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def generate_graph(self, pos, batch, cell, natoms=None):
        if natoms is None:
            natoms = torch.bincount(batch, minlength=cell.shape[0])
        edge_index, _, _, _ = ops.radius_graph_matpes(pos, cell, natoms, batch, self.max_radius, self.max_neighbors, 2)
        dvec = pos[edge_index[1]] - pos[edge_index[0]]
        return edge_index, torch.norm(dvec, dim=1), dvec

    def prepare(self, data):
        """Data-dependent head of a forward pass (graphs.GraphedTrainStep): the periodic neighbour list."""
        with torch.no_grad():
            return {"edge_index": self.generate_graph(data["pos"].detach(), data["batch"], data["cell"],
                                                      data["natoms"])[0]}

    def _compute_rl_ij(self, edge_distance_vec):
        return ops.edge_sh(edge_distance_vec, max(self.lmax_list))

    def _init_t_ij(self, x_embedding, edge_dist_feat, edge_index):
        h_all = x_embedding[:, 0, :]
        h_sum = ops.linear(h_all[edge_index[0]] + h_all[edge_index[1]], self.h_proj.weight, self.h_proj.bias)
        return h_sum * ops.linear(edge_dist_feat, self.W_erp.weight, self.W_erp.bias)

    def forward(self, data):
        self.batch_size = len(data["natoms"])
        self.dtype, self.device = data["pos"].dtype, data["pos"].device
        atomic_numbers = data["atomic_numbers"].long()
        num_atoms = atomic_numbers.shape[0]
        pos = data["pos"]
        if "edge_index" in data:
            edge_index = data["edge_index"]
            edge_vec = pos[edge_index[1]] - pos[edge_index[0]]
            edge_distance = torch.norm(edge_vec, dim=1)
        else:
            edge_index, edge_distance, edge_vec = self.generate_graph(pos, data["batch"], data["cell"], data["natoms"])
        frames = init_edge_rot_mat(edge_vec)
        for rot in self.SO3_rotation:
            rot.set_wigner(frames)
        rl_ij = self._compute_rl_ij(edge_vec)

        x = SO3_Embedding(num_atoms, self.lmax_list, self.sphere_channels, self.device, self.dtype)
        x.embedding[:, 0, :] = self.sphere_embedding(atomic_numbers)
        rbf = self.distance_expansion(edge_distance)
        edge_feat = rbf
        if self.share_atom_edge_embedding and self.use_atom_edge_embedding:
            edge_feat = torch.cat((rbf, self.source_embedding(atomic_numbers[edge_index[0]]),
                                   self.target_embedding(atomic_numbers[edge_index[1]])), dim=1)
        x.embedding = x.embedding + self.edge_degree_embedding(atomic_numbers, edge_feat, edge_index).embedding
        t_ij = self._init_t_ij(x.embedding, rbf, edge_index)
        phi_kw = dict(phi_r=edge_feat) if self._phi_every_layer else {}
        for block in self.blocks:
            x, t_ij = block(x, atomic_numbers, edge_feat, edge_index, batch=data["batch"], t_ij=t_ij, rl_ij=rl_ij,
                            **phi_kw)
        x.embedding = self.norm(x.embedding)
        if getattr(self, "global_attn", None) is not None:          # config-5 twin: all-to-all attention after the norm
            x.embedding = self.global_attn(x.embedding, data["batch"], pos)
        node_energy = self.energy_block(x).embedding[:, 0, 0]
        energy_total = segment_sum(node_energy, data["batch"], self.batch_size)
        energy_out = (energy_total / data["natoms"].to(node_energy.dtype)).unsqueeze(1)
        return {"energy": energy_out, "energy_total": energy_total, "pos": pos}
