"""EquiformerV2_QM9 (reference models/equiformerv2_qm9.py:81-788): per-molecule property heads, no PBC.
Same constructor arguments, forward(data) contract, `generate_graph` return tuple and state_dict keys.
The non-periodic radius graph (equiformerv2_qm9.py:423-525: strict 0 < d < r_c, per-destination
nearest `max_neighbors`) is built on the GPU by `ops.radius_graph`."""
import torch
import torch.nn as nn

from .. import ops
from ..EquiformerV2Functions.drop import set_num_graphs
from ..EquiformerV2Functions.edge_rot_mat import init_edge_rot_mat
from ..EquiformerV2Functions.input_block import EdgeDegreeEmbedding
from ..EquiformerV2Functions.layer_norm import get_normalization_layer
from ..EquiformerV2Functions.radial_function import RadialFunction
from ..EquiformerV2Functions.so3 import CoefficientMappingModule, SO3_Embedding, SO3_Rotation
from ..EquiformerV2Functions.transformer_block import FeedForwardNetwork, TransBlockV2
from .common import GaussianSmearing, build_so3_grid, init_linear, segment_sum

_AVG_NUM_NODES_QM9 = 18.0
_AVG_DEGREE_QM9 = 6.0


class EquiformerV2_QM9(nn.Module):
    def __init__(self, num_targets=12, use_pbc=False, regress_forces=False, otf_graph=True, max_neighbors=50,
                 max_radius=5.0, max_num_elements=10, num_layers=8, sphere_channels=128, attn_hidden_channels=128,
                 num_heads=8, attn_alpha_channels=32, attn_value_channels=16, ffn_hidden_channels=512,
                 norm_type="rms_norm_sh", lmax_list=[4], mmax_list=[2], grid_resolution=None,
                 num_sphere_samples=128, edge_channels=128, use_atom_edge_embedding=True,
                 share_atom_edge_embedding=False, use_m_share_rad=False, distance_function="gaussian",
                 num_distance_basis=512, attn_activation="scaled_silu", use_s2_act_attn=False,
                 use_attn_renorm=True, ffn_activation="scaled_silu", use_gate_act=False, use_grid_mlp=False,
                 use_sep_s2_act=True, alpha_drop=0.1, drop_path_rate=0.05, proj_drop=0.0, weight_init="normal"):
        super().__init__()
        assert weight_init in ["normal", "uniform"]
        assert distance_function in ["gaussian"]
        for k, v in list(locals().items()):
            if k not in ("self", "__class__"):
                setattr(self, k, v)
        self.cutoff = max_radius
        self.device = "cpu"
        self.grad_forces = False
        self.num_resolutions = len(lmax_list)
        self.sphere_channels_all = self.num_resolutions * sphere_channels
        if share_atom_edge_embedding:
            assert use_atom_edge_embedding
            self.block_use_atom_edge_embedding = False
        else:
            self.block_use_atom_edge_embedding = use_atom_edge_embedding

        self.sphere_embedding = nn.Embedding(max_num_elements, self.sphere_channels_all)
        # 600 Gaussians of width 2 regardless of num_distance_basis (equiformerv2_qm9.py:261-266)
        self.distance_expansion = GaussianSmearing(0.0, self.cutoff, 600, 2.0)
        self.edge_channels_list = [int(self.distance_expansion.num_output)] + [edge_channels] * 2
        if share_atom_edge_embedding and use_atom_edge_embedding:
            self.source_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            self.target_embedding = nn.Embedding(max_num_elements, self.edge_channels_list[-1])
            self.edge_channels_list[0] += 2 * self.edge_channels_list[-1]
        else:
            self.source_embedding = self.target_embedding = None

        self.SO3_rotation = nn.ModuleList([SO3_Rotation(l) for l in lmax_list])
        self.mappingReduced = CoefficientMappingModule(lmax_list, mmax_list)
        self.SO3_grid = build_so3_grid(lmax_list, grid_resolution)
        self.edge_degree_embedding = EdgeDegreeEmbedding(
            sphere_channels, lmax_list, mmax_list, self.SO3_rotation, self.mappingReduced, max_num_elements,
            self.edge_channels_list, self.block_use_atom_edge_embedding, rescale_factor=_AVG_DEGREE_QM9)
        self.blocks = nn.ModuleList([
            TransBlockV2(sphere_channels, attn_hidden_channels, num_heads, attn_alpha_channels, attn_value_channels,
                         ffn_hidden_channels, sphere_channels, lmax_list, mmax_list, self.SO3_rotation,
                         self.mappingReduced, self.SO3_grid, max_num_elements, self.edge_channels_list,
                         self.block_use_atom_edge_embedding, use_m_share_rad, attn_activation, use_s2_act_attn,
                         use_attn_renorm, ffn_activation, use_gate_act, use_grid_mlp, use_sep_s2_act, norm_type,
                         alpha_drop, drop_path_rate, proj_drop)
            for _ in range(num_layers)])
        self.norm = get_normalization_layer(norm_type, lmax=max(lmax_list), num_channels=sphere_channels)
        self.output_blocks = nn.ModuleList([
            FeedForwardNetwork(sphere_channels, ffn_hidden_channels, 1, lmax_list, mmax_list, self.SO3_grid,
                               ffn_activation, use_gate_act, use_grid_mlp, use_sep_s2_act)
            for _ in range(num_targets)])
        # reference: normal init leaves the default nn.Linear init in 'uniform' mode (equiformerv2_qm9.py:712-721)
        self.apply(lambda m: init_linear(m, "normal" if weight_init == "normal" else "keep"))
        self.apply(self._uniform_init_rad_func_linear_weights)

    def _uniform_init_rad_func_linear_weights(self, m):
        if isinstance(m, RadialFunction):
            m.apply(lambda mm: init_linear(mm, "uniform") if isinstance(mm, nn.Linear) else None)

    @property
    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def generate_graph(self, data):
        """-> (edge_index, edge_distance, edge_distance_vec, cell_offsets=None, None, neighbors)."""
        if "edge_index" in data:
            ei, d, v = data["edge_index"], data["edge_distance"], data["edge_distance_vec"]
        else:
            ei, d, v = ops.radius_graph(data["pos"], data["natoms"], data["batch"], self.max_radius,
                                        self.max_neighbors)
        neighbors = torch.bincount(ei[1], minlength=data["pos"].shape[0])
        return ei, d, v, None, None, neighbors

    def _init_edge_rot_mat(self, data, edge_index, edge_distance_vec):
        return init_edge_rot_mat(edge_distance_vec)

    def prepare(self, data):
        """Data-dependent, host-synchronising head of a forward pass (graphs.GraphedTrainStep): radius graph (one
        read-back of the edge count) and edge frames; everything after it has shapes fixed by (atoms, edges)."""
        if "edge_index" in data:
            ei, d, v = data["edge_index"], data["edge_distance"], data["edge_distance_vec"]
        else:
            ei, d, v = ops.radius_graph(data["pos"], data["natoms"], data["batch"], self.max_radius, self.max_neighbors)
        return {"edge_index": ei, "edge_distance": d, "edge_distance_vec": v,
                "edge_frames": self._init_edge_rot_mat(data, ei, v)}

    def forward(self, data):
        set_num_graphs(len(data["natoms"]))          # GraphDropPath: no batch.max() read-back (drop.py)
        try:
            return self._forward(data)
        finally:
            set_num_graphs(None)

    def _forward(self, data):
        self.batch_size = len(data["natoms"])
        self.dtype, self.device = data["pos"].dtype, data["pos"].device
        atomic_numbers = data["atomic_numbers"].long()
        num_atoms = atomic_numbers.shape[0]
        if "edge_frames" in data:
            edge_index, edge_distance, edge_vec, frames = (data["edge_index"], data["edge_distance"],
                                                           data["edge_distance_vec"], data["edge_frames"])
        else:
            edge_index, edge_distance, edge_vec, _, _, _ = self.generate_graph(data)
            frames = self._init_edge_rot_mat(data, edge_index, edge_vec)
        for rot in self.SO3_rotation:
            rot.set_wigner(frames)

        x = SO3_Embedding(num_atoms, self.lmax_list, self.sphere_channels, self.device, self.dtype)
        x.embedding[:, 0, :] = self.sphere_embedding(atomic_numbers)
        rbf = self.distance_expansion(edge_distance)
        if self.share_atom_edge_embedding and self.use_atom_edge_embedding:
            rbf = torch.cat((rbf, self.source_embedding(atomic_numbers[edge_index[0]]),
                             self.target_embedding(atomic_numbers[edge_index[1]])), dim=1)
        x.embedding = x.embedding + self.edge_degree_embedding(atomic_numbers, rbf, edge_index).embedding
        for block in self.blocks:
            x = block(x, atomic_numbers, rbf, edge_index, batch=data["batch"])
        x.embedding = self.norm(x.embedding)

        preds = [segment_sum(head(x).embedding[:, 0, 0], data["batch"], self.batch_size)
                 for head in self.output_blocks]
        return torch.stack(preds, dim=1)

    @torch.jit.ignore
    def no_weight_decay(self):
        from ..EquiformerV2Functions.so3 import SO3_LinearV2
        names = {n for n, _ in self.named_parameters()}
        out = []
        for mod_name, mod in self.named_modules():
            if isinstance(mod, (nn.Linear, SO3_LinearV2, nn.LayerNorm)):
                for p_name, _ in mod.named_parameters():
                    if isinstance(mod, (nn.Linear, SO3_LinearV2)) and "weight" in p_name:
                        continue
                    full = mod_name + "." + p_name
                    assert full in names
                    out.append(full)
        return set(out)
