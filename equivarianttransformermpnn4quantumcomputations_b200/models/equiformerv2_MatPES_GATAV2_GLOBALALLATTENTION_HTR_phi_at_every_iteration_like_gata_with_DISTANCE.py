"""EquiformerV2_MatPES, GATAV2-phi + global all-to-all node attention (reference
models/equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE.py; BASELINE
config 5): the phi-every-layer model with `global_attn = GlobalNodeAttentionHTR_with_ROPE(sphere_channels, lmax,
num_heads, dropout=alpha_drop)` (:231-237) applied to the normed embedding before the energy head (:406-407).
Note the reference instantiates the ROPE class although the file name says `with_DISTANCE`."""
from ..NewFunctions.GATA_and_all2all.activation import GlobalNodeAttentionHTR_with_ROPE, set_structure_sizes
from .equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata import EquiformerV2_MatPES as _Phi
from .equiformerv2_MatPESv2 import init_edge_rot_mat  # noqa: F401


class EquiformerV2_MatPES(_Phi):
    def __init__(self, *args, use_global_attn=True, global_attn_rope=True, **kwargs):
        super().__init__(*args, **kwargs)
        self.global_attn = (GlobalNodeAttentionHTR_with_ROPE(sphere_channels=self.sphere_channels,
                                                             lmax=max(self.lmax_list), num_heads=self.num_heads,
                                                             dropout=self.alpha_drop)
                            if use_global_attn else None)

    def prepare(self, data):
        """Neighbour list + the structure sizes on the host (one read-back here instead of one inside the all-to-all
        attention): with both in `data`, the rest of the step has static shapes (graphs.GraphedTrainStep; the sizes
        are part of the graph signature)."""
        out = super().prepare(data)
        out["natoms_host"] = tuple(int(n) for n in data["natoms"].tolist())
        return out

    def forward(self, data):
        set_structure_sizes(data.get("natoms_host"))
        try:
            return super().forward(data)
        finally:
            set_structure_sizes(None)
