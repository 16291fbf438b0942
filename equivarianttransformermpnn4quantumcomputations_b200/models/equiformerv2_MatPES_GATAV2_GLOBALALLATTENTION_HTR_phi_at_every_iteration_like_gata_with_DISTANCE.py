"""EquiformerV2_MatPES, GATAV2-phi + global all-to-all node attention (reference
models/equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE.py; BASELINE
config 5): the phi-every-layer model with `global_attn = GlobalNodeAttentionHTR_with_ROPE(sphere_channels, lmax,
num_heads, dropout=alpha_drop)` (:231-237) applied to the normed embedding before the energy head (:406-407).
Note the reference instantiates the ROPE class although the file name says `with_DISTANCE`."""
from ..NewFunctions.GATA_and_all2all.activation import GlobalNodeAttentionHTR_with_ROPE
from .equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata import EquiformerV2_MatPES as _Phi
from .equiformerv2_MatPESv2 import init_edge_rot_mat  # noqa: F401


class EquiformerV2_MatPES(_Phi):
    def __init__(self, *args, use_global_attn=True, global_attn_rope=True, **kwargs):
        super().__init__(*args, **kwargs)
        self.global_attn = (GlobalNodeAttentionHTR_with_ROPE(sphere_channels=self.sphere_channels,
                                                             lmax=max(self.lmax_list), num_heads=self.num_heads,
                                                             dropout=self.alpha_drop)
                            if use_global_attn else None)
