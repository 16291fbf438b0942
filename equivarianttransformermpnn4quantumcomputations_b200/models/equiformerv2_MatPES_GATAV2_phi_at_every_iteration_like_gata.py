"""EquiformerV2_MatPES, GATAV2 with the phi factor refined in every block (reference
models/equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata.py; BASELINE config 4): the GATAV2 model with blocks
from `NewFunctions/Gotennets_GATA_phi_refined_every_layer` (`num_rbf`, `phi_r=edge_dist_feat`, :207,:428) and
`_AVG_DEGREE_MATPES = 50.51` (:60)."""
from ..NewFunctions.Gotennets_GATA_phi_refined_every_layer.transformer_block import TransBlockV2
from .equiformerv2_MatPES_GATAV2 import EquiformerV2_MatPES as _GATAV2
from .equiformerv2_MatPESv2 import init_edge_rot_mat  # noqa: F401

_AVG_DEGREE_MATPES = 50.51


class EquiformerV2_MatPES(_GATAV2):
    _avg_degree = _AVG_DEGREE_MATPES
    _phi_every_layer = True
    _block_cls = TransBlockV2
