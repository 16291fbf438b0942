"""`NewFunctions/Gotennets_GATA_phi_refined_every_layer`: the GATA fork whose value activation carries the extra
`phi_proj(phi_r)` factor in every block (reference activation.py:314,352; used by BASELINE configs 4-5 through
equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata.py)."""
