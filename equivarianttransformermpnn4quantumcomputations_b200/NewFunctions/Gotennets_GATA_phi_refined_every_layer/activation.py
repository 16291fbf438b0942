"""Same classes as `Gotennet_morethaninspired.activation`; GATAValueActivation requires `num_rbf` here
(reference Gotennets_GATA_phi_refined_every_layer/activation.py:296-352)."""
from ..Gotennet_morethaninspired.activation import *  # noqa: F401,F403
from ..Gotennet_morethaninspired.activation import HTR, GATAValueActivation as _Base


class GATAValueActivation(_Base):
    def __init__(self, sphere_channels, hidden_channels, edge_channels, lmax, mmax, num_rbf):
        super().__init__(sphere_channels, hidden_channels, edge_channels, lmax, mmax, num_rbf=num_rbf)

    def forward(self, attn_output, t_ij, h_j, X_j, rl_ij, phi_r):
        return super().forward(attn_output, t_ij, h_j, X_j, rl_ij, phi_r=phi_r)
