"""Equivariant normalisation layers with the reference's class names and parameter / buffer keys
(reference layer_norm.py).  The three variants the models use run ONE kernel pair
(`eqv2_equiv_norm_fwd/bwd`) parameterised by degree groups (see csrc/norm.cu); like the reference
(`@autocast(enabled=False)`) they always compute in fp32."""
import torch
import torch.nn as nn

from .. import ops


def get_normalization_layer(norm_type, lmax, num_channels, eps=1e-5, affine=True, normalization="component"):
    assert norm_type in ["layer_norm", "layer_norm_sh", "rms_norm_sh"]
    cls = {"layer_norm": EquivariantLayerNormArray,
           "layer_norm_sh": EquivariantLayerNormArraySphericalHarmonics,
           "rms_norm_sh": EquivariantRMSNormArraySphericalHarmonicsV2}[norm_type]
    return cls(lmax, num_channels, eps, affine, normalization)


def get_l_to_all_m_expand_index(lmax):
    return torch.tensor([l for l in range(lmax + 1) for _ in range(2 * l + 1)], dtype=torch.long)


def _check(affine, normalization):
    assert normalization in ["norm", "component"]
    if not affine or normalization != "component":
        raise NotImplementedError("equivariant norms: only affine=True, normalization='component' "
                                  "(what every reference model uses) has a kernel")


class EquivariantLayerNormArray(nn.Module):
    """'layer_norm' (layer_norm.py:38-108): one RMS per degree; l = 0 centred; bias on l = 0."""

    def __init__(self, lmax, num_channels, eps=1e-5, affine=True, normalization="component"):
        super().__init__()
        _check(affine, normalization)
        self.lmax, self.num_channels, self.eps = lmax, num_channels, eps
        self.affine, self.normalization = affine, normalization
        self.affine_weight = nn.Parameter(torch.ones(lmax + 1, num_channels))
        self.affine_bias = nn.Parameter(torch.zeros(num_channels))

    def __repr__(self):
        return f"{self.__class__.__name__}(lmax={self.lmax}, num_channels={self.num_channels}, eps={self.eps})"

    def forward(self, node_input):
        return ops.equiv_norm(node_input.float(), self.affine_weight, self.affine_bias,
                                     "layer_norm", self.lmax, self.eps)


class EquivariantLayerNormArraySphericalHarmonics(nn.Module):
    """'layer_norm_sh' (layer_norm.py:112-201): nn.LayerNorm on l = 0, one shared degree-balanced RMS
    for all l > 0."""

    def __init__(self, lmax, num_channels, eps=1e-5, affine=True, normalization="component",
                 std_balance_degrees=True):
        super().__init__()
        _check(affine, normalization)
        if not std_balance_degrees:
            raise NotImplementedError("layer_norm_sh: std_balance_degrees=False is never used")
        self.lmax, self.num_channels, self.eps = lmax, num_channels, eps
        self.affine, self.normalization, self.std_balance_degrees = affine, normalization, True
        self.norm_l0 = nn.LayerNorm(num_channels, eps=eps, elementwise_affine=True)
        self.affine_weight = nn.Parameter(torch.ones(lmax, num_channels))
        bw = torch.zeros((lmax + 1) ** 2 - 1, 1)
        for l in range(1, lmax + 1):
            bw[l * l - 1:(l + 1) ** 2 - 1] = 1.0 / (2 * l + 1)
        self.register_buffer("balance_degree_weight", bw / max(lmax, 1))

    def __repr__(self):
        return (f"{self.__class__.__name__}(lmax={self.lmax}, num_channels={self.num_channels}, "
                f"eps={self.eps}, std_balance_degrees={self.std_balance_degrees})")

    def forward(self, node_input):
        w = torch.cat([self.norm_l0.weight.view(1, -1), self.affine_weight], dim=0)
        return ops.equiv_norm(node_input.float(), w, self.norm_l0.bias, "layer_norm_sh", self.lmax, self.eps)


class EquivariantRMSNormArraySphericalHarmonics(nn.Module):
    """Un-centred RMS norm (layer_norm.py:205-262).  Importable for API parity; no model selects it
    (`get_normalization_layer('rms_norm_sh')` returns the V2 class below)."""

    def __init__(self, lmax, num_channels, eps=1e-5, affine=True, normalization="component"):
        super().__init__()
        raise NotImplementedError("EquivariantRMSNormArraySphericalHarmonics (V1) is not reachable from "
                                  "get_normalization_layer; use 'rms_norm_sh' (V2)")


class EquivariantRMSNormArraySphericalHarmonicsV2(nn.Module):
    """'rms_norm_sh' (layer_norm.py:265-351): l = 0 centred, one degree-balanced RMS over all K rows."""

    def __init__(self, lmax, num_channels, eps=1e-5, affine=True, normalization="component", centering=True,
                 std_balance_degrees=True):
        super().__init__()
        _check(affine, normalization)
        if not (centering and std_balance_degrees):
            raise NotImplementedError("rms_norm_sh: only centering=True, std_balance_degrees=True is used")
        self.lmax, self.num_channels, self.eps = lmax, num_channels, eps
        self.affine, self.normalization = affine, normalization
        self.centering, self.std_balance_degrees = True, True
        self.affine_weight = nn.Parameter(torch.ones(lmax + 1, num_channels))
        self.affine_bias = nn.Parameter(torch.zeros(num_channels))
        self.register_buffer("expand_index", get_l_to_all_m_expand_index(lmax))
        bw = torch.zeros((lmax + 1) ** 2, 1)
        for l in range(lmax + 1):
            bw[l * l:(l + 1) ** 2] = 1.0 / (2 * l + 1)
        self.register_buffer("balance_degree_weight", bw / (lmax + 1))

    def __repr__(self):
        return (f"{self.__class__.__name__}(lmax={self.lmax}, num_channels={self.num_channels}, eps={self.eps}, "
                f"centering={self.centering}, std_balance_degrees={self.std_balance_degrees})")

    def forward(self, node_input):
        return ops.equiv_norm(node_input.float(), self.affine_weight, self.affine_bias,
                                     "rms_norm_sh", self.lmax, self.eps)
