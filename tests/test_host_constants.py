"""Host-side constants of the product (J matrices, S2 grid matrices), computed by recurrences /
quadrature in `_so3_math.py`, against the independent construction in oracle/sh_basis.py and the
matrices the unmodified reference produced (golden components)."""
import numpy as np
import pytest
import torch

from conftest import golden
from helpers import pkg
from oracle import sh_basis


def test_jd_matches_oracle_and_is_involutive():
    sm = pkg("_so3_math")
    ref = sh_basis.make_jd(6)
    for l, J in enumerate(sm.jd_blocks(6)):
        assert np.allclose(J, ref[l], atol=1e-10), l
        assert np.allclose(J @ J, np.eye(2 * l + 1), atol=1e-10)
        assert np.allclose(J, J.T, atol=1e-12)


@pytest.mark.parametrize("lm", [(4, 2), (4, 4), (6, 2), (6, 6), (2, 2), (3, 2), (3, 3)])
def test_grid_matrices_match_reference(lm):
    sm = pkg("_so3_math")
    c = golden("components.pt")
    tg, fg = sm.s2_grid_matrices(lm[0], lm[1], 18, 18)
    assert np.allclose(tg, c[f"to_grid_{lm[0]}_{lm[1]}"].numpy(), atol=2e-6)
    assert np.allclose(fg, c[f"from_grid_{lm[0]}_{lm[1]}"].numpy(), atol=2e-6)


def test_state_dict_keys_match_reference_parameter_names():
    """Strict load_state_dict of reference checkpoints needs identical parameter names."""
    from helpers import build_oc20, build_qm9
    for name, build in (("oc20_small_rms_norm_sh.pt", build_oc20), ("oc20_small_layer_norm_sh.pt", build_oc20),
                        ("qm9_small.pt", build_qm9)):
        fx = golden(name)
        model = build(fx["hyper"], torch.device("cpu"))
        assert set(dict(model.named_parameters())) == set(fx["params"])
        sd = model.state_dict()
        assert "blocks.0.ga.so2_conv_1.fc_m0.weight" in sd and "distance_expansion.offset" in sd
        assert any(k.endswith("to_grid_mat") for k in sd) and any(k.endswith("mapping.to_m") for k in sd)
