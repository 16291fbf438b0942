"""eqv2_edge_frames against the reference formulas (edge_rot_mat.py:13-80 restated in oracle/eqv2_oracle.py::edge_rot_mat,
pinned by tests/golden/components.pt from the unmodified reference; deterministic variant equiformerv2_MatPESv2.py:41-66
restated in oracle.edge_rot_mat_deterministic)."""
import pytest
import torch

from conftest import golden
from helpers import fixed_rand_like, pkg
from oracle import eqv2_oracle as O


def test_random_helper_frames_match_reference_vectors(backend):
    fx = golden("components.pt")
    mod = pkg("EquiformerV2Functions.edge_rot_mat")
    with fixed_rand_like(fx["rand_vec"] + 0.5):
        R = mod.init_edge_rot_mat(backend.to(fx["edge_vec"]))
    assert R.shape == fx["rot"].shape and not R.requires_grad
    assert float((R.cpu() - fx["rot"]).abs().max()) < 2e-6
    # orthonormal, middle row = edge direction
    I = torch.eye(3).expand_as(fx["rot"])
    assert float((R.cpu() @ R.cpu().transpose(1, 2) - I).abs().max()) < 1e-5
    d = fx["edge_vec"] / fx["edge_vec"].norm(dim=1, keepdim=True)
    assert float((R.cpu()[:, 1] - d).abs().max()) < 1e-6


def test_parallel_helper_is_swapped_and_short_edges_are_reported(backend, capsys):
    ops = pkg("ops")
    vec = torch.tensor([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.3, 0.3, 0.3], [0.00001, 0.0, 0.0]])
    draw = torch.tensor([[0.4, 0.0, 0.0], [0.0, -0.3, 0.0], [0.1, 0.1, 0.1], [0.0, 0.2, 0.1]])   # (anti)parallel helpers
    R = ops.edge_frames(backend.to(vec), backend.to(draw)).cpu()
    ref = O.edge_rot_mat(vec, draw)
    assert float((R - ref).abs().max()) < 2e-6
    assert "Error edge_vec_0_distance" in capsys.readouterr().out


def test_deterministic_frames_match_reference_formula(backend):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(4)
    vec = torch.randn(257, 3, generator=gen) * 3
    vec[:3] = torch.tensor([[1.0, 0.0, 0.0], [0.0, -2.0, 0.0], [0.5, 0.5, 0.5]])      # axis-aligned bonds, exact ties
    R = ops.edge_frames(backend.to(vec), None).cpu()
    ref = O.edge_rot_mat_deterministic(vec)
    assert float((R - ref).abs().max()) < 2e-6
