"""Mathematical invariants that stand in for the golden vectors the reference does not ship
(SURVEY App. B.4): they pin the oracle where third-party arithmetic (e3nn, PyG) is un-vendored."""
import math

import pytest
import torch

from oracle import eqv2_oracle as O


def _frames(n, seed):
    gen = torch.Generator().manual_seed(seed)
    vec = torch.randn(n, 3, generator=gen)
    return vec, O.edge_rot_mat(vec, torch.rand(n, 3, generator=gen) - 0.5)


def test_frame_maps_edge_direction_to_y():
    vec, R = _frames(50, 0)
    y = torch.einsum("eij,ej->ei", R, vec / vec.norm(dim=1, keepdim=True))
    assert torch.allclose(y, torch.tensor([0.0, 1.0, 0.0]).expand_as(y), atol=1e-5)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(50, 3, 3), atol=1e-5)


@pytest.mark.parametrize("lmax", [2, 4, 6])
def test_wigner_orthogonal_and_l1_is_rotation(lmax):
    _, R = _frames(30, 1)
    W = O.rotation_to_wigner(R, lmax)
    K = (lmax + 1) ** 2
    assert torch.allclose(W @ W.transpose(1, 2), torch.eye(K).expand(30, K, K), atol=2e-5)
    assert torch.allclose(W[:, 0, 0], torch.ones(30), atol=1e-6)
    # polar axis y, m = (-1, 0, +1) <-> (x, y, z): the l = 1 block is the frame itself (SURVEY App. B.4 i)
    assert torch.allclose(W[:, 1:4, 1:4], R, atol=1e-5)


@pytest.mark.parametrize("lmax", [2, 4, 6])
def test_wigner_rotates_edge_harmonics_onto_m0(lmax):
    from oracle import sh_basis
    vec, R = _frames(20, 2)
    W = O.rotation_to_wigner(R, lmax)
    Y = torch.tensor(sh_basis.real_sh_xyz(lmax, (vec / vec.norm(dim=1, keepdim=True)).double().numpy()), dtype=torch.float32)
    Yr = torch.einsum("eij,ej->ei", W, Y)
    m0 = torch.tensor([l * l + l for l in range(lmax + 1)])
    mask = torch.ones((lmax + 1) ** 2, dtype=torch.bool)
    mask[m0] = False
    assert float(Yr[:, mask].abs().max()) < 2e-5


@pytest.mark.parametrize("lmax", [2, 4, 6])
def test_from_grid_inverts_to_grid(lmax):
    tg, fg = O.s2_grid_mats(lmax, lmax, 18)
    K = (lmax + 1) ** 2
    ident = torch.einsum("bai,baj->ij", fg.double(), tg.double())
    assert torch.allclose(ident, torch.eye(K, dtype=torch.float64), atol=1e-5)


def test_m_primary_roundtrip_and_softmax_rows():
    lay = O.Layout(6, 2)
    x = torch.arange(lay.Kr)
    assert torch.equal(x[lay.to_m][lay.to_l], x)
    gen = torch.Generator().manual_seed(3)
    logits = torch.randn(40, 4, generator=gen)
    idx = torch.randint(0, 7, (40,), generator=gen)
    a = O.segment_softmax(logits, idx, 7)
    sums = torch.zeros(7, 4).index_add_(0, idx, a)
    present = torch.bincount(idx, minlength=7) > 0
    assert torch.allclose(sums[present], torch.ones_like(sums[present]), atol=1e-6)


def test_model_energy_is_rotation_invariant():
    """Base family only (GATA models are not invariant)."""
    from conftest import golden
    fx = golden("oc20_small_rms_norm_sh.pt")
    h = fx["hyper"]
    hp = O.Hyper(lmax=h["lmax"], mmax=h["mmax"], C=h["C"], H=h["H"], heads=h["heads"], alpha_ch=h["alpha_ch"],
                 value_ch=h["value_ch"], ffn_hidden=h["ffn_hidden"], edge_ch=h["edge_ch"], num_layers=h["num_layers"],
                 norm_type=h["norm_type"], cutoff=h["cutoff"])
    inp = fx["inputs"]
    a = 0.7
    Q = torch.tensor([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1.0]])
    Q = Q @ torch.tensor([[1, 0, 0], [0, math.cos(0.3), -math.sin(0.3)], [0, math.sin(0.3), math.cos(0.3)]])
    outs = []
    for rot in (torch.eye(3), Q):
        vec = fx["edge_vec"] @ rot.T
        gen = torch.Generator().manual_seed(0)
        e, f = O.oc20_forward(fx["params"], hp, inp["atomic_numbers"], inp["batch"], len(inp["natoms"]), fx["edge_index"],
                              fx["edge_distance"], vec, torch.rand(vec.shape, generator=gen) - 0.5)
        outs.append((e, f))
    # the S2 activation is only approximately SO(2)-equivariant: 1e-3 level
    assert float((outs[0][0] - outs[1][0]).abs().max() / outs[0][0].abs().max()) < 5e-3
    assert float((outs[0][1] @ Q.T - outs[1][1]).abs().max() / outs[0][1].abs().max()) < 2e-2
