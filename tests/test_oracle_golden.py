"""Pins oracle/eqv2_oracle.py (the CPU restatement) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py): outputs and parameter gradients of the OC20- and QM9-shaped models,
Wigner-D matrices, S2 grid matrices, edge frames and graphs."""
import pytest
import torch

from conftest import golden
from oracle import eqv2_oracle as O


def _hyper(h):
    return O.Hyper(lmax=h["lmax"], mmax=h["mmax"], C=h["C"], H=h["H"], heads=h["heads"], alpha_ch=h["alpha_ch"],
                   value_ch=h["value_ch"], ffn_hidden=h["ffn_hidden"], edge_ch=h["edge_ch"], num_layers=h["num_layers"],
                   norm_type=h["norm_type"], grid_res=h["grid_res"], num_rbf=h["num_rbf"], cutoff=h["cutoff"],
                   max_elements=h["max_elements"], max_neighbors=h["max_neighbors"],
                   num_targets=h.get("num_targets", 1))


def _rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / (b.detach().double().abs().max() + 1e-30))


@pytest.mark.parametrize("norm_type", ["rms_norm_sh", "layer_norm_sh", "layer_norm"])
def test_oracle_oc20_matches_reference(norm_type):
    fx = golden(f"oc20_small_{norm_type}.pt")
    hp = _hyper(fx["hyper"])
    P = {k: v.clone().requires_grad_(True) for k, v in fx["params"].items()}
    inp = fx["inputs"]
    energy, forces = O.oc20_forward(P, hp, inp["atomic_numbers"], inp["batch"], len(inp["natoms"]), fx["edge_index"],
                                    fx["edge_distance"], fx["edge_vec"], fx["rand_vec"])
    assert _rel(energy, fx["energy"]) < 2e-6 and _rel(forces, fx["forces"]) < 2e-6
    (energy.sum() + (forces * torch.linspace(-1, 1, forces.numel()).view_as(forces)).sum()).backward()
    for k, g in fx["grads"].items():
        assert _rel(P[k].grad, g) < 5e-5, k


def test_oracle_qm9_matches_reference():
    fx = golden("qm9_small.pt")
    hp = _hyper(fx["hyper"])
    hp.avg_degree = 6.0                       # _AVG_DEGREE_QM9 (equiformerv2_qm9.py:82)
    P = {k: v.clone().requires_grad_(True) for k, v in fx["params"].items()}
    inp = fx["inputs"]
    pred = O.qm9_forward(P, hp, inp["atomic_numbers"], inp["batch"], len(inp["natoms"]), fx["edge_index"],
                         fx["edge_distance"], fx["edge_vec"], fx["rand_vec"])
    assert _rel(pred, fx["pred"]) < 2e-6


def test_oracle_graph_builders_match_reference_graphs():
    fx = golden("qm9_small.pt")
    inp = fx["inputs"]
    ei, d, v = O.radius_graph_qm9(inp["pos"], inp["batch"], fx["hyper"]["cutoff"], fx["hyper"]["max_neighbors"])
    assert torch.equal(ei, fx["edge_index"]) and torch.equal(d, fx["edge_distance"])
    fx = golden("oc20_small_rms_norm_sh.pt")
    inp = fx["inputs"]
    ei, d, v = O.radius_graph_pbc_fairchem(inp["pos"], inp["cell"], inp["batch"], inp["natoms"], fx["hyper"]["cutoff"],
                                           fx["hyper"]["max_neighbors"])
    assert torch.equal(ei, fx["edge_index"]) and torch.allclose(v, fx["edge_vec"])


def test_oracle_components_match_reference():
    c = golden("components.pt")
    R = O.edge_rot_mat(c["edge_vec"], c["rand_vec"])
    assert torch.allclose(R, c["rot"], atol=1e-6)
    for lmax in (2, 4, 6):
        W = O.rotation_to_wigner(c["rot"], lmax)
        assert torch.allclose(W, c[f"wigner_l{lmax}"], atol=2e-5), lmax
    for (l, m) in ((4, 2), (4, 4), (6, 2), (6, 6), (2, 2), (3, 2), (3, 3)):
        tg, fg = O.s2_grid_mats(l, m, 18)
        assert torch.allclose(tg, c[f"to_grid_{l}_{m}"], atol=1e-6)
        assert torch.allclose(fg, c[f"from_grid_{l}_{m}"], atol=1e-6)


def test_oracle_matpes_builders_match_reference_graphs():
    fx = golden("matpes_v2_small.pt")
    inp, hp = fx["inputs"], fx["hyper"]
    for version, pre in ((1, "v1_"), (2, "")):
        ei, d, v, _ = O.radius_graph_matpes(inp["pos"], inp["cell"], inp["batch"], hp["cutoff"], hp["max_neighbors"], version)
        n = inp["pos"].shape[0]
        oa = O.canonical_edge_order(ei, n, tiebreak=d)
        ob = O.canonical_edge_order(fx[pre + "edge_index"], n, tiebreak=fx[pre + "edge_distance"])
        assert torch.equal(ei[:, oa], fx[pre + "edge_index"][:, ob]), version
        assert torch.allclose(d[oa], fx[pre + "edge_distance"][ob], atol=1e-6)


def test_oracle_matpes_v2_train_step_matches_reference():
    """energy, autograd forces (create_graph) and the parameter gradients of a loss on both."""
    fx = golden("matpes_v2_small.pt")
    h = fx["hyper"]
    hp = _hyper(h)
    hp.avg_degree = 12.0                      # _AVG_DEGREE_MATPES (equiformerv2_MatPESv2.py:68)
    P = {k: v.clone().requires_grad_(True) for k, v in fx["params"].items()}
    inp = fx["inputs"]
    pos = inp["pos"].clone().requires_grad_(True)
    e_tot = O.matpes_v2_forward(P, hp, inp["atomic_numbers"], inp["batch"], inp["natoms"], pos, fx["edge_index"])
    assert _rel(e_tot, fx["energy_total"]) < 2e-6
    forces = -torch.autograd.grad(e_tot.sum(), pos, create_graph=True)[0]
    assert _rel(forces, fx["forces"]) < 1e-5
    energy = (e_tot / inp["natoms"].float()).unsqueeze(1)
    wf = torch.linspace(-1, 1, forces.numel()).view_as(forces)
    we = torch.linspace(0.5, 1.5, energy.numel()).view_as(energy)
    ((energy * we).sum() + (forces * wf).sum()).backward()
    for k, g in fx["grads"].items():
        assert _rel(P[k].grad, g) < 1e-4, k
