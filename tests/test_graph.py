"""Neighbour-list kernels vs the oracle builders and the graphs the unmodified reference built
(golden fixtures): indices bit-exact after the canonical (dst, src[, distance]) sort; distances and
vectors to fp32 rounding."""
import pytest
import torch

from conftest import golden
from helpers import pkg
from oracle import eqv2_oracle as O


def _canon(ei, d, v, n):
    order = O.canonical_edge_order(ei.cpu(), n, tiebreak=d.detach().cpu())
    return ei.cpu()[:, order], d.cpu()[order], v.cpu()[order]


def test_radius_graph_matches_reference_qm9_graph(backend):
    fx = golden("qm9_small.pt")
    ops = pkg("ops")
    inp = backend.to(fx["inputs"])
    hp = fx["hyper"]
    ei, d, v = ops.radius_graph(inp["pos"], inp["natoms"], inp["batch"], hp["cutoff"], hp["max_neighbors"])
    n = inp["pos"].shape[0]
    a = _canon(ei, d, v, n)
    b = _canon(fx["edge_index"], fx["edge_distance"], fx["edge_vec"], n)
    assert a[0].shape == b[0].shape and torch.equal(a[0], b[0])
    assert torch.allclose(a[1], b[1], rtol=0, atol=1e-6)
    assert torch.allclose(a[2], b[2], rtol=0, atol=1e-6)
    # builder output is dst-sorted (the CSR the attention kernels use)
    assert bool((ei[1][1:] >= ei[1][:-1]).all())


@pytest.mark.parametrize("max_nb", [None, 3, 50])
def test_radius_graph_matches_oracle(backend, max_nb):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(17)
    natoms = torch.tensor([7, 1, 12, 20])
    pos = torch.cat([torch.rand(int(k), 3, generator=gen) * 6.0 for k in natoms])
    batch = torch.repeat_interleave(torch.arange(len(natoms)), natoms)
    ref = O.radius_graph_qm9(pos, batch, 3.5, max_nb)
    ei, d, v = ops.radius_graph(backend.to(pos), backend.to(natoms), backend.to(batch), 3.5, max_nb)
    a, b = _canon(ei, d, v, len(pos)), _canon(*ref, len(pos))
    assert torch.equal(a[0], b[0])
    assert torch.allclose(a[1], b[1], atol=1e-6) and torch.allclose(a[2], b[2], atol=1e-6)


def test_radius_graph_empty_and_single(backend):
    ops = pkg("ops")
    pos = backend.to(torch.tensor([[0.0, 0, 0], [10.0, 0, 0], [0, 10.0, 0]]))
    natoms = backend.to(torch.tensor([1, 2]))
    batch = backend.to(torch.tensor([0, 1, 1]))
    ei, d, v = ops.radius_graph(pos, natoms, batch, 5.0, 10)
    assert ei.shape == (2, 0) and d.shape == (0,) and v.shape == (0, 3)


@pytest.mark.parametrize("norm_type", ["rms_norm_sh"])
def test_radius_graph_pbc_matches_reference_oc20_graph(backend, norm_type):
    fx = golden(f"oc20_small_{norm_type}.pt")
    ops = pkg("ops")
    inp = backend.to(fx["inputs"])
    hp = fx["hyper"]
    ei, d, v = ops.radius_graph_pbc(inp["pos"], inp["cell"], inp["natoms"], inp["batch"], hp["cutoff"],
                                    hp["max_neighbors"])
    n = inp["pos"].shape[0]
    a = _canon(ei, d, v, n)
    b = _canon(fx["edge_index"], fx["edge_distance"], fx["edge_vec"], n)
    assert a[0].shape == b[0].shape and torch.equal(a[0], b[0])
    assert torch.allclose(a[1], b[1], atol=2e-6) and torch.allclose(a[2], b[2], atol=2e-6)


@pytest.mark.parametrize("strict", [False, True])
def test_radius_graph_pbc_matches_oracle_skewed_cells(backend, strict):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(5)
    cells = torch.stack([4.0 * torch.eye(3) + 0.4 * torch.randn(3, 3, generator=gen),
                         torch.diag(torch.tensor([3.0, 5.0, 9.0])) + 0.2 * torch.randn(3, 3, generator=gen)])
    natoms = torch.tensor([5, 3])
    pos = torch.cat([torch.rand(int(k), 3, generator=gen) @ cells[g] for g, k in enumerate(natoms)])
    batch = torch.repeat_interleave(torch.arange(2), natoms)
    ref = O.radius_graph_pbc_fairchem(pos, cells, batch, natoms, 5.0, 6, strict=strict)
    ei, d, v = ops.radius_graph_pbc(backend.to(pos), backend.to(cells), backend.to(natoms), backend.to(batch), 5.0, 6,
                                    strict=strict)
    a, b = _canon(ei, d, v, len(pos)), _canon(*ref, len(pos))
    assert a[0].shape == b[0].shape and torch.equal(a[0], b[0])
    assert torch.allclose(a[1], b[1], atol=2e-6) and torch.allclose(a[2], b[2], atol=2e-6)


def test_radius_graph_pbc_nonstrict_truncation_semantics(backend):
    """Independent numpy pin of the fairchem-core truncation rule the kernel follows: on SQUARED distances,
    a centre with more than max_nb candidates keeps d2 <= d2_sorted[max_nb] + 0.01, i.e. at least max_nb + 1 edges;
    self images need d2 > 1e-4; the cutoff is inclusive."""
    import numpy as np
    ops = pkg("ops")
    rng = np.random.default_rng(3)
    cell = np.diag([4.1, 4.3, 4.7]) + 0.1 * rng.normal(size=(3, 3))
    n, cutoff, max_nb = 7, 6.0, 9
    p = rng.uniform(0, 1, (n, 3)) @ cell
    pos, cells = torch.tensor(p, dtype=torch.float32), torch.tensor(cell, dtype=torch.float32).unsqueeze(0)
    natoms, batch = torch.tensor([n]), torch.zeros(n, dtype=torch.long)
    ei, d, v = ops.radius_graph_pbc(backend.to(pos), backend.to(cells), backend.to(natoms), backend.to(batch), cutoff, max_nb)
    ei, d = ei.cpu().numpy(), d.cpu().numpy()
    p32, c32 = pos.double().numpy(), cells[0].double().numpy()
    R = 4
    offs = np.array([[a, b, c] for a in range(-R, R + 1) for b in range(-R, R + 1) for c in range(-R, R + 1)], float) @ c32
    for i in range(n):
        d2 = (((p32[None, :, None, :] + offs[None, None, :, :]) - p32[i]) ** 2).sum(-1)[0]        # [n, S]
        cand = np.sort(d2[(d2 <= cutoff * cutoff) & (d2 > 1e-4)])
        keep = cand[cand <= cand[max_nb] + 0.01] if len(cand) > max_nb else cand
        mine = np.sort(d[ei[1] == i].astype(np.float64) ** 2)
        assert len(mine) == len(keep) and len(mine) >= min(len(cand), max_nb + 1), (i, len(mine), len(keep))
        assert np.allclose(mine, keep, rtol=1e-5)


def test_segment_sum(backend):
    ops = pkg("ops")
    v = torch.randn(11, requires_grad=True)
    batch = torch.tensor([0, 0, 0, 1, 3, 3, 3, 3, 4, 4, 4])
    vb = backend.to(v.detach()).requires_grad_(True)
    out = ops.segment_sum_nodes(vb, backend.to(batch), 6)
    ref = torch.zeros(6).index_add_(0, batch, v)
    assert torch.allclose(out.cpu(), ref.detach(), atol=1e-6)
    w = torch.arange(6.0)
    (out * backend.to(w)).sum().backward()
    assert torch.allclose(vb.grad.cpu(), w[batch])


def test_edge_plan_csr(backend):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(2)
    ei = torch.randint(0, 9, (2, 40), generator=gen)
    plan = ops.EdgePlan(backend.to(ei), 9)
    for idx, perm, rowptr in ((ei[1], plan.perm_dst, plan.rowptr_dst), (ei[0], plan.perm_src, plan.rowptr_src)):
        order = torch.sort(idx, stable=True)[1]
        assert torch.equal(perm.cpu().long(), order)
        counts = torch.bincount(idx, minlength=9)
        assert torch.equal(rowptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))


@pytest.mark.parametrize("version", [1, 2])
def test_radius_graph_matpes_matches_reference_graph(backend, version):
    fx = golden("matpes_v2_small.pt")
    ops = pkg("ops")
    inp = backend.to(fx["inputs"])
    hp = fx["hyper"]
    pre = "v1_" if version == 1 else ""
    ei, d, v, img = ops.radius_graph_matpes(inp["pos"], inp["cell"], inp["natoms"], inp["batch"], hp["cutoff"],
                                            hp["max_neighbors"], version)
    n = inp["pos"].shape[0]
    a = _canon(ei, d, v, n)
    b = _canon(fx[pre + "edge_index"], fx[pre + "edge_distance"], fx[pre + "edge_vec"], n)
    assert a[0].shape == b[0].shape and torch.equal(a[0], b[0])
    assert torch.allclose(a[1], b[1], atol=2e-6)
    if version == 2:                      # v1 duplicates of one (src, dst) pair differ only by image: compare as sets
        assert torch.allclose(a[2], b[2], atol=2e-6)
    assert int(img.min()) >= 0 and int(img.max()) <= 26


@pytest.mark.parametrize("version", [1, 2])
def test_radius_graph_matpes_matches_oracle_random_cells(backend, version):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(31)
    cells = torch.stack([3.5 * torch.eye(3) + 0.3 * torch.randn(3, 3, generator=gen) for _ in range(3)])
    natoms = torch.tensor([4, 9, 1])
    pos = torch.cat([torch.rand(int(k), 3, generator=gen) @ cells[g] for g, k in enumerate(natoms)])
    batch = torch.repeat_interleave(torch.arange(3), natoms)
    for max_nb in (None, 5, 12):
        ref = O.radius_graph_matpes(pos, cells, batch, 4.0, max_nb, version)
        ei, d, v, img = ops.radius_graph_matpes(backend.to(pos), backend.to(cells), backend.to(natoms), backend.to(batch),
                                                4.0, max_nb, version)
        a, b = _canon(ei, d, v, len(pos)), _canon(ref[0], ref[1], ref[2], len(pos))
        assert a[0].shape == b[0].shape and torch.equal(a[0], b[0]), (version, max_nb)
        assert torch.allclose(a[1], b[1], atol=2e-6)
