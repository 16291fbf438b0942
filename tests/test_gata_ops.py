"""HTR inner product and GATAValueActivation assembly kernels (csrc/gata.cu) against the reference expressions
(NewFunctions/Gotennet_morethaninspired/activation.py:197-201 `vector_rejection`, :236-250 the degree loop; :370-414 the
output assembly), evaluated in float64 with torch autograd: value, first derivatives, and the derivative OF the first
derivatives (what the double backward of the force loss needs)."""
import pytest
import torch

from helpers import pkg, rel_err


def _htr_ref(q, k, rl, lmax):
    def rej(rep, r):
        ru = r.unsqueeze(-1)
        return rep - (rep * ru).sum(dim=1, keepdim=True) * ru
    w, off = 0.0, 0
    for l in range(1, lmax + 1):
        n = 2 * l + 1
        r = rl[:, off:off + n]
        w = w + (rej(q[:, off:off + n], r) * rej(k[:, off:off + n], -r)).sum(dim=1) / n
        off += n
    return w


def _gata_ref(comb, Xp, rl, lmax, mmax):
    C = Xp.shape[-1]
    chunks = comb.split(C, dim=-1)
    out = [torch.nn.functional.silu(chunks[0]).unsqueeze(1)]
    off = 0
    for l in range(1, lmax + 1):
        n, w = 2 * l + 1, min(2 * l + 1, 2 * mmax + 1)
        out.append(chunks[l].unsqueeze(1) * rl[:, off:off + w].unsqueeze(-1) + chunks[lmax + l].unsqueeze(1) * Xp[:, off:off + w])
        off += n
    return torch.cat(out, dim=1)


def _second_order_check(fn_mine, fn_ref, inputs, dev):
    """value, gradients, and gradients of a random functional of the gradients, mine (fp32, kernels) vs ref (fp64)."""
    gen = torch.Generator().manual_seed(7)
    mine = [t.clone().to(dev).requires_grad_(True) for t in inputs]
    ref = [t.double().requires_grad_(True) for t in inputs]
    ym, yr = fn_mine(*mine), fn_ref(*ref)
    assert rel_err(ym, yr) < 2e-6
    go = torch.randn(yr.shape, generator=gen)
    gm = torch.autograd.grad(ym, mine, go.to(dev), create_graph=True)
    gr = torch.autograd.grad(yr, ref, go.double(), create_graph=True)
    for a, b in zip(gm, gr):
        assert rel_err(a, b) < 5e-6
    ws = [torch.randn(t.shape, generator=gen) for t in inputs]
    sm = sum((a * w.to(dev)).sum() for a, w in zip(gm, ws))
    sr = sum((b * w.double()).sum() for b, w in zip(gr, ws))
    hm = torch.autograd.grad(sm, mine, allow_unused=True)
    hr = torch.autograd.grad(sr, ref, allow_unused=True)
    for a, b in zip(hm, hr):
        if b is None or float(b.abs().max()) == 0.0:
            assert a is None or float(a.abs().max()) < 1e-6
        else:
            assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("lmax,H", [(4, 128), (2, 32), (3, 16)])
def test_htr_inner_matches_reference_to_second_order(backend, lmax, H):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(lmax)
    E, M = 37, (lmax + 1) ** 2 - 1
    q, k = torch.randn(E, M, H, generator=gen), torch.randn(E, M, H, generator=gen)
    rl = torch.randn(E, M, generator=gen) * 0.6             # NOT unit vectors: exercises the (2 - |r|^2) factor
    rl_dev = rl.to(backend.device)
    _second_order_check(lambda a, b: ops.htr_inner(a, b, rl_dev, lmax),
                        lambda a, b: _htr_ref(a, b, rl.double(), lmax), [q, k], backend.device)


@pytest.mark.parametrize("lmax,mmax,H", [(4, 4, 128), (4, 2, 32), (3, 3, 16), (2, 1, 32)])
def test_gata_value_matches_reference_to_second_order(backend, lmax, mmax, H):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(10 * lmax + mmax)
    E, M = 29, (lmax + 1) ** 2 - 1
    comb = torch.randn(E, (1 + 2 * lmax) * H, generator=gen)
    Xp = torch.randn(E, M, H, generator=gen)
    rl = torch.randn(E, M, generator=gen)
    rl_dev = rl.to(backend.device)
    y = ops.gata_value(comb.to(backend.device), Xp.to(backend.device), rl_dev, lmax, mmax)
    assert y.shape == (E, ops.gata_rows(lmax, mmax), H)
    _second_order_check(lambda a, b: ops.gata_value(a, b, rl_dev, lmax, mmax),
                        lambda a, b: _gata_ref(a, b, rl.double(), lmax, mmax), [comb, Xp], backend.device)
