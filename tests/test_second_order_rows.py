"""Closed-form second-order kernels of the small per-row operators (VERDICT r1 missing #4) against torch autograd in
float64: the derivative OF the first backward pass -- what `loss.backward()` needs when the forces in the loss were
themselves obtained by `autograd.grad(create_graph=True)` (train_MatPES_GATAWandB.py:67-91)."""
import pytest
import torch

from helpers import pkg, rel_err


@pytest.fixture(autouse=True)
def _kernel_path_is_taken(monkeypatch):
    """Every test here must reach the closed-form kernel, not the generic torch route kept for cotangents of parameter
    gradients: record the C-ABI entry points that were called."""
    _lib = pkg("_lib")
    calls = []
    orig = _lib.call

    def spy(name, *a, **k):
        calls.append(name)
        return orig(name, *a, **k)

    monkeypatch.setattr(_lib, "call", spy)
    yield calls
    assert any(n.endswith("_bwd2") for n in calls), sorted(set(calls))


@pytest.mark.parametrize("rows,width", [(37, 128), (5, 16), (64, 200)])
def test_ln_silu_double_backward(backend, rows, width):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(rows + width)
    x = torch.randn(rows, width, generator=gen) * 1.5 + 0.3
    w = torch.randn(width, generator=gen) * 0.5 + 1.0
    b = torch.randn(width, generator=gen) * 0.2
    go = torch.randn(rows, width, generator=gen)
    uu = torch.randn(rows, width, generator=gen)
    Fn = torch.nn.functional

    ref = [t.double().requires_grad_(True) for t in (x, w, b)]
    yr = Fn.silu(Fn.layer_norm(ref[0], (width,), ref[1], ref[2], 1e-5))
    gor = go.double().requires_grad_(True)
    (gxr,) = torch.autograd.grad(yr, ref[0], gor, create_graph=True)
    hr = torch.autograd.grad((gxr * uu.double()).sum(), ref + [gor])

    dev = backend.device
    mine = [t.clone().to(dev).requires_grad_(True) for t in (x, w, b)]
    ym = ops.ln_silu(mine[0], mine[1], mine[2], 1e-5)
    assert rel_err(ym, yr) < 2e-6
    gom = go.clone().to(dev).requires_grad_(True)
    (gxm,) = torch.autograd.grad(ym, mine[0], gom, create_graph=True)
    assert rel_err(gxm, gxr) < 5e-6
    hm = torch.autograd.grad((gxm * uu.to(dev)).sum(), mine + [gom])
    for a, r, name in zip(hm, hr, ("x", "weight", "bias", "grad_out")):
        assert rel_err(a, r) < 2e-5, name


@pytest.mark.parametrize("norm_type", ["rms_norm_sh", "layer_norm_sh", "layer_norm"])
@pytest.mark.parametrize("lmax,C", [(4, 32), (2, 16)])
def test_equiv_norm_double_backward(backend, norm_type, lmax, C):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(lmax * 100 + C)
    N, K = 7, (lmax + 1) ** 2
    x = torch.randn(N, K, C, generator=gen) + 0.2
    w = torch.randn(lmax + 1, C, generator=gen) * 0.3 + 1.0
    b = torch.randn(C, generator=gen) * 0.1
    go = torch.randn(N, K, C, generator=gen)
    uu = torch.randn(N, K, C, generator=gen)
    expr = ops._equiv_norm_expr(norm_type, lmax, 1e-5)

    ref = [t.double().requires_grad_(True) for t in (x, w, b)]
    yr = expr(*ref)
    gor = go.double().requires_grad_(True)
    (gxr,) = torch.autograd.grad(yr, ref[0], gor, create_graph=True)
    hr = torch.autograd.grad((gxr * uu.double()).sum(), [ref[0], ref[1], gor])

    dev = backend.device
    mine = [t.clone().to(dev).requires_grad_(True) for t in (x, w, b)]
    ym = ops.equiv_norm(mine[0], mine[1], mine[2], norm_type, lmax, 1e-5)
    assert rel_err(ym, yr) < 2e-6
    gom = go.clone().to(dev).requires_grad_(True)
    (gxm,) = torch.autograd.grad(ym, mine[0], gom, create_graph=True)
    assert rel_err(gxm, gxr) < 5e-6
    hm = torch.autograd.grad((gxm * uu.to(dev)).sum(), [mine[0], mine[1], gom])
    for a, r, name in zip(hm, hr, ("x", "weight", "grad_out")):
        assert rel_err(a, r) < 2e-5, name


@pytest.mark.parametrize("use_ln", [True, False])
@pytest.mark.parametrize("heads,ach", [(8, 64), (2, 8), (4, 32)])
def test_attn_alpha_double_backward(backend, use_ln, heads, ach):
    """alpha = segment_softmax_dst(alpha_dot . SmoothLeakyReLU(LayerNorm(Ya)))  (transformer_block.py:311-315): the
    derivative of dL/dYa w.r.t. (Ya, LN weight, LN bias, alpha_dot, grad_out), kernels vs torch autograd in float64."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(heads * 100 + ach + int(use_ln))
    N, E = 9, 61
    ei = torch.randint(0, N, (2, E), generator=gen)
    Ya = torch.randn(E, heads * ach, generator=gen)
    ln_w = (torch.randn(ach, generator=gen) * 0.3 + 1.0) if use_ln else None
    ln_b = (torch.randn(ach, generator=gen) * 0.2) if use_ln else None
    ad = torch.randn(heads, ach, generator=gen) * 0.3
    go = torch.randn(E, heads, generator=gen)
    uu = torch.randn(E, heads * ach, generator=gen)

    expr = ops._alpha_expr(heads, ach, 1e-5, ei[1], N)
    leaves = [Ya, ln_w, ln_b, ad]
    ref = [t.double().requires_grad_(True) if t is not None else None for t in leaves]
    yr = expr(*ref)
    gor = go.double().requires_grad_(True)
    (gxr,) = torch.autograd.grad(yr, ref[0], gor, create_graph=True)
    live_r = [t for t in ref if t is not None] + [gor]
    hr = torch.autograd.grad((gxr * uu.double()).sum(), live_r)

    dev = backend.device
    plan = ops.edge_plan(ei.to(dev), N)
    mine = [t.clone().to(dev).requires_grad_(True) if t is not None else None for t in leaves]
    ym = ops.attn_alpha(mine[0], mine[1], mine[2], mine[3], plan, heads, ach)
    assert rel_err(ym, yr) < 5e-6
    gom = go.clone().to(dev).requires_grad_(True)
    (gxm,) = torch.autograd.grad(ym, mine[0], gom, create_graph=True)
    assert rel_err(gxm, gxr) < 1e-5
    live_m = [t for t in mine if t is not None] + [gom]
    hm = torch.autograd.grad((gxm * uu.to(dev)).sum(), live_m)
    names = (["Ya", "ln_w", "ln_b", "alpha_dot"] if use_ln else ["Ya", "alpha_dot"]) + ["grad_out"]
    for a, r, name in zip(hm, hr, names):
        assert rel_err(a, r) < 5e-5, name


def test_rbf_double_backward(backend):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(5)
    E, R, cutoff = 83, 600, 6.0
    d = torch.rand(E, generator=gen) * cutoff
    offset = torch.linspace(0.0, cutoff, R)
    coeff = -0.5 / (2.0 * float(offset[1] - offset[0])) ** 2
    go = torch.randn(E, R, generator=gen)
    uu = torch.randn(E, generator=gen)

    dr = d.double().requires_grad_(True)
    yr = torch.exp(coeff * (dr.view(-1, 1) - offset.double().view(1, -1)) ** 2)
    gor = go.double().requires_grad_(True)
    (gdr,) = torch.autograd.grad(yr, dr, gor, create_graph=True)
    hr = torch.autograd.grad((gdr * uu.double()).sum(), [dr, gor])

    dev = backend.device
    dm = d.clone().to(dev).requires_grad_(True)
    ym = ops.rbf(dm, offset.to(dev), coeff)
    assert rel_err(ym, yr) < 2e-6
    gom = go.clone().to(dev).requires_grad_(True)
    (gdm,) = torch.autograd.grad(ym, dm, gom, create_graph=True)
    assert rel_err(gdm, gdr) < 1e-5
    hm = torch.autograd.grad((gdm * uu.to(dev)).sum(), [dm, gom])
    assert rel_err(hm[0], hr[0]) < 2e-5 and rel_err(hm[1], hr[1]) < 2e-5
