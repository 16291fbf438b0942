"""Fused GaussianSmearing + first radial-MLP layer (north_star piece 2) against the dense evaluation the
reference performs (equiformerv2_oc20.py:43-60 -> torch.cat of transformer_block.py:241-248 -> nn.Linear of
radial_function.py:29): forward <= 1e-6, weight / bias / embedding gradients <= 1e-5 of each tensor's largest entry."""
import pytest
import torch

from helpers import pkg, rel_err


@pytest.mark.parametrize("R,cutoff,H,Ce,V", [(600, 12.0, 128, 128, 90), (600, 5.0, 64, 64, 10), (50, 6.0, 32, 16, 7)])
def test_rbf_linear_matches_dense(backend, R, cutoff, H, Ce, V):
    ops = pkg("ops")
    common = pkg("models.common")
    gen = torch.Generator().manual_seed(R + H)
    E, N = 700, 60
    smear = common.GaussianSmearing(0.0, cutoff, R, 2.0).to(backend.device)
    dist = (torch.rand(E, generator=gen) * cutoff * 0.98 + 0.01)
    dist[:4] = torch.tensor([0.0, cutoff, cutoff * 0.5, 1e-4])          # both ends of the basis
    ei = torch.randint(0, N, (2, E), generator=gen)
    Z = torch.randint(0, V, (N,), generator=gen)
    W1 = (torch.randn(H, R + 2 * Ce, generator=gen) / (R + 2 * Ce) ** 0.5)
    b1 = torch.randn(H, generator=gen) * 0.1
    src_w = torch.randn(V, Ce, generator=gen) * 0.5
    dst_w = torch.randn(V, Ce, generator=gen) * 0.5
    go = torch.randn(E, H, generator=gen)

    # dense reference in float64 on the host
    d64 = dist.double()
    rbf = torch.exp(smear.coeff * (d64.view(-1, 1) - smear.offset.cpu().double().view(1, -1)) ** 2)
    leaves = [t.double().requires_grad_(True) for t in (W1, b1, src_w, dst_w)]
    x = torch.cat([rbf, leaves[2][Z[ei[0]]], leaves[3][Z[ei[1]]]], dim=1)
    ref = x @ leaves[0].t() + leaves[1]
    (ref * go.double()).sum().backward()

    dev = backend.device
    mine = [t.clone().to(dev).requires_grad_(True) for t in (W1, b1, src_w, dst_w)]
    rbf_t = smear(dist.to(dev))
    assert getattr(rbf_t, "_eqv2_rbf", None) is not None and rbf_t._eqv2_rbf.band <= 16
    plan = ops.edge_plan(ei.to(dev), N)
    zs, csr_s, zd, csr_d = plan.element_types(Z.to(dev), V)
    src = ops.rbf_source_of(rbf_t)
    assert src is not None
    feat = ops.FusedEdgeFeatures(src, mine[2], mine[3], zs, csr_s, zd, csr_d)
    assert feat.width == R + 2 * Ce
    out = ops.rbf_linear(feat, mine[0], mine[1])
    assert rel_err(out, ref) < 1e-6
    (out * go.to(dev)).sum().backward()
    for a, b, name in zip(mine, leaves, ("W1", "b1", "source_embedding", "target_embedding")):
        assert rel_err(a.grad, b.grad) < 1e-5, name


def test_fused_first_layer_is_used_by_the_model_and_keeps_parity(backend):
    """The OC20 model takes the fused path by default (first-order step): same golden parity, and the dense
    [E, 600] x_edge GEMM is gone from the launch list."""
    from conftest import golden
    from helpers import build_oc20, fixed_rand_like, load_params
    ops, _lib = pkg("ops"), pkg("_lib")
    fx = golden("oc20_small_rms_norm_sh.pt")
    model = build_oc20(fx["hyper"], backend.device)
    load_params(model, fx["params"])
    data = backend.to(dict(fx["inputs"], edge_index=fx["edge_index"], edge_distance=fx["edge_distance"],
                           edge_distance_vec=fx["edge_vec"]))
    n0 = _lib.launch_count()
    calls = []
    orig = _lib.call

    def spy(name, *a, **k):
        calls.append(name)
        return orig(name, *a, **k)

    _lib.call = spy
    try:
        with fixed_rand_like(fx["rand_vec"] + 0.5):
            energy, forces = model(data)
        w = torch.linspace(-1, 1, forces.numel(), device=forces.device).view_as(forces)
        (energy.sum() + (forces * w).sum()).backward()
    finally:
        _lib.call = orig
    assert rel_err(energy, fx["energy"]) < 1e-5 and rel_err(forces, fx["forces"]) < 1e-5
    assert calls.count("eqv2_rbf_linear_fwd") == fx["hyper"]["num_layers"] + 2       # blocks + force head + degree embedding
    assert calls.count("eqv2_rbf_linear_wgrad") == fx["hyper"]["num_layers"] + 2
    for k, p in model.named_parameters():
        if "rad_func.net.0" in k or "source_embedding" in k or "target_embedding" in k:
            g_ref = fx["grads"][k]
            assert float((p.grad.cpu() - g_ref).abs().max()) <= 5e-5 * float(g_ref.abs().max()), k
