"""The five global (all-to-all) node attention classes of reference NewFunctions/GATA_and_all2all/activation.py:419-1567
 against the UNMODIFIED reference classes evaluated on the host CPU in the same test: same state_dict
(strict), same inputs -- three structures of different sizes in one batch -- outputs <= 1e-5, parameter gradients <= 1e-4
of each tensor's largest entry, position gradients (the HTR variants differentiate the pair harmonics) <= 1e-4."""
import importlib
import os
import sys

import pytest
import torch

from helpers import pkg, rel_err
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not os.path.isdir(ref_loader.REF_ROOT), reason="reference copy missing (oracle/make_ref.py)")

C, LMAX, HEADS = 16, 2, 4
VARIANTS = {
    "GlobalNodeAttention": (dict(d_model=C, num_heads=HEADS, dropout=0.0, use_rope=True, rope_dim=6), "scalar"),
    "GlobalNodeAttentionFullEquivariant": (dict(sphere_channels=C, lmax=LMAX, num_heads=HEADS, dropout=0.0), "nopos"),
    "GlobalNodeAttentionHTR": (dict(sphere_channels=C, lmax=LMAX, num_heads=HEADS, dropout=0.0), "full"),
    "GlobalNodeAttentionHTR_with_distance": (dict(sphere_channels=C, lmax=LMAX, num_heads=HEADS, dropout=0.0, num_rbf=8,
                                                  rbf_cutoff=6.0), "full"),
    "GlobalNodeAttentionHTR_with_ROPE": (dict(sphere_channels=C, lmax=LMAX, num_heads=HEADS, dropout=0.0, num_rbf=8,
                                              rbf_cutoff=6.0, use_rope=True, rope_dim=6), "full"),
}


def _reference_class(name):
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("EquiformerV2Functions", "NewFunctions")}
    ref_loader.install()
    try:
        mod = importlib.import_module("NewFunctions.GATA_and_all2all.activation")
        assert (getattr(mod, "__file__", "") or "").startswith(ref_loader.REF_ROOT)
        return getattr(mod, name)
    finally:
        for k in list(sys.modules):
            if k.split(".")[0] in ("EquiformerV2Functions", "NewFunctions"):
                del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_matches_reference_class(backend, name):
    kw, kind = VARIANTS[name]
    gen = torch.Generator().manual_seed(len(name))
    counts = [5, 9, 3]
    N, K = sum(counts), (LMAX + 1) ** 2
    batch = torch.repeat_interleave(torch.arange(3), torch.tensor(counts))
    pos = torch.randn(N, 3, generator=gen) * 2.0
    x = torch.randn(N, K, C, generator=gen) if kind != "scalar" else torch.randn(N, C, generator=gen)
    go = torch.randn(x.shape, generator=gen)

    torch.manual_seed(0)
    ref = _reference_class(name)(**kw)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=gen))
    pr = pos.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr, batch) if kind == "nopos" else ref(xr, batch, pr)
    (yr * go).sum().backward()

    mine = getattr(pkg("NewFunctions.GATA_and_all2all.activation"), name)(**kw).to(backend.device)
    mine.load_state_dict(ref.state_dict(), strict=True)
    pm = pos.clone().to(backend.device).requires_grad_(True)
    xm = x.clone().to(backend.device).requires_grad_(True)
    bm = batch.to(backend.device)
    ym = mine(xm, bm) if kind == "nopos" else mine(xm, bm, pm)
    assert rel_err(ym, yr) < 1e-5
    (ym * go.to(backend.device)).sum().backward()
    assert rel_err(xm.grad, xr.grad) < 1e-4
    if kind != "nopos" and pr.grad is not None and float(pr.grad.abs().max()) > 0:
        assert rel_err(pm.grad, pr.grad) < 1e-4
    rp = dict(ref.named_parameters())
    # gradients that are mathematically zero (softmax is invariant to a key bias) are rounding noise in both
    # implementations: tolerance = 1e-4 of the tensor's own scale + 2e-6 of the largest gradient of the module
    top = max(float(p.grad.abs().max()) for p in rp.values() if p.grad is not None)
    for k, p in mine.named_parameters():
        g_ref = rp[k].grad
        if g_ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert float((p.grad.cpu() - g_ref).abs().max()) <= 1e-4 * float(g_ref.abs().max()) + 2e-6 * top, k
