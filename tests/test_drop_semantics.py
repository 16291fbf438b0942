"""Stochastic-depth / dropout mask semantics against the UNMODIFIED reference drop.py (VERDICT r1:
"every parity test runs rates 0").  The RNG stays in torch on purpose: with the same seed the drop-in layers must draw
the same masks and produce bit-identical outputs as the reference layers (reference drop.py:16-68,119-149), in training
mode, in eval mode, with and without the announced number of graphs (the drop-in's capturable replacement of the
reference's `batch.max() + 1` read-back)."""
import importlib
import os
import sys

import pytest
import torch

from helpers import pkg
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not os.path.isdir(ref_loader.REF_ROOT), reason="reference copy missing")


@pytest.fixture
def ref_drop():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] == "EquiformerV2Functions"}
    ref_loader.install()
    mod = importlib.import_module("EquiformerV2Functions.drop")
    assert mod.__file__.startswith(ref_loader.REF_ROOT)
    yield mod
    for k in list(sys.modules):
        if k.split(".")[0] == "EquiformerV2Functions":
            del sys.modules[k]
    sys.modules.update(saved)


def _batch():
    return torch.repeat_interleave(torch.arange(5), torch.tensor([3, 7, 2, 5, 4]))


@pytest.mark.parametrize("announce", [False, True])
@pytest.mark.parametrize("p", [0.05, 0.5])
def test_graph_drop_path_matches_reference(ref_drop, p, announce):
    mine = pkg("EquiformerV2Functions.drop")
    batch = _batch()
    x = torch.randn(len(batch), 9, 8, generator=torch.Generator().manual_seed(1))
    a, b = ref_drop.GraphDropPath(p), mine.GraphDropPath(p)
    for training in (True, False):
        a.train(training), b.train(training)
        for seed in range(4):
            torch.manual_seed(seed)
            ya = a(x, batch)
            torch.manual_seed(seed)
            if announce:
                mine.set_num_graphs(5)
            try:
                yb = b(x, batch)
            finally:
                mine.set_num_graphs(None)
            assert torch.equal(ya, yb)
            if not training:
                assert torch.equal(yb, x)
    # the masks are per graph: all atoms of a structure share one factor in {0, 1/(1-p)}
    a.train(True)
    torch.manual_seed(11)
    y = a(torch.ones(len(batch), 1, 1), batch).view(-1)
    for g in range(5):
        vals = y[batch == g].unique()
        assert len(vals) == 1 and (float(vals) == 0.0 or abs(float(vals) - 1 / (1 - p)) < 1e-6)


@pytest.mark.parametrize("drop_graph", [False, True])
def test_equivariant_dropout_matches_reference(ref_drop, drop_graph):
    mine = pkg("EquiformerV2Functions.drop")
    batch = _batch()
    x = torch.randn(len(batch), 9, 8, generator=torch.Generator().manual_seed(2))
    a = ref_drop.EquivariantDropoutArraySphericalHarmonics(0.3, drop_graph)
    b = mine.EquivariantDropoutArraySphericalHarmonics(0.3, drop_graph)
    for training in (True, False):
        a.train(training), b.train(training)
        for seed in range(4):
            torch.manual_seed(seed)
            ya = a(x, batch)
            torch.manual_seed(seed)
            yb = b(x, batch)
            assert torch.equal(ya, yb)
    # one mask per (node | graph, channel), shared by every (l, m) row
    a.train(True)
    torch.manual_seed(3)
    y = a(torch.ones_like(x), batch)
    assert torch.equal(y, y[:, :1, :].expand_as(y))


def test_drop_path_function_matches_reference(ref_drop):
    mine = pkg("EquiformerV2Functions.drop")
    x = torch.randn(12, 4, 4, generator=torch.Generator().manual_seed(3))
    for p in (0.0, 0.1, 0.9):
        for training in (True, False):
            torch.manual_seed(5)
            ya = ref_drop.drop_path(x, p, training)
            torch.manual_seed(5)
            yb = mine.drop_path(x, p, training)
            assert torch.equal(ya, yb)


def test_whole_model_train_mode_matches_reference_with_same_seed(backend):
    """The OC20 model in TRAINING mode with all three regularisers on (attention dropout, stochastic depth, projection
    dropout): same seed => same masks => same energies / forces as the unmodified reference (CPU generator; the kernel
    source runs under the emulator).  Pins the RNG consumption order and the layout the attention-dropout mask is drawn on."""
    if backend.name != "emu":
        pytest.skip("CPU generator semantics: compared on the CPU only (the CUDA generator draws different numbers)")
    from conftest import golden
    from helpers import fixed_rand_like, load_params, rel_err
    fx = golden("oc20_small_rms_norm_sh.pt")
    hp = fx["hyper"]
    kw = dict(max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
              num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
              attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"],
              ffn_hidden_channels=hp["ffn_hidden"], norm_type=hp["norm_type"], lmax_list=[hp["lmax"]],
              mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"], alpha_drop=0.3,
              drop_path_rate=0.2, proj_drop=0.1)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] == "EquiformerV2Functions"}
    ref_loader.install()
    try:
        mod = importlib.import_module("equiformerv2_oc20")
        ref_model = mod.EquiformerV2_OC20(**kw).train()
        with torch.no_grad():
            for k, p in ref_model.named_parameters():
                p.copy_(fx["params"][k])
        with fixed_rand_like(fx["rand_vec"] + 0.5):
            torch.manual_seed(77)
            e_ref, f_ref = ref_model(dict(fx["inputs"]))
    finally:
        for k in list(sys.modules):
            if k.split(".")[0] == "EquiformerV2Functions" or k == "equiformerv2_oc20":
                del sys.modules[k]
        sys.modules.update(saved)
    model = pkg("models.equiformerv2_oc20").EquiformerV2_OC20(**kw).train()
    load_params(model, fx["params"])
    data = dict(fx["inputs"], edge_index=fx["edge_index"], edge_distance=fx["edge_distance"],
                edge_distance_vec=fx["edge_vec"])
    with fixed_rand_like(fx["rand_vec"] + 0.5):
        torch.manual_seed(77)
        e, f = model(data)
    assert rel_err(e, e_ref) < 1e-5 and rel_err(f, f_ref) < 1e-5
    # and the masks matter: another seed gives a different result
    with fixed_rand_like(fx["rand_vec"] + 0.5):
        torch.manual_seed(78)
        e2, _ = model(data)
    assert rel_err(e2, e_ref) > 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("p", [0.05, 0.5])
def test_fused_graph_drop_path_kernel_is_bit_identical_to_the_reference_expression(p):
    """eqv2_drop_path_scale (one launch) against the reference's expression `x * (ones.div(keep) * floor(keep + rand))[batch]`
    (drop.py:16-27,49-68) with the same seed on the GPU: outputs and input gradients bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mine = pkg("EquiformerV2Functions.drop")
    pkg("_lib")._state["lib"] = None
    batch = _batch().cuda()
    x0 = torch.randn(len(batch), 9, 8, generator=torch.Generator().manual_seed(1)).cuda()
    g0 = torch.randn(len(batch), 9, 8, generator=torch.Generator().manual_seed(2)).cuda()
    mod = mine.GraphDropPath(p).train(True)
    for seed in range(4):
        xa, xb = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        torch.manual_seed(seed)
        ones = torch.ones((5, 1, 1), device="cuda")
        ya = xa * mine.drop_path(ones, p, True)[batch]                 # the reference expression (torch ops)
        torch.manual_seed(seed)
        mine.set_num_graphs(5)
        try:
            yb = mod(xb, batch)                                        # the kernel
        finally:
            mine.set_num_graphs(None)
        assert torch.equal(ya, yb)
        ya.backward(g0), yb.backward(g0)
        assert torch.equal(xa.grad, xb.grad)
