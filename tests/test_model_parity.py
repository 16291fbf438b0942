"""End-to-end parity of the CUDA path against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py): same parameters (by the reference's own state_dict keys), same graph, same
edge-frame draw.  Tolerance: 1e-5 relative on outputs (north_star, fp32 mode), 5e-5 relative (to the
largest entry of each tensor) on parameter gradients (5e-5 since round 2; measured 9e-6)."""
import pytest
import torch

from conftest import golden
from helpers import build_oc20, build_qm9, fixed_rand_like, load_params, rel_err

OUT_TOL = 1e-5
GRAD_TOL = 5e-5


def _graph_inputs(fx, backend):
    data = dict(fx["inputs"])
    data["edge_index"] = fx["edge_index"]
    data["edge_distance"] = fx["edge_distance"]
    data["edge_distance_vec"] = fx["edge_vec"]
    return backend.to(data)


def _check_grads(model, fx):
    bad = []
    # gradients that are mathematically zero (e.g. a key bias under softmax) are rounding noise in both paths:
    # measure them against the overall gradient scale, not against themselves
    floor = 1e-7 * max(float(g.abs().max()) for g in fx["grads"].values())
    # softmax is invariant to a bias on the keys: d/d(k_proj.bias) is EXACTLY zero mathematically and pure rounding
    # noise (1e-8) numerically in both implementations
    zero_by_construction = ("global_attn.k_proj.bias",)
    for k, p in model.named_parameters():
        g_ref = fx["grads"].get(k)
        if g_ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        if k.endswith(zero_by_construction):
            assert float(p.grad.abs().max()) < 1e-5 * floor / 1e-7 * 1e-2, k
            continue
        e = float((p.grad.detach().double().cpu() - g_ref.double()).abs().max() / max(float(g_ref.abs().max()), floor))
        if e > GRAD_TOL:
            bad.append((k, e))
    assert not bad, bad


@pytest.mark.parametrize("norm_type", ["rms_norm_sh", "layer_norm_sh", "layer_norm"])
def test_oc20_small_matches_reference(backend, norm_type):
    fx = golden(f"oc20_small_{norm_type}.pt")
    model = build_oc20(fx["hyper"], backend.device)
    load_params(model, fx["params"])
    data = _graph_inputs(fx, backend)
    with fixed_rand_like(fx["rand_vec"] + 0.5):
        energy, forces = model(data)
    assert rel_err(energy, fx["energy"]) < OUT_TOL
    assert rel_err(forces, fx["forces"]) < OUT_TOL
    w = torch.linspace(-1, 1, forces.numel(), device=forces.device).view_as(forces)
    (energy.sum() + (forces * w).sum()).backward()
    _check_grads(model, fx)


def test_qm9_small_matches_reference(backend):
    fx = golden("qm9_small.pt")
    model = build_qm9(fx["hyper"], backend.device)
    load_params(model, fx["params"])
    data = _graph_inputs(fx, backend)
    with fixed_rand_like(fx["rand_vec"] + 0.5):
        pred = model(data)
    assert rel_err(pred, fx["pred"]) < OUT_TOL
    w = torch.linspace(-1, 1, pred.numel(), device=pred.device).view_as(pred)
    (pred * w).sum().backward()
    _check_grads(model, fx)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("tf32x3", 1e-5), ("tf32", 5e-3)])
def test_oc20_small_other_gemm_modes(mode, tol):
    """fp32 = exact FFMA engine everywhere; tf32 = single-pass TF32 tensor cores (stated tolerance 5e-3)."""
    from conftest import Backend
    from helpers import pkg
    ops = pkg("ops")
    be = Backend("cuda")
    fx = golden("oc20_small_rms_norm_sh.pt")
    ops.set_gemm_mode(mode)
    try:
        model = build_oc20(fx["hyper"], be.device)
        load_params(model, fx["params"])
        data = _graph_inputs(fx, be)
        with fixed_rand_like(fx["rand_vec"] + 0.5):
            energy, forces = model(data)
        assert rel_err(energy, fx["energy"]) < tol
        assert rel_err(forces, fx["forces"]) < tol
    finally:
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)


@pytest.mark.gpu
@pytest.mark.parametrize("variant,fixture", [("oc20", "oc20_small_rms_norm_sh.pt"), ("gatav2", "matpes_gatav2_small.pt"),
                                             ("gatav2_phi", "matpes_gatav2_phi_small.pt")])
def test_single_pass_f16_mode_stated_tolerance(variant, fixture):
    """Reduced-precision GEMM mode of BASELINE configs[3] ("bf16/TF32 GEMM mode vs fp32 tolerance check"): `f16` = the
    f16x3 engine reading only the hi planes -- ONE fp16 tensor-core pass (11-bit significands, fp32 accumulation), every
    contraction forced onto it (threshold 0).  STATED TOLERANCE against the fp32 golden vectors of the unmodified
    reference: 1e-2 relative on energies and forces (measured on B200: OC20 7.7e-5 / 8.7e-4, GATAV2 3.0e-4 / 2.7e-4,
    GATAV2-phi 4.6e-3 / 1.1e-3)."""
    import helpers
    from conftest import Backend
    from helpers import pkg
    ops, _lib = pkg("ops"), pkg("_lib")
    be = Backend("cuda")
    fx = golden(fixture)
    old = ops.F16_MIN_MACS
    ops.F16_MIN_MACS = 0
    ops.set_gemm_mode("f16")
    try:
        _lib.start_kernel_timing()
        if variant == "oc20":
            model = build_oc20(fx["hyper"], be.device)
            load_params(model, fx["params"])
            with fixed_rand_like(fx["rand_vec"] + 0.5):
                energy, forces = model(_graph_inputs(fx, be))
        else:
            model = getattr(helpers, "build_" + variant)(fx["hyper"], be.device)
            load_params(model, fx["params"])
            data = be.to(dict(fx["inputs"]))
            pos = data["pos"].clone().requires_grad_(True)
            out = model(dict(data, pos=pos))
            energy = out["energy"]
            forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
            forces.sum().backward()         # the double backward must run in this mode too
        prof = _lib.stop_kernel_timing()
        e_err, f_err = rel_err(energy, fx["energy"]), rel_err(forces, fx["forces"])
        print(f"f16 single pass [{variant}]: energy {e_err:.2e} forces {f_err:.2e}")
        assert prof.get("eqv2_gemm_f16_ex", {}).get("calls", 0) >= 10, sorted(prof)
        assert e_err < 1e-2 and f_err < 1e-2
        assert e_err > 1e-7 or f_err > 1e-7         # it really is the reduced-precision engine
    finally:
        ops.F16_MIN_MACS = old
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["oc20_small_rms_norm_sh.pt", "qm9_small.pt"])
def test_small_fixtures_with_every_contraction_on_the_f16x3_engine(fixture):
    """The golden fixtures are small: by default their contractions fall below the tensor-core engine's size threshold
    and run on the FFMA engine.  Here the threshold is 0, so every addressable GEMM (forward, dgrad, wgrad, degree slabs)
    goes through operand split + tcgen05 kind::f16 and is compared with the unmodified reference directly."""
    from conftest import Backend
    from helpers import pkg
    ops, _lib = pkg("ops"), pkg("_lib")
    be = Backend("cuda")
    fx = golden(fixture)
    old = ops.F16_MIN_MACS
    ops.F16_MIN_MACS = 0
    ops.set_gemm_mode("f16x3")
    try:
        _lib.start_kernel_timing()
        if fixture.startswith("oc20"):
            model = build_oc20(fx["hyper"], be.device)
            load_params(model, fx["params"])
            with fixed_rand_like(fx["rand_vec"] + 0.5):
                energy, forces = model(_graph_inputs(fx, be))
            assert rel_err(energy, fx["energy"]) < OUT_TOL and rel_err(forces, fx["forces"]) < OUT_TOL
            w = torch.linspace(-1, 1, forces.numel(), device=forces.device).view_as(forces)
            (energy.sum() + (forces * w).sum()).backward()
        else:
            model = build_qm9(fx["hyper"], be.device)
            load_params(model, fx["params"])
            with fixed_rand_like(fx["rand_vec"] + 0.5):
                pred = model(_graph_inputs(fx, be))
            assert rel_err(pred, fx["pred"]) < OUT_TOL
            (pred * torch.linspace(-1, 1, pred.numel(), device=pred.device).view_as(pred)).sum().backward()
        prof = _lib.stop_kernel_timing()
        assert prof.get("eqv2_gemm_f16", {}).get("calls", 0) >= 10, sorted(prof)
        _check_grads(model, fx)
    finally:
        ops.F16_MIN_MACS = old
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)


@pytest.mark.gpu
@pytest.mark.parametrize("variant,fixture", [("matpes_v2", "matpes_v2_small.pt"), ("gatav2", "matpes_gatav2_small.pt"),
                                             ("gatav2_phi", "matpes_gatav2_phi_small.pt"),
                                             ("gatav2_global", "matpes_gatav2_global_small.pt")])
def test_matpes_family_double_backward_with_every_contraction_on_the_f16x3_engine(variant, fixture):
    """Same as above for the MatPES pattern (configs 3-5): energy, autograd forces and the double-backward parameter
    gradients of the unmodified reference, with every GEMM of all three passes on the f16x3 engine."""
    import helpers
    from conftest import Backend
    from helpers import pkg
    ops, _lib = pkg("ops"), pkg("_lib")
    be = Backend("cuda")
    fx = golden(fixture)
    old = ops.F16_MIN_MACS
    ops.F16_MIN_MACS = 0
    ops.set_gemm_mode("f16x3")
    try:
        model = getattr(helpers, "build_" + variant)(fx["hyper"], be.device)
        load_params(model, fx["params"])
        data = be.to(dict(fx["inputs"]))
        pos = data["pos"].clone().requires_grad_(True)
        _lib.start_kernel_timing()
        out = model(dict(data, pos=pos))
        assert rel_err(out["energy"], fx["energy"]) < OUT_TOL
        forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
        assert rel_err(forces, fx["forces"]) < OUT_TOL
        wf = torch.linspace(-1, 1, forces.numel(), device=forces.device).view_as(forces)
        we = torch.linspace(0.5, 1.5, out["energy"].numel(), device=forces.device).view_as(out["energy"])
        ((out["energy"] * we).sum() + (forces * wf).sum()).backward()
        prof = _lib.stop_kernel_timing()
        assert prof.get("eqv2_gemm_f16", {}).get("calls", 0) >= 30, sorted(prof)
        _check_grads(model, fx)
    finally:
        ops.F16_MIN_MACS = old
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)


def test_matpes_v2_train_step_matches_reference(backend):
    """MatPES pattern (train_MatPES_GATAWandB.py:67-91): energy, forces = -autograd.grad(E, pos, create_graph=True),
    then the gradient of a loss on energy AND forces w.r.t. every parameter (double backward), against the golden
    vectors of the unmodified reference.  The graph is built by the CUDA 27-image builder."""
    from helpers import build_matpes_v2
    fx = golden("matpes_v2_small.pt")
    model = build_matpes_v2(fx["hyper"], backend.device)
    load_params(model, fx["params"])
    data = backend.to(dict(fx["inputs"]))
    pos = data["pos"].clone().requires_grad_(True)
    out = model(dict(data, pos=pos))
    assert rel_err(out["energy"], fx["energy"]) < OUT_TOL
    forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
    assert rel_err(forces, fx["forces"]) < OUT_TOL
    wf = torch.linspace(-1, 1, forces.numel(), device=forces.device).view_as(forces)
    we = torch.linspace(0.5, 1.5, out["energy"].numel(), device=forces.device).view_as(out["energy"])
    ((out["energy"] * we).sum() + (forces * wf).sum()).backward()
    _check_grads(model, fx)


def test_matpes_v1_forces_and_stress_match_reference(backend):
    """BASELINE config 3 as named (equiformerv2_MatPES.py:373-488): forces and Voigt stress computed inside forward by
    autograd (strain applied to positions and cell), graph from the CUDA 27-image builder (version 1).  The reference
    can only run the two passes separately; each is compared with its own golden output, and the
    combined call -- which the reference cannot make -- must reproduce both.
    Tolerance: the reference's fp32 forces / stress are themselves 1.2e-5 / 1.6e-5 away from the same reference evaluated
    in float64 (fixture fields *_f64), so the check is made against the float64 values with a bound of twice the
    reference's own fp32 deviation (never tighter than the 1e-5 of north_star), plus a 5e-5 sanity bound on the
    fp32-vs-fp32 difference."""
    from helpers import build_matpes_v1
    fx = golden("matpes_v1_small.pt")
    data = backend.to(dict(fx["inputs"]))
    draw = fx["rand_vec"] + 0.5
    tol_f = max(1e-5, 2 * rel_err(fx["forces"], fx["forces_f64"]))
    tol_s = max(1e-5, 2 * rel_err(fx["stress"], fx["stress_f64"]))

    def check(out, forces, stress):
        if forces:
            assert rel_err(out["forces"], fx["forces_f64"]) < tol_f
            assert rel_err(out["forces"], fx["forces"]) < 5e-5
        if stress:
            assert rel_err(out["stress"], fx["stress_f64"]) < tol_s
            assert rel_err(out["stress"], fx["stress"]) < 5e-5

    mf = build_matpes_v1(fx["hyper"], backend.device, regress_forces=True, regress_stress=False)
    load_params(mf, fx["params"])
    with fixed_rand_like(draw):
        of = mf(dict(data, pos=data["pos"].clone()))
    assert rel_err(of["energy"], fx["energy"]) < OUT_TOL and rel_err(of["energy"], fx["energy_f64"]) < OUT_TOL
    check(of, True, False)
    ms = build_matpes_v1(fx["hyper"], backend.device, regress_forces=False, regress_stress=True).eval()
    load_params(ms, fx["params"])
    with fixed_rand_like(draw):
        os_ = ms(dict(data, pos=data["pos"].clone()))
    assert rel_err(os_["energy"], fx["energy_stress_pass"]) < OUT_TOL
    check(os_, False, True)
    mb = build_matpes_v1(fx["hyper"], backend.device).eval()
    load_params(mb, fx["params"])
    with fixed_rand_like(draw):
        ob = mb(dict(data, pos=data["pos"].clone()))
    check(ob, True, True)


@pytest.mark.parametrize("variant", ["gatav2", "gatav2_phi", "gatav2_global"])
def test_matpes_gatav2_train_step_matches_reference(backend, variant):
    """BASELINE config 4 family (equiformerv2_MatPES_GATAV2.py: HTR edge stream + GATA value activation): energy,
    autograd forces and double-backward parameter gradients against the unmodified reference; parameters the
    reference leaves without gradient (so2_conv_1.so2_m_conv.*) must stay without gradient."""
    import helpers
    if backend.name == "emu" and variant != "gatav2_phi":
        pytest.skip("CPU suite budget: the emulator runs the phi variant (superset of the base one; the global-attention "
                    "classes have their own emulator test, tests/test_global_attention.py); all three run on the GPU")
    name = {"gatav2": "matpes_gatav2_small.pt", "gatav2_phi": "matpes_gatav2_phi_small.pt",
            "gatav2_global": "matpes_gatav2_global_small.pt"}[variant]      # the last one is BASELINE config 5
    fx = golden(name)
    model = getattr(helpers, "build_" + variant)(fx["hyper"], backend.device)
    load_params(model, fx["params"])
    data = backend.to(dict(fx["inputs"]))
    pos = data["pos"].clone().requires_grad_(True)
    out = model(dict(data, pos=pos))
    assert rel_err(out["energy"], fx["energy"]) < OUT_TOL
    forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
    assert rel_err(forces, fx["forces"]) < OUT_TOL
    wf = torch.linspace(-1, 1, forces.numel(), device=forces.device).view_as(forces)
    we = torch.linspace(0.5, 1.5, out["energy"].numel(), device=forces.device).view_as(out["energy"])
    ((out["energy"] * we).sum() + (forces * wf).sum()).backward()
    _check_grads(model, fx)
    assert any(k not in fx["grads"] for k, _ in model.named_parameters())
