"""GPU only: a mid-size OC20-shaped model (large enough that every dense contraction runs on the tcgen05
engine, including the two-level strided SO3_LinearV2 problems) must give the same energies / forces /
parameter gradients in the fp32-class tensor-core modes (3xTF32 and the fp16-split f16x3 default) as with the exact FFMA engine -- the 1e-5 bound of the fp32 mode."""
import pytest
import torch

from helpers import pkg, rel_err


def _run(mode, data, seed_frames):
    ops = pkg("ops")
    oc20 = pkg("models.equiformerv2_oc20")
    ops.set_gemm_mode(mode)
    try:
        torch.manual_seed(0)
        model = oc20.EquiformerV2_OC20(num_layers=2, sphere_channels=64, attn_hidden_channels=32, num_heads=4,
                                       attn_alpha_channels=32, attn_value_channels=16, ffn_hidden_channels=64,
                                       lmax_list=[4], mmax_list=[2], edge_channels=64, alpha_drop=0.0, drop_path_rate=0.0,
                                       max_radius=8.0).cuda()
        gen = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.02 * torch.randn(p.shape, generator=gen).cuda())
        torch.manual_seed(seed_frames)
        energy, forces = model(data)
        w = torch.linspace(-1, 1, forces.numel(), device="cuda").view_as(forces)
        (energy.sum() + (forces * w).sum()).backward()
        return energy.detach(), forces.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}
    finally:
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)


@pytest.mark.gpu
def test_tensor_core_mode_matches_ffma_mode():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _lib = pkg("_lib")
    _lib._state["lib"] = None
    syn = pkg("synthetic")
    data = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.oc20_batch(4, seed=5).items()}
    e0, f0, g0 = _run("fp32", data, 7)
    _lib.start_kernel_timing()
    e1, f1, g1 = _run("tf32x3", data, 7)
    prof = _lib.stop_kernel_timing()
    assert prof.get("eqv2_gemm_tc", {}).get("calls", 0) >= 20, "the tensor-core engine did not run"
    assert rel_err(e1, e0) < 1e-5 and rel_err(f1, f0) < 1e-5
    bad = [(k, rel_err(g1[k], g0[k])) for k in g0 if rel_err(g1[k], g0[k]) > 1e-4]
    assert not bad, bad
    e2, f2, _ = _run("tf32", data, 7)
    assert rel_err(e2, e0) < 5e-3 and rel_err(f2, f0) < 5e-3
    _lib.start_kernel_timing()
    e3, f3, g3 = _run("f16x3", data, 7)
    prof = _lib.stop_kernel_timing()
    assert prof.get("eqv2_gemm_f16", {}).get("calls", 0) >= 20, "the f16x3 engine did not run"
    assert rel_err(e3, e0) < 1e-5 and rel_err(f3, f0) < 1e-5
    bad = [(k, rel_err(g3[k], g0[k])) for k in g0 if rel_err(g3[k], g0[k]) > 1e-4]
    assert not bad, bad


@pytest.mark.gpu
def test_s2_written_operand_planes_match_the_fp32_route():
    """The S2 activation writing conv2's A operand as fp16 hi/lo planes (eqv2_s2sep_fwd_planes, bound from max |Y| and the
    activation's operator norm) against the same block with Z as an fp32 tensor + operand split: outputs and every parameter
    gradient of one attention block + force head agree to fp32 rounding."""
    import importlib
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from conftest import PKG
    ops = pkg("ops")
    _lib = importlib.import_module(PKG + "._lib")
    _lib._state["lib"] = None
    ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)
    oc20, syn = pkg("models.equiformerv2_oc20"), pkg("synthetic")
    torch.manual_seed(0)
    model = oc20.EquiformerV2_OC20(num_layers=1, sphere_channels=128, attn_hidden_channels=64, num_heads=8,
                                   attn_alpha_channels=64, attn_value_channels=16, ffn_hidden_channels=128, lmax_list=[6],
                                   mmax_list=[2], alpha_drop=0.0, drop_path_rate=0.0).cuda()
    data = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.oc20_batch(2, seed=11).items()}

    def run(flag):
        ops._FEATURES["s2_planes"] = flag
        ops.reset_caches()
        torch.manual_seed(5)
        model.zero_grad(set_to_none=True)
        _lib.start_kernel_timing()
        energy, forces = model(data)
        (energy.sum() + (forces * torch.linspace(-1, 1, forces.numel(), device="cuda").view_as(forces)).sum()).backward()
        prof = _lib.stop_kernel_timing()
        return energy.detach().clone(), forces.detach().clone(), [p.grad.clone() for p in model.parameters()], prof

    try:
        e1, f1, g1, prof1 = run(True)
        e0, f0, g0, prof0 = run(False)
    finally:
        ops._FEATURES["s2_planes"] = True
    assert prof1.get("eqv2_s2sep_fwd_planes", {}).get("calls", 0) >= 2 and "eqv2_s2sep_fwd_planes" not in prof0
    rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert rel(e1, e0) < 2e-6 and rel(f1, f0) < 5e-6, (rel(e1, e0), rel(f1, f0))
    worst = max(rel(a, b) for a, b in zip(g1, g0))
    assert worst < 2e-5, worst
