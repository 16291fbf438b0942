"""GEMM engines vs an fp64 reference: the exact FFMA engine (emulator + GPU) and the tcgen05 tensor-core
engines (GPU only; 3xTF32 and the fp16-split f16x3 engine must stay fp32-class, 1xTF32 within its stated 2e-3)."""
import ctypes

import pytest
import torch

from conftest import PKG
from helpers import pkg


def _run(ops, _lib, backend, M, N, K, transA, transB, engine, bias=False, accumulate=False, split_k=1, groups=1):
    gen = torch.Generator().manual_seed(M * 7 + N * 3 + K + 2 * transA + transB)
    descs, refs, outs = [], [], []
    keep = []
    for g in range(groups):
        A = torch.randn((K, M) if transA else (M, K), generator=gen)
        B = torch.randn((N, K) if transB else (K, N), generator=gen)
        b = torch.randn(N, generator=gen) if bias else None
        C0 = torch.randn(M, N, generator=gen) if accumulate else torch.zeros(M, N)
        ref = (A.double().t() if transA else A.double()) @ (B.double().t() if transB else B.double())
        if bias:
            ref = ref + b.double()
        if accumulate:
            ref = ref + C0.double()
        Ad, Bd, Cd = backend.to(A), backend.to(B), backend.to(C0.clone())
        bd = backend.to(b) if bias else None
        keep += [Ad, Bd, Cd, bd]
        descs.append(ops._desc(Ad, Bd, Cd, bd, M, N, K, transA, transB, ops._plain(A.shape[1]), ops._plain(B.shape[1]),
                               ops._plain(N), accumulate=int(accumulate and split_k == 1)))
        refs.append(ref)
        outs.append(Cd)
    arr = (_lib.GemmDesc * groups)(*descs)
    if engine == "f16x3":
        assert all(ops._f16_addressable(d) for d in descs)
        ops._run_gemm_f16(descs, split_k, 0.0, 0.0)
    elif engine == "fp32":
        _lib.call("eqv2_gemm_f32", ctypes.cast(arr, ctypes.c_void_p), groups, split_k, _lib.stream_ptr())
    else:
        if not all(ops._tc_addressable(d) for d in descs):
            import pytest
            pytest.skip("operand alignment outside the tensor-core engine's contract (served by the FFMA engine)")
        _lib.call("eqv2_gemm_tc", ctypes.cast(arr, ctypes.c_void_p), groups, split_k, 0 if engine == "tf32x3" else 1,
                  _lib.stream_ptr())
    errs = []
    for ref, out in zip(refs, outs):
        errs.append(float((out.double().cpu() - ref).abs().max() / ref.abs().max()))
    return max(errs)


SHAPES = [(128, 128, 32), (256, 128, 64), (200, 136, 96), (77, 40, 36), (384, 256, 1792), (1000, 544, 960)]


@pytest.mark.parametrize("M,N,K", SHAPES[:4])
@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_ffma_engine(backend, M, N, K, tA, tB):
    ops, _lib = pkg("ops"), pkg("_lib")
    assert _run(ops, _lib, backend, M, N, K, tA, tB, "fp32", bias=True) < 2e-6


def test_ffma_engine_split_k_and_groups(backend):
    ops, _lib = pkg("ops"), pkg("_lib")
    assert _run(ops, _lib, backend, 96, 72, 300, 1, 0, "fp32", split_k=3, groups=2) < 2e-6
    assert _run(ops, _lib, backend, 96, 72, 64, 0, 1, "fp32", accumulate=True) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_tensor_core_engine_3xtf32(M, N, K, tA, tB):
    from conftest import Backend
    ops, _lib = pkg("ops"), pkg("_lib")
    err = _run(ops, _lib, Backend("cuda"), M, N, K, tA, tB, "tf32x3", bias=True)
    assert err < 3e-6, err


@pytest.mark.gpu
@pytest.mark.parametrize("tA,tB", [(0, 1), (1, 0)])
def test_tensor_core_engine_1xtf32_and_options(tA, tB):
    from conftest import Backend
    ops, _lib = pkg("ops"), pkg("_lib")
    be = Backend("cuda")
    assert _run(ops, _lib, be, 384, 256, 1792, tA, tB, "tf32") < 2e-3
    assert _run(ops, _lib, be, 300, 200, 4096, tA, tB, "tf32x3", split_k=4, groups=3) < 3e-6
    assert _run(ops, _lib, be, 300, 200, 512, tA, tB, "tf32x3", accumulate=True, bias=True) < 3e-6


F16_SHAPES = SHAPES + [(130, 260, 520), (64, 8, 8), (13120 // 8, 384, 1280), (640, 1, 128), (640, 128, 1), (1, 128, 640)]


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", F16_SHAPES)
@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_f16x3_engine(M, N, K, tA, tB):
    from conftest import Backend
    ops, _lib = pkg("ops"), pkg("_lib")
    err = _run(ops, _lib, Backend("cuda"), M, N, K, tA, tB, "f16x3", bias=True)
    assert err < 3e-6, err


@pytest.mark.gpu
@pytest.mark.parametrize("tA,tB", [(0, 1), (1, 0)])
def test_f16x3_engine_options(tA, tB):
    from conftest import Backend
    ops, _lib = pkg("ops"), pkg("_lib")
    be = Backend("cuda")
    assert _run(ops, _lib, be, 300, 200, 4096, tA, tB, "f16x3", split_k=4, groups=3) < 3e-6
    assert _run(ops, _lib, be, 300, 200, 1000, tA, tB, "f16x3", accumulate=True, groups=2) < 3e-6
    assert _run(ops, _lib, be, 2000, 1100, 9000, tA, tB, "f16x3", bias=True) < 3e-6      # 144 tiles < 148 CTAs, 141 k-blocks
    assert _run(ops, _lib, be, 128, 128, 13120, tA, tB, "f16x3", split_k=32) < 3e-6


@pytest.mark.gpu
def test_f16x3_wide_dynamic_range():
    """Rows 2^-20 .. 2^10 apart in magnitude: the per-tensor power-of-two scale keeps the max-normalised error of every
    output row fp32-class down to 2^-18 of the tensor maximum (documented bound of the split)."""
    from conftest import Backend
    ops, _lib = pkg("ops"), pkg("_lib")
    be = Backend("cuda")
    gen = torch.Generator().manual_seed(5)
    M, N, K = 512, 256, 1024
    A = torch.randn(M, K, generator=gen) * torch.logspace(-5, 3, M).view(-1, 1)
    B = torch.randn(N, K, generator=gen)
    C = torch.zeros(M, N)
    Ad, Bd, Cd = be.to(A), be.to(B), be.to(C)
    d = ops._desc(Ad, Bd, Cd, None, M, N, K, 0, 1, ops._plain(K), ops._plain(K), ops._plain(N))
    ops._run_gemm_f16([d], 1, 0.0, 0.0)
    ref = A.double() @ B.double().t()
    row_err = (Cd.double().cpu() - ref).abs().max(1).values / ref.abs().max(1).values
    big = A.abs().max(1).values >= A.abs().max() * 2.0 ** -18
    assert float(row_err[big].max()) < 3e-6
    assert float(row_err.max()) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("lmax,N,Ci,Co", [(6, 640, 128, 128), (3, 77, 40, 24), (2, 300, 64, 136)])
def test_f16x3_slab_linear_matches_ffma(lmax, N, Ci, Co):
    """SO3_LinearV2 (so3.py:698-743) on the f16x3 engine -- node tensor repacked slab by slab by the split kernel, C
    written back through the two-level row map -- against the exact FFMA engine: output and all three gradients."""
    ops = pkg("ops")
    K = (lmax + 1) ** 2
    gen = torch.Generator().manual_seed(lmax * 100 + N)
    x0 = torch.randn(N, K, Ci, generator=gen).cuda()
    W0 = (torch.randn(lmax + 1, Co, Ci, generator=gen) / Ci ** 0.5).cuda()
    b0 = torch.randn(Co, generator=gen).cuda()
    g = torch.randn(N, K, Co, generator=gen).cuda()
    res = {}
    try:
        for mode in ("fp32", "f16x3"):
            ops.set_gemm_mode(mode)
            x, W, b = x0.clone().requires_grad_(True), W0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
            pkg("_lib").start_kernel_timing()
            y = ops.so3_linear(x, W, b)
            (y * g).sum().backward()
            prof = pkg("_lib").stop_kernel_timing()
            res[mode] = (y.detach(), x.grad, W.grad, b.grad)
            if mode == "f16x3" and N * K * Ci * Co >= (1 << 24):
                assert prof.get("eqv2_gemm_f16", {}).get("calls", 0) == 3, prof.keys()
    finally:
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)
    for a, b_ in zip(res["f16x3"], res["fp32"]):
        assert float((a - b_).abs().max() / b_.abs().max()) < 3e-6


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,col_off,n", [(13489, 4608, 0, 4608), (1000, 2432, 576, 1024), (37, 200, 8, 120), (513, 70, 3, 9)])
def test_planes_colsum_matches_fp32_column_sums(rows, cols, col_off, n):
    """Bias gradients read from operand planes (eqv2_planes_colsum: vectorised 16-byte path and the scalar fallback)."""
    import importlib
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ops = pkg("ops")
    _lib = importlib.import_module(PKG + "._lib")
    ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)
    gen = torch.Generator().manual_seed(rows)
    t = (torch.randn(rows, cols, generator=gen) * 3).cuda()
    sp = ops._splits_for([ops.OperandSrc(t, rows, cols)])[(id(t), 0)]
    S = max(1, min(rows // 16, -(-2368 // ((n + 127) // 128))))
    partial = torch.empty(S * n, dtype=torch.float32, device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.call("eqv2_planes_colsum", sp.buf.data_ptr(), sp.plane, sp.cols_pad, col_off, rows, n, S, sp.absmax.data_ptr(),
              partial.data_ptr(), out.data_ptr(), _lib.stream_ptr(), n_kernels=2)
    ref = t[:, col_off:col_off + n].double().sum(0)
    assert float((out.double() - ref).abs().max()) <= 2e-6 * float(t.abs().max()) * rows ** 0.5 + 1e-4
