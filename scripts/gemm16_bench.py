"""Diagnostic (GPU): the f16x3 engine on the OC20 shapes -- accuracy vs fp64, operand-split cost, GEMM-only throughput
(splits prepared), next to the tf32x3 engine.   python scripts/gemm16_bench.py [quick]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import ops, _lib


def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def one(name, M, N, K, tA, tB, split=1, check=True):
    A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    d = ops._desc(A, B, C, None, M, N, K, tA, tB, ops._plain(A.shape[1]), ops._plain(B.shape[1]), ops._plain(N))
    out = [f"{name:16s}"]
    ops._run_gemm_f16([d], 1, 0, 0)
    torch.cuda.synchronize()
    if check:
        rows = slice(0, min(M, 512))
        ref = ((A.double().t() if tA else A.double())[rows] @ (B.double().t() if tB else B.double()))
        out.append(f"err {float((C[rows].double() - ref).abs().max() / ref.abs().max()):.1e}")
    t_split = timeit(lambda: ops._splits_for(list(d.src)))
    with ops.split_scope([]):
        ops._splits_for(list(d.src))              # registers both splits in the scope
        def run16():
            if split > 1: C.zero_()
            ops._run_gemm_f16([d], split, 0, 0)
        ms = timeit(run16)
    out.append(f"split {t_split:.3f} ms | f16x3 {2.0 * M * N * K / ms / 1e9:7.1f} TF/s ({ms:.3f} ms)")
    arr = (_lib.GemmDesc * 1)(d)
    ms = timeit(lambda: _lib.call("eqv2_gemm_tc", ctypes.cast(arr, ctypes.c_void_p), 1, split, 0, _lib.stream_ptr()))
    out.append(f"tf32x3 {2.0 * M * N * K / ms / 1e9:7.1f} TF/s ({ms:.3f} ms)")
    print(" | ".join(out), flush=True)


E = 13120
shapes = {"conv1 fwd m0": (E, 1024, 1792, 0, 1), "conv1 dgrad m0": (E, 1792, 1024, 0, 0),
          "conv1 wgrad m0": (1024, 1792, E, 1, 0), "conv2 fwd m1": (E, 1536, 768, 0, 1),
          "rad last": (E, 4608, 128, 0, 1), "square 8192": (8192, 8192, 8192, 0, 1)}
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = {"small": (256, 256, 512, 0, 1), "conv1 fwd m0": shapes["conv1 fwd m0"]}
for name, (M, N, K, tA, tB) in shapes.items():
    one(name, M, N, K, tA, tB)
