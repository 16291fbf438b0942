"""Diagnostic (GPU): throughput of the GEMM engines on the OC20 shapes, per mode."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import ops, _lib

def bench(M, N, K, tA, tB, engine, mode=0, split=1, reps=5):
    A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    d = ops._desc(A, B, C, None, M, N, K, tA, tB, ops._plain(A.shape[1]), ops._plain(B.shape[1]), ops._plain(N))
    arr = (_lib.GemmDesc * 1)(d)
    def run():
        if engine == "fp32":
            _lib.call("eqv2_gemm_f32", ctypes.cast(arr, ctypes.c_void_p), 1, split, _lib.stream_ptr())
        else:
            _lib.call("eqv2_gemm_tc", ctypes.cast(arr, ctypes.c_void_p), 1, split, mode, _lib.stream_ptr())
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return 2.0 * M * N * K / ms / 1e9, ms

E = 13120
for name, (M, N, K, tA, tB) in {"conv1 fwd m0": (E, 1024, 1792, 0, 1), "conv1 dgrad m0": (E, 1792, 1024, 0, 0),
                                 "conv1 wgrad m0": (1024, 1792, E, 1, 0), "conv2 fwd m1": (E, 1536, 768, 0, 1),
                                 "rad last": (E, 4608, 128, 0, 1), "square 8192": (8192, 8192, 8192, 0, 1)}.items():
    out = [f"{name:16s}"]
    for label, eng, mode in (("3xTF32", "tc", 0), ("3x no-promote", "tc", 2), ("1xTF32", "tc", 1), ("FFMA", "fp32", 0)):
        if eng == "fp32" and M * N * K > 3e11: continue
        tf, ms = bench(M, N, K, tA, tB, eng, mode)
        out.append(f"{label}: {tf:7.1f} TF/s ({ms:.3f} ms)")
    print(" | ".join(out), flush=True)
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
for flag in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = flag
    for _ in range(2): a @ b
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [a @ b for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    print("cuBLAS 8192^3 allow_tf32=", flag, 2 * 8192 ** 3 / (e0.elapsed_time(e1) / 5) / 1e9, "TF/s")
