"""Diagnostic (GPU): accumulation accuracy of the tcgen05 kind::tf32 engine vs K, against fp64."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import ops, _lib

def run(A, B, engine):
    M, K = A.shape; N = B.shape[0]
    C = torch.zeros(M, N, device="cuda")
    d = ops._desc(A, B, C, None, M, N, K, 0, 1, ops._plain(K), ops._plain(K), ops._plain(N))
    arr = (_lib.GemmDesc * 1)(d)
    if engine == "fp32":
        _lib.call("eqv2_gemm_f32", ctypes.cast(arr, ctypes.c_void_p), 1, 1, _lib.stream_ptr())
    else:
        _lib.call("eqv2_gemm_tc", ctypes.cast(arr, ctypes.c_void_p), 1, 1, 0 if engine == "tf32x3" else 1, _lib.stream_ptr())
    return C

def tf32_round(x):
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)

torch.manual_seed(0)
for K in (32, 128, 512, 2048, 8192):
    A = torch.randn(256, K, device="cuda"); B = torch.randn(256, K, device="cuda")
    At, Bt = tf32_round(A), tf32_round(B)
    ref = A.double() @ B.double().t(); reft = At.double() @ Bt.double().t()
    def e(C, r): return float((C.double() - r).abs().max() / r.abs().max()), float((C.double() - r).mean() / r.abs().mean())
    print(K, "ffma", e(run(A, B, "fp32"), ref), "3x", e(run(A, B, "tf32x3"), ref), "1x(exact inputs)", e(run(At, Bt, "tf32"), reft),
          "torch fp32", e(A @ B.t(), ref))
# positive inputs expose a systematic (truncation) bias
for K in (512, 2048):
    A = torch.rand(256, K, device="cuda"); B = torch.rand(256, K, device="cuda")
    At, Bt = tf32_round(A), tf32_round(B)
    reft = At.double() @ Bt.double().t()
    C = run(At, Bt, "tf32")
    print("positive", K, float(((C.double() - reft) / reft).mean()), float(((C.double() - reft) / reft).abs().max()))
