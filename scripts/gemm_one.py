"""One GEMM shape / mode for ncu:  python scripts/gemm_one.py M N K tA tB mode"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import ops, _lib
M, N, K, tA, tB, mode = [int(a) for a in sys.argv[1:7]]
A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((N, K) if tB else (K, N), device="cuda")
C = torch.zeros(M, N, device="cuda")
d = ops._desc(A, B, C, None, M, N, K, tA, tB, ops._plain(A.shape[1]), ops._plain(B.shape[1]), ops._plain(N))
arr = (_lib.GemmDesc * 1)(d)
for _ in range(3):
    _lib.call("eqv2_gemm_tc", ctypes.cast(arr, ctypes.c_void_p), 1, 1, mode, _lib.stream_ptr())
torch.cuda.synchronize()
print("ok", float(C.abs().sum()))
