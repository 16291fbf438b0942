"""One GEMM shape on the f16x3 engine for ncu:  python scripts/gemm16_one.py M N K tA tB"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import ops
M, N, K, tA, tB = [int(a) for a in sys.argv[1:6]]
A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((N, K) if tB else (K, N), device="cuda")
C = torch.zeros(M, N, device="cuda")
d = ops._desc(A, B, C, None, M, N, K, tA, tB, ops._plain(A.shape[1]), ops._plain(B.shape[1]), ops._plain(N))
with ops.split_scope([]):
    for _ in range(3):
        ops._run_gemm_f16([d], 1, 0, 0)
torch.cuda.synchronize()
print("ok", float(C.abs().sum()))
