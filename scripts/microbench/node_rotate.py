"""Microbenchmark (GPU): the node-centric rotate kernels at the OC20 shape (640 atoms, ~13.5 k edges, C = 128, lmax 6 /
mmax 2) under the tunables of csrc/rotate.cu (EQV2_NODE_PIPE / EQV2_NODE_CW / EQV2_NODE_STAGES).  Inputs exceed L2
(dA 400 MB + rad 249 MB), so consecutive launches see cold data."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
N, deg, C, lmax, mmax = 640, 21, 128, 6, 2
E = N * deg
dst = torch.arange(N, device=dev).repeat_interleave(deg)
src = (dst // 80) * 80 + torch.randint(0, 80, (E,), device=dev)
plan = ops.EdgePlan(torch.stack([src, dst]), N)
lay = ops.CoeffLayout.get(lmax, mmax)
WS = sum((2 * l + 1) ** 2 for l in range(lmax + 1))
wig = torch.randn(E, WS, device=dev)
rad = torch.randn(E, lay.nslot * 2 * C, device=dev)
gA = torch.randn(E, lay.Kr * 2 * C, device=dev)
x = torch.randn(N, lay.K, C, device=dev)
val = torch.randn(E, lay.Kr * C, device=dev)
alpha = torch.rand(E, 8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for pipe, cw, st in [("0", "", ""), ("1", "128", "2"), ("1", "64", "2"), ("1", "32", "2"), ("1", "128", "3"), ("1", "64", "3")]:
    os.environ["EQV2_NODE_PIPE"] = pipe
    for k, v in (("EQV2_NODE_CW", cw), ("EQV2_NODE_STAGES", st)):
        if v:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    t_dx = timeit(lambda: ops._gr_bwd(x, rad, gA, plan, wig, lmax, mmax, want_dx=True, want_drad=False))
    t_rf = timeit(lambda: ops._rir_fwd(val, alpha, plan, wig, lmax, mmax, lay.Kr, 8, 1.0, C))
    print(f"pipe={pipe} CW={cw or '-':>3} stages={st or '-'}   gather_rotate_dx {t_dx:7.1f} us   rotinv_reduce_fwd {t_rf:7.1f} us", flush=True)
