"""Diagnostic (GPU): per-kernel device time of ONE eager OC20 train step from torch.profiler (CUPTI), aggregated by kernel
name -- the cheap companion of the ncu launch list (profiles/).   python scripts/profile_step.py [--layers N] [--top K]"""
import argparse, collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from equivarianttransformermpnn4quantumcomputations_b200 import synthetic
from equivarianttransformermpnn4quantumcomputations_b200.models import equiformerv2_oc20 as oc20

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--structures", type=int, default=None)
ap.add_argument("--config", default="oc20", choices=sorted(bench.CONFIGS))
a = ap.parse_args()
CFG = bench.CONFIGS[a.config]
kw = dict(CFG["kw"])
if a.config == "oc20":
    kw["num_layers"] = a.layers
torch.manual_seed(0)
dev = torch.device("cuda")
import importlib
model = getattr(importlib.import_module(bench.PKG + ".models." + CFG["module"]), CFG["cls"])(**kw).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
data = {k: (v.to(dev) if torch.is_tensor(v) else v)
        for k, v in bench.make_batch(CFG, a.structures or CFG["structures"], seed=1000).items()}


def step():
    loss = bench.forward_loss(CFG, model, data)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from equivarianttransformermpnn4quantumcomputations_b200 import ops
ops._GEMM_LOG = []
step()
torch.cuda.synchronize()
seen = collections.Counter(str(x) for x in ops._GEMM_LOG)
ops._GEMM_LOG = None
print("GEMM launches declined by the f16x3 engine (M, N, K, tA, tB, addressable):")
for k, c in seen.most_common(20):
    print(f"  {c:4d} x {k}")
ops._GEMM16_LOG = []
ops._SPLIT_LOG = []
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
glog, ops._GEMM16_LOG = ops._GEMM16_LOG, None
slog, ops._SPLIT_LOG = ops._SPLIT_LOG, None
sagg = collections.Counter()
for r, c, given in slog:
    sagg[(r, c, given)] += 1
print("operand splits (rows, cols, maximum known) x count, MB read+written per step:")
for (r, c, given), n in sorted(sagg.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[1])[:16]:
    print(f"  {n:4d} x ({r}, {c}, {given})  {n * r * c * (8 if given else 12) / 1e6:8.1f} MB")
gev = sorted([ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and "gemm_f16_kernel" in ev.name],
             key=lambda ev: ev.time_range.start)
if len(gev) == len(glog):
    per = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for ev, (shapes, sk) in zip(gev, glog):
        key = (shapes if len(shapes) <= 2 else (shapes[0], "... %d groups" % len(shapes)), sk)
        r = per[key]
        r[0] += 1; r[1] += ev.device_time; r[2] += sum(2.0 * m * n * k for m, n, k, _, _ in shapes)
    print("f16x3 GEMM launches by shape (M, N, K, tA, tB) x split_k:")
    for key, (c, t, fl) in sorted(per.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"  {t / 1e3:7.3f} ms {c:4d} x {fl / t / 1e6:6.1f} TFLOP/s  {key}")
else:
    print("gemm events / log mismatch", len(gev), len(glog))
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("at::native::", "")
        agg[n[:100]][0] += 1
        agg[n[:100]][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"total device time {tot / 1e3:.2f} ms in {sum(v[0] for v in agg.values())} kernels/memops")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}% {c:5d}  {k}")
