"""Diagnostic (GPU): which ATen operators (torch glue around the C-ABI kernels) one eager OC20 train step still launches,
aggregated by operator name and input shapes, with the Python frame that called them.
    python scripts/profile_aten.py [--config oc20] [--top 40]"""
import argparse, collections, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="oc20", choices=sorted(bench.CONFIGS))
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
CFG = bench.CONFIGS[a.config]
torch.manual_seed(0)
dev = torch.device("cuda")
model = getattr(importlib.import_module(bench.PKG + ".models." + CFG["module"]), CFG["cls"])(**CFG["kw"]).to(dev)
opt = importlib.import_module(bench.PKG + ".optim").FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-3,
                                                                max_grad_norm=100.0, ema_decay=0.999)
data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.make_batch(CFG, CFG["structures"], seed=1000).items()}


def step():
    loss = bench.forward_loss(CFG, model, data)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, None])
PKGDIR = bench.PKG
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
        continue
    dt = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    if dt <= 0 or ev.cpu_children and any(c.name.startswith("aten::") and (getattr(c, "device_time_total", 0) or 0) > 0 for c in ev.cpu_children):
        continue            # count leaf operators only
    frame = next((f for f in (ev.stack or []) if PKGDIR in f or "bench.py" in f), "(autograd engine)")
    key = (ev.name, str(ev.input_shapes)[:70], frame.split("/")[-1][:70])
    r = agg[key]
    r[0] += 1
    r[1] += dt
tot = sum(v[1] for v in agg.values())
print(f"ATen leaf operators with device time: {sum(v[0] for v in agg.values())} calls, {tot / 1e3:.2f} ms")
for k, (c, t, _) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"{t / 1e3:7.3f} ms {c:4d} x {k[0]:28s} {k[1]:70s} {k[2]}")
