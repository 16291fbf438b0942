import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equivarianttransformermpnn4quantumcomputations_b200 import synthetic, graphs, ops
from equivarianttransformermpnn4quantumcomputations_b200.models import equiformerv2_oc20 as oc20

def _loss(out, d):
    return (out[0] - d["energy"]).abs().mean() + (out[1] - d["forces"]).abs().mean()

def make():
    torch.manual_seed(0)
    m = oc20.EquiformerV2_OC20(num_layers=2, sphere_channels=32, attn_hidden_channels=16, num_heads=2,
                               attn_alpha_channels=16, attn_value_channels=8, ffn_hidden_channels=32, lmax_list=[3],
                               mmax_list=[2], edge_channels=32, alpha_drop=0.0, drop_path_rate=0.0, max_radius=8.0).cuda()
    return m, torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True)

mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
ops.set_gemm_mode(mode)
a = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synthetic.oc20_batch(2, seed=5).items()}
m0, o0 = make(); m1, o1 = make()
st = graphs.GraphedTrainStep(m1, _loss, o1)
for it in range(3):
    torch.manual_seed(10 + it)
    l0 = _loss(m0(a), a); o0.zero_grad(set_to_none=True); l0.backward()
    torch.manual_seed(10 + it)
    # graphed, but look at the grads before the optimizer runs
    full = st._with_prepared(a)
    sig = (int(full["pos"].shape[0]), int(full["edge_index"].shape[1]), len(full["natoms"]))
    if sig not in st.graphs: st._capture(full, sig)
    g, static, loss, grads, _ = st.graphs[sig]
    for k, v in full.items():
        if torch.is_tensor(v): static[k].copy_(v)
    g.replay()
    torch.cuda.synchronize()
    bad = []
    for (n, p0), p1 in zip(m0.named_parameters(), m1.parameters()):
        if p0.grad is None or p1.grad is None:
            bad.append((n, "none", p0.grad is None, p1.grad is None)); continue
        e = float((p0.grad - p1.grad).abs().max() / (p0.grad.abs().max() + 1e-20))
        if e > 1e-4: bad.append((n, e))
    print("iter", it, "loss", float(l0), float(loss), "param diff", max(float((x - y).abs().max()) for x, y in zip(m0.parameters(), m1.parameters())), "bad grads", len(bad), bad[:6])
    o0.step(); o1.step()
