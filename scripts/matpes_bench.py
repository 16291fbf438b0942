"""BASELINE configs 3-5 (MatPES family; `--model base | gatav2 | gatav2_phi | global`, default base): train step = forward energy -> forces by
autograd.grad(create_graph=True) -> L1(E) + L1(F) -> loss.backward() (double backward) -> AdamW,
on synthetic 30-atom bulk cells (6 A cutoff, max 20 neighbours), batch 8 per GPU (config_cosinelearning.py:76).
Prints one JSON line: GPU structures/s, per-entry-point time shares, and the CPU oracle port on a bounded sample."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
PKG = "equivarianttransformermpnn4quantumcomputations_b200"
_lib = importlib.import_module(PKG + "._lib")
syn = importlib.import_module(PKG + ".synthetic")
MODELS = {"base": "equiformerv2_MatPESv2", "gatav2": "equiformerv2_MatPES_GATAV2",
          "gatav2_phi": "equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata",
          "global": "equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE"}
WHICH = next((a.split("=")[1] for a in sys.argv if a.startswith("--model=")), "base")
mp = importlib.import_module(PKG + ".models." + MODELS[WHICH])
ATOMS = int(next((a.split("=")[1] for a in sys.argv if a.startswith("--atoms=")), 200 if WHICH == "global" else 30))

KW = dict(max_neighbors=20, max_radius=6.0, num_layers=6, sphere_channels=128, attn_hidden_channels=128, num_heads=8,
          attn_alpha_channels=32, attn_value_channels=16, ffn_hidden_channels=512, lmax_list=[4], mmax_list=[2],
          edge_channels=128, alpha_drop=0.0, drop_path_rate=0.0)
if WHICH != "base":      # configs/MatPES/config_cosinelearningGATA.py / ...MoreGATA_all2all.py: mmax 4
    KW["mmax_list"] = [4]


def train_step(model, opt, data, w_e=1.0, w_f=1.0):
    pos = data["pos"].detach().requires_grad_(True)
    out = model(dict(data, pos=pos))
    forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
    loss = w_e * (out["energy"] - data["energy"]).abs().mean() + w_f * (forces - data["forces"]).abs().mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def forward_loss(model, data, w_e=1.0, w_f=1.0):
    pos = data["pos"].detach().requires_grad_(True)
    out = model(dict(data, pos=pos))
    forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
    return w_e * (out["energy"] - data["energy"]).abs().mean() + w_f * (forces - data["forces"]).abs().mean()


def main(B=8, steps=5, warmup=3):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = mp.EquiformerV2_MatPES(**KW).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    host = syn.matpes_batch(B, seed=7, n_atoms=ATOMS)
    data = {k: v.to(dev) for k, v in host.items()}
    graphed = "--no-graph" not in sys.argv
    if graphed:     # forward + force gradient + loss + double backward replayed from a CUDA graph (graphs.py)
        graphs = importlib.import_module(PKG + ".graphs")
        stepper = graphs.GraphedTrainStep(model, None, opt, forward_loss=lambda d: forward_loss(model, d))
        run = lambda: stepper(data)
    else:
        run = lambda: train_step(model, opt, data)
    for _ in range(warmup):
        loss = run()
    torch.cuda.synchronize()
    E = int(model.generate_graph(data["pos"], data["batch"], data["cell"], data["natoms"])[0].shape[1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if graphed:     # release the graph's private memory pool before the eager profiling pass (config 5: ~100 GB)
        import gc
        opt.zero_grad(set_to_none=True)
        loss = loss.detach().clone()
        del stepper, run
        gc.collect()
        torch.cuda.empty_cache()
    _lib.start_kernel_timing()
    train_step(model, opt, data)
    prof = _lib.stop_kernel_timing()
    tot = sum(r["ms"] for r in prof.values())
    shares = {k: round(v["ms"] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    if WHICH != "base":
        print(json.dumps({"workload": f"MatPES {WHICH} ({MODELS[WHICH]}; lmax 4, mmax 4, 6 blocks) train step with autograd "
                                      "forces (double backward)", "structures_per_s": B / (ms / 1e3), "ms_per_step": ms,
                          "launch": "CUDA graph replay" if graphed else "eager", "structures": B,
                          "atoms": int(data["pos"].shape[0]), "edges": E, "loss": float(loss), "params": model.num_params,
                          "kernel_time_shares": shares, "kernel_ms_total": tot}))
        return
    # CPU oracle port: 2 cells, same pattern
    from oracle import eqv2_oracle as O
    torch.set_num_threads(os.cpu_count())
    cm = mp.EquiformerV2_MatPES(**KW)
    P = dict(cm.named_parameters())
    hp = O.Hyper(lmax=4, mmax=2, C=128, H=128, heads=8, alpha_ch=32, value_ch=16, ffn_hidden=512, edge_ch=128, num_layers=6,
                 norm_type="rms_norm_sh", cutoff=6.0, max_neighbors=20, max_elements=100)
    hp.avg_degree = 12.0
    small = syn.matpes_batch(2, seed=9)
    t0 = time.perf_counter()
    ei, _, _, _ = O.radius_graph_matpes(small["pos"], small["cell"], small["batch"], 6.0, 20, 2)
    pos = small["pos"].clone().requires_grad_(True)
    et = O.matpes_v2_forward(P, hp, small["atomic_numbers"], small["batch"], small["natoms"], pos, ei)
    f = -torch.autograd.grad(et.sum(), pos, create_graph=True)[0]
    ((et / small["natoms"]).unsqueeze(1) - small["energy"]).abs().mean().add((f - small["forces"]).abs().mean()).backward()
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "MatPES EquiformerV2 (lmax 4, mmax 2, 6 blocks) train step with autograd forces (double backward)",
                      "structures_per_s": B / (ms / 1e3), "ms_per_step": ms, "launch": "CUDA graph replay" if graphed else "eager", "structures": B, "atoms": int(data["pos"].shape[0]),
                      "edges": E, "loss": float(loss), "params": model.num_params, "kernel_time_shares": shares,
                      "kernel_ms_total": tot,
                      "cpu_baseline": {"structures_per_s": 2 / dt, "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"2 cells ({ei.shape[1]} edges) fwd + force grad + double bwd, {dt:.1f} s"}}))


if __name__ == "__main__":
    main()
