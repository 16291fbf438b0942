#!/usr/bin/env python
"""Benchmark of the hot path on the five BASELINE.json workloads.

  python bench.py --gpus 1 --steps K --warmup W                 # headline: OC20 S2EF train step (BASELINE configs[1])
  python bench.py --config {oc20,qm9,matpes,gatav2,global} ...  # the other BASELINE configs, same contract
  python bench.py --impl reference ...                          # the UNMODIFIED reference on the host CPU cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...         # data parallel, one rank per GPU (NCCL)

A "step" = neighbour list + forward + L1 losses + backward (+ double backward for the MatPES family, whose forces are
-dE/dpos by autograd) + AdamW update on a batch of synthetic structures of the config's shape, fp32,
random-init weights of the reference architecture.  Prints ONE JSON line (rank 0):
  value      : structures/s with inputs resident in HBM
  e2e        : the same through the public call with pinned HOST buffers copied in every step and the loss read back
  roofline   : the dominant kernel against the measured peak; rooflines: the eight most expensive entry points, each with
               algorithmic FLOPs / bytes computed in code by the op wrappers (ops.py `work=`), CUDA-event times
  cpu_baseline : the reference's own CPU implementation (oracle/_ref = byte-identical copy of the reference's Python
               files, else the oracle port) on a bounded sample of the same workload, warm-up + median of 3
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
PKG = "equivarianttransformermpnn4quantumcomputations_b200"

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the committed `ncu --set full`
# capture of the round-2 build (profiles/r02_ncu_kernel_summary.txt: conv1 forward, all five m-groups in one launch,
# 157 GFLOP; 425 MB read + 115 MB written against 554 MB algorithmic).  A constant of that capture, not of this run (ncu
# cannot run inside the timed bench): reported with its source.
TRAFFIC = {"eqv2_gemm_f16": (5.397e8, "profiles/r02_ncu_kernel_summary.txt (conv1 forward launch, E = 13 489; algorithmic 554 MB)")}
FFMA_PEAK_TFLOPS = 73.0      # measured FFMA / FFMA2 peak of this pool's B200 (scripts/microbench/ffma2.cu)
UNIT = "structures/s"

_MATPES_KW = dict(max_neighbors=20, max_radius=6.0, max_num_elements=100, num_layers=6, sphere_channels=128,
                  attn_hidden_channels=128, num_heads=8, attn_alpha_channels=32, attn_value_channels=16,
                  ffn_hidden_channels=512, lmax_list=[4], mmax_list=[2], grid_resolution=18, edge_channels=128,
                  alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)

CONFIGS = {
    # BASELINE configs[1] -- the configuration the metric is quoted on (equiformerv2_oc20.py ctor defaults)
    "oc20": dict(
        metric="oc20_s2ef_train_structures_per_s", module="equiformerv2_oc20", cls="EquiformerV2_OC20", kind="direct",
        kw=dict(max_neighbors=20, max_radius=12.0, max_num_elements=90, num_layers=12, sphere_channels=128,
                attn_hidden_channels=64, num_heads=8, attn_alpha_channels=64, attn_value_channels=16,
                ffn_hidden_channels=128, norm_type="rms_norm_sh", lmax_list=[6], mmax_list=[2], grid_resolution=18,
                edge_channels=128, alpha_drop=0.1, drop_path_rate=0.05, proj_drop=0.0),
        structures=8, ref_structures=4, blocks_extra=1,
        workload="OC20 S2EF EquiformerV2 (lmax 6, mmax 2, 12 blocks + force head) train step (graph build + fwd + L1 "
                 "loss + bwd + AdamW), ~80-atom slabs, 12 A cutoff, max 20 neighbours, attention dropout 0.10 / "
                 "stochastic depth 0.05"),
    # BASELINE configs[0] (configs/QM9/config_equiformerV2_mu_alpha_homo_lumo_osv.py)
    "qm9": dict(
        metric="qm9_train_structures_per_s", module="equiformerv2_qm9", cls="EquiformerV2_QM9", kind="qm9",
        kw=dict(num_targets=6, max_neighbors=500, max_radius=5.0, max_num_elements=10, num_layers=6, sphere_channels=96,
                attn_hidden_channels=48, num_heads=4, attn_alpha_channels=64, attn_value_channels=24,
                ffn_hidden_channels=96, lmax_list=[4], mmax_list=[4], grid_resolution=18, edge_channels=64,
                alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0),
        structures=128, ref_structures=16, blocks_extra=0,
        workload="QM9 EquiformerV2 (lmax 4, mmax 4, 96 channels, 6 blocks, 6 property heads) train step, 128 molecules "
                 "of 9-29 atoms, 5 A cutoff"),
    # BASELINE configs[2] (configs/MatPES/config_cosinelearning.py; forces by autograd -> double backward)
    "matpes": dict(
        metric="matpes_train_structures_per_s", module="equiformerv2_MatPESv2", cls="EquiformerV2_MatPES", kind="matpes",
        kw=_MATPES_KW, structures=8, ref_structures=2, blocks_extra=0, atoms=30,
        workload="MatPES EquiformerV2 (lmax 4, mmax 2, 6 blocks) train step with autograd forces (fwd + force gradient + "
                 "double backward + AdamW), 30-atom bulk cells, 6 A cutoff, max 20 neighbours"),
    # BASELINE configs[3] (configs/MatPES/config_cosinelearningMoreGATA.py: batch 16, lmax = mmax = 4)
    "gatav2": dict(
        metric="matpes_gatav2_train_structures_per_s",
        module="equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata", cls="EquiformerV2_MatPES", kind="matpes",
        kw=dict(_MATPES_KW, mmax_list=[4]), structures=16, ref_structures=2, blocks_extra=0, atoms=30,
        workload="MatPES GATAV2 phi-at-every-iteration (HTR edge stream + GATA value activation; lmax 4, mmax 4, 6 blocks) "
                 "train step with autograd forces, 30-atom bulk cells, 6 A cutoff, max 20 neighbours"),
    # BASELINE configs[4] (200-atom cells, dense all-pairs attention)
    "global": dict(
        metric="matpes_global_attention_train_structures_per_s",
        module="equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE",
        cls="EquiformerV2_MatPES", kind="matpes", kw=dict(_MATPES_KW, mmax_list=[4]), structures=4, ref_structures=1,
        blocks_extra=0, atoms=200,
        workload="MatPES GATAV2 + global all-to-all node attention (lmax 4, mmax 4, 6 blocks) train step with autograd "
                 "forces, 200-atom cells (local graph + dense all-pairs attention per cell)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="oc20", choices=sorted(CONFIGS))
    ap.add_argument("--structures", type=int, default=None, help="structures per GPU per step (default: the config's)")
    ap.add_argument("--ref-structures", type=int, default=None,
                    help="structures per step of the CPU reference arm (bounded sample of the workload)")
    ap.add_argument("--layers", type=int, default=None, help="(debug) override the number of blocks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-dropout", action="store_true", help="alpha_drop = drop_path_rate = 0")
    ap.add_argument("--bucket", default=None, help="pad every batch to multiples ATOMS,EDGES (e.g. 64,512) with a masked "
                    "ghost structure so that one captured graph serves a whole bucket of batch sizes")
    ap.add_argument("--grad-sync", default="overlap", choices=["overlap", "captured", "flat"],
                    help="N > 1: 'overlap' = bucket all-reduces issued from inside the (captured) backward pass on a side "
                         "stream, gradients live in the flat buckets; 'flat' = copy / all-reduce / copy back after it")
    ap.add_argument("--sync-bucket-mb", type=float, default=32.0, help="N > 1: size of the gradient all-reduce buckets")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: clip_grad_norm_ + AdamW + EMA as one multi-tensor pass (optim.FusedAdamW, the reference's "
                         "optimizer-side step); torch: torch.optim.AdamW(fused=True) alone (the round-1 step)")
    ap.add_argument("--disable", default="", help="(A/B measurements) comma list: planes (producer-written operand "
                    "planes), c_absmax (max |C| in the GEMM epilogue)")
    ap.add_argument("--gemm-mode", default=None, choices=["f16x3", "f16", "tf32x3", "tf32", "fp32"],
                    help="GEMM engine (default: the package default, f16x3 = fp32-class accuracy)")
    return ap.parse_args()


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "200", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].strip() == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(rows[0][1]) if rows else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload pieces shared by both arms
def model_kw(cfg, args):
    kw = dict(cfg["kw"])
    if args.layers:
        kw["num_layers"] = args.layers
    if args.no_dropout:
        kw.update(alpha_drop=0.0, drop_path_rate=0.0)
    return kw


def make_batch(cfg, n, seed):
    import importlib
    syn = importlib.import_module(PKG + ".synthetic")
    if cfg["kind"] == "direct":
        return syn.oc20_batch(n, seed=seed)
    if cfg["kind"] == "qm9":
        d = syn.qm9_batch(n, seed=seed)
        d["targets"] = torch.randn(n, cfg["kw"]["num_targets"], generator=torch.Generator().manual_seed(seed + 1))
        return d
    return syn.matpes_batch(n, seed=seed, n_atoms=cfg["atoms"])


def forward_loss(cfg, model, data):
    """The scalar a train step back-propagates (reference train scripts: L1 on energy and forces,
    train_oc20v2_parallel.py:150-176, train_MatPES_GATAWandB.py:67-91; L1 on the property vector for QM9)."""
    if "atom_mask" in data:          # batch padded to an (atoms, edges) bucket: the ghost structure is masked out
        import importlib
        mm = importlib.import_module(PKG + ".batching").masked_mean
        mean_s = lambda x: mm(x, data["structure_mask"])
        mean_a = lambda x: mm(x, data["atom_mask"])
    else:
        mean_s = mean_a = lambda x: x.mean()
    if cfg["kind"] == "direct":
        energy, forces = model(data)
        return mean_s((energy - data["energy"]).abs()) + mean_a((forces - data["forces"]).abs())
    if cfg["kind"] == "qm9":
        return mean_s((model(data) - data["targets"]).abs())
    pos = data["pos"].detach().requires_grad_(True)
    out = model(dict(data, pos=pos))
    forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
    return mean_s((out["energy"] - data["energy"]).abs()) + mean_a((forces - data["forces"]).abs())


def config_dict(cfg, kw, structures, world):
    """Identical for both arms (the reference arm describes its bounded sample in cpu_baseline.sample)."""
    return {"workload": cfg["workload"], "structures_per_gpu": structures, "layers": kw["num_layers"],
            "parallelism": f"dp{world}",
            "l2": "step working set (weights + >1 GB of activations) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
def reference_stepper(cfg, kw):
    """-> (step(data) -> loss, kind, note): the unmodified reference model on the host CPU when its byte-identical copy is
    present (oracle/_ref, or /root/reference in the build container), else the oracle port (OC20 / MatPES base only)."""
    import importlib
    from oracle import ref_loader
    if os.path.isdir(ref_loader.REF_ROOT):
        ref_loader.install()
        mod = importlib.import_module(cfg["module"])
        torch.manual_seed(0)
        model = getattr(mod, cfg["cls"])(**kw)
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-3)
        shadow = {n: p.data.clone() for n, p in model.named_parameters() if p.requires_grad}

        def step(data):     # the reference's optimizer-side step (train_oc20v2_parallel.py:171-186): clip, AdamW, EMA
            loss = forward_loss(cfg, model, data)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 100.0)
            opt.step()
            for n, p in model.named_parameters():
                if p.requires_grad:
                    shadow[n] = ((1.0 - 0.999) * p.data + 0.999 * shadow[n]).clone()
            return float(loss)
        return step, "reference", "unmodified reference model files (oracle/_ref copy) + third-party shims, torch CPU"
    if cfg["kind"] != "direct":
        return None, None, "reference copy (oracle/_ref) missing and the oracle port covers only the OC20 model"
    from oracle import eqv2_oracle as O
    model = getattr(importlib.import_module(PKG + ".models." + cfg["module"]), cfg["cls"])(**kw)   # parameter container
    P = dict(model.named_parameters())
    hp = O.Hyper(lmax=kw["lmax_list"][0], mmax=kw["mmax_list"][0], C=kw["sphere_channels"], H=kw["attn_hidden_channels"],
                 heads=kw["num_heads"], alpha_ch=kw["attn_alpha_channels"], value_ch=kw["attn_value_channels"],
                 ffn_hidden=kw["ffn_hidden_channels"], edge_ch=kw["edge_channels"], num_layers=kw["num_layers"],
                 norm_type=kw["norm_type"], cutoff=kw["max_radius"], max_neighbors=kw["max_neighbors"])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)

    def step(data):
        n = len(data["natoms"])
        ei, dist, vec = O.radius_graph_pbc_fairchem(data["pos"], data["cell"], data["batch"], data["natoms"],
                                                    kw["max_radius"], kw["max_neighbors"])
        energy, forces = O.oc20_forward(P, hp, data["atomic_numbers"], data["batch"], n, ei, dist, vec,
                                        torch.rand(vec.shape) - 0.5)
        loss = (energy - data["energy"]).abs().mean() + (forces - data["forces"]).abs().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss)
    return step, "port", "oracle/eqv2_oracle.py port of the reference forward + autograd backward + AdamW"


def run_reference(args):
    """The reference's own CPU implementation of the path, all host threads, K steps of a bounded sample of the
    workload (`--ref-structures` structures of the config's shape per step)."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cfg = CONFIGS[args.config]
    kw = model_kw(cfg, args)
    world = args.gpus
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.ref_structures or cfg["ref_structures"]
    step, kind, note = reference_stepper(cfg, kw)
    structures = args.structures or cfg["structures"]
    base = {"impl": "reference", "metric": cfg["metric"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(cfg, kw, structures, world)}
    if step is None:
        print(json.dumps(dict(base, unavailable=note)), flush=True)
        return
    times = []
    for it in range(args.warmup + args.steps):
        data = make_batch(cfg, n, seed=1000 + it)
        t0 = time.perf_counter()
        step(data)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    med = sorted(times)[len(times) // 2]
    sample = (f"{n} of the {structures} structures per step ({int(data['pos'].shape[0])} atoms), {note}; "
              f"median step {med:.2f} s")
    line = dict(base, value=value, ms_per_step=1e3 * total / len(times),
                cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """The reference arm as a child process (its own interpreter: thread pool, sys.modules aliases of the reference
    package names and the torch.load patch stay out of the measured process): 1 warm-up + 3 timed steps, median-based."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--steps", "3",
           "--warmup", "1"]
    if args.layers:
        cmd += ["--layers", str(args.layers)]
    if args.no_dropout:
        cmd += ["--no-dropout"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env, timeout=900)
        line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        return line.get("cpu_baseline") or {"unavailable": line.get("unavailable")}
    except Exception as exc:     # noqa: BLE001 -- a missing baseline must not void the measured GPU line
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


# ------------------------------------------------------------------------------------------------
_BOUND = {      # entry point -> which roofline bounds it
    "eqv2_gemm_f16": "tensor", "eqv2_gemm_tc": "tensor", "eqv2_gemm_f32": "ffma",
    "eqv2_s2sep_fwd": "ffma", "eqv2_s2sep_bwd": "ffma", "eqv2_s2sep_bwd2": "ffma", "eqv2_s2act_fwd": "ffma",
    "eqv2_s2act_bwd": "ffma",
}


def roofline_of(name, r, pk, tot, note=None):
    bound = _BOUND.get(name, "hbm")
    ms = r["ms"]
    out = {"kernel": name, "bound": bound, "launches_per_step": r["calls"], "avg_launch_ms": ms / r["calls"],
           "share_of_step_kernel_time": ms / tot}
    if bound == "hbm":
        ach = r["bytes"] / (ms * 1e-3) / 1e9 if r["bytes"] > 0 else None
        out.update(achieved=ach, peak=pk["hbm"], unit="GB/s", frac=(ach / pk["hbm"]) if ach else None,
                   algorithmic_bytes_per_step=r["bytes"] or None, peak_source=pk["src"] + " HBM copy")
    else:
        peak = pk["tensor"] if bound == "tensor" else FFMA_PEAK_TFLOPS
        ach = r["flops"] / (ms * 1e-3) / 1e12 if r["flops"] > 0 else None
        out.update(achieved=ach, peak=peak, unit="TFLOP/s", frac=(ach / peak) if ach else None,
                   algorithmic_flops_per_step=r["flops"] or None,
                   peak_source=(pk["src"] + " bf16 sustained") if bound == "tensor" else
                   "FFMA peak measured by scripts/microbench/ffma2.cu")
    tr = TRAFFIC.get(name)
    out["traffic"] = tr[0] if tr else None
    if tr:
        out["traffic_source"] = tr[1]
    if note:
        out["note"] = note
    return out


def run_b200(args):
    import importlib
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    _lib = importlib.import_module(PKG + "._lib")
    _lib.lib()
    ops = importlib.import_module(PKG + ".ops")
    if args.gemm_mode:
        ops.set_gemm_mode(args.gemm_mode)
    for f in filter(None, args.disable.split(",")):
        ops._FEATURES[f] = False
    engine_note = {
        "f16x3": ("tcgen05 kind::f16 on operands pre-split into scaled fp16 hi/lo planes (3 passes, TMA, persistent CTAs, "
                  "fp32 register promotion: fp32-class accuracy); short reductions and degree slabs on the FFMA engine",
                  "fp32-accurate 3xFP16: 3 tensor-core passes at the fp16/bf16 rate -> ceiling = peak/3"),
        "f16": ("tcgen05 kind::f16 single pass on the hi plane (REDUCED precision, stated tolerance 2e-3)",
                "single fp16 pass"),
        "tf32x3": ("tcgen05 kind::tf32, 3xTF32 split with fp32 promotion (fp32-class accuracy); short reductions and "
                   "degree slabs on the FFMA engine",
                   "fp32-accurate 3xTF32: 3 tensor-core passes at the TF32 rate (1/2 of bf16) -> ceiling = peak/6"),
        "tf32": ("tcgen05 kind::tf32 single pass (REDUCED precision, stated tolerance 5e-3)", "TF32 rate = 1/2 of bf16"),
        "fp32": ("FFMA engine only", "no tensor cores"),
    }[ops.gemm_mode()]

    cfg = CONFIGS[args.config]
    kw = model_kw(cfg, args)
    torch.manual_seed(0)
    model = getattr(importlib.import_module(PKG + ".models." + cfg["module"]), cfg["cls"])(**kw).to(dev)
    net = model
    if world > 1 and args.no_graph:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    if args.optimizer == "fused":       # reference: clip_grad_norm_ -> AdamW -> EMA.update (train_oc20v2_parallel.py:177-186)
        optim = importlib.import_module(PKG + ".optim")
        opt = optim.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-3, max_grad_norm=100.0, ema_decay=0.999)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)

    B = args.structures or cfg["structures"]
    host = make_batch(cfg, B, seed=1000 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items() if torch.is_tensor(v)}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    resident = {k: v.to(dev) for k, v in pinned.items()}
    stats = {}

    def eager_step(data):
        loss = forward_loss(cfg, net, data)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    # The step (forward + loss + backward) is replayed from a CUDA graph; neighbour list / edge frames / AdamW stay eager
    # (graphs.py).  Data parallel: every rank replays its own graph on its own structures, then the gradients are
    # all-reduced in flat NCCL buckets.  --no-graph: eager launches (DistributedDataParallel when N > 1).
    use_graph = not args.no_graph
    stepper = None
    if use_graph:
        graphs = importlib.import_module(PKG + ".graphs")
        sync = None
        if world > 1:
            parallel = importlib.import_module(PKG + ".parallel")
            if args.grad_sync in ("overlap", "captured"):     # captured: inside the graph, but after the backward pass
                sync = parallel.OverlappedGradientAllReducer(model.parameters(), bucket_mb=args.sync_bucket_mb,
                                                             overlap=args.grad_sync == "overlap")
            else:
                sync = parallel.GradientAllReducer(model.parameters(), bucket_mb=64).reduce
        bucket = tuple(int(v) for v in args.bucket.split(",")) if args.bucket else None
        stepper = graphs.GraphedTrainStep(model, None, opt, grad_sync=sync, bucket=bucket,
                                          forward_loss=lambda d: forward_loss(cfg, model, d))
        step = stepper
    else:
        step = eager_step

    def step_e2e():
        data = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        return float(step(data).item())          # device -> host read of the step's result

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        stats["host_ms"] = 1e3 * (time.perf_counter() - t0) / n       # time the host needs to ENQUEUE one step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            stats["rank_ms_per_step"] = [round(float(t.item()) / n, 3) for t in every]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_e2e()
    with torch.no_grad():
        E = int(model.prepare(resident)["edge_index"].shape[1])
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.reset_launch_count()
    ms_dev = timed(lambda: step(resident), args.steps)
    host_ms = stats["host_ms"]
    rank_ms = stats.get("rank_ms_per_step")
    rank_edges = None
    if dist is not None:
        mine = torch.tensor([E], device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        rank_edges = [int(t.item()) for t in every]
    launches = _lib.launch_count()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler is not None else None

    # per-entry-point device time (separate pass: event pairs around every C-ABI call)
    barrier()
    if stepper is not None:
        launches += stepper.captured_launches() * args.steps      # kernels inside the replayed graphs
        stepper.release()               # the eager profiling pass below needs the memory of the graph pool
    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    _lib.start_kernel_timing()
    eager_step(resident)
    prof = _lib.stop_kernel_timing()

    n_struct = B * world
    value = n_struct * args.steps / (ms_dev / 1e3)
    e2e = n_struct * args.steps / (ms_e2e / 1e3)
    blocks = kw["num_layers"] + cfg["blocks_extra"]
    if rank == 0:
        pk = peaks()
        tot = sum(r["ms"] for r in prof.values()) or 1.0
        ranked = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
        shares = {k: round(v["ms"] / tot, 4) for k, v in ranked[:8]}
        rooflines = [roofline_of(k, r, pk, tot) for k, r in ranked[:8]]
        roof = roofline_of(ranked[0][0], ranked[0][1], pk, tot, note=engine_note[1] if ranked[0][0].startswith("eqv2_gemm") else None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args)
        conf = config_dict(cfg, kw, B, world)
        line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": conf,
                "workload_detail": {"atoms_per_gpu": int(host["pos"].shape[0]), "edges_per_gpu": E,
                                    "params": model.num_params,
                                    "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2),
                                    "launch": (("CUDA graph replay of forward+loss+backward" +
                                                (" with the bucketed NCCL gradient all-reduce captured inside the backward "
                                                 "pass on a side stream" if world > 1 and args.grad_sync == "overlap" else "") +
                                                "; neighbour list, edge frames" +
                                                (", gradient all-reduce (NCCL, flat buckets)" if world > 1 and
                                                 args.grad_sync != "overlap" else "") + " and AdamW eager") if use_graph
                                               else "eager (every kernel enqueued from Python; DDP when N > 1)"),
                                    "gemm_engine": engine_note[0],
                                    "optimizer": ("FusedAdamW: gradient-norm clip (100) + AdamW (wd 1e-3) + EMA (0.999) in one "
                                                  "multi-tensor pass" if args.optimizer == "fused"
                                                  else "torch.optim.AdamW(fused=True), no clip, no EMA")},
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms,
                "ranks": ({"ms_per_step": rank_ms, "edges": rank_edges, "grad_sync": args.grad_sync}
                          if world > 1 else None),
                "edge_msgs_per_s": E * world * blocks * args.steps / (ms_dev / 1e3),
                "roofline": roof, "rooflines": rooflines, "kernel_time_shares": shares, "clocks": clocks,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
