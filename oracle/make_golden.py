"""ORACLE tooling: generate tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference/models, imported through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container only:  python -m oracle.make_golden
The fixtures are committed; the GPU box never needs /root/reference.
Each fixture stores: hyper-parameters, the reference model's *parameters* (state_dict keys),
inputs, the graph the reference built, the random edge-frame draw it consumed
(edge_rot_mat.py:28), outputs, and parameter gradients of a scalar loss.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


class RandRecorder:
    """Records every torch.rand_like draw (the edge-frame helper vectors)."""

    def __enter__(self):
        self.draws = []
        self._orig = torch.rand_like

        def rec(*a, **k):
            out = self._orig(*a, **k)
            self.draws.append(out.clone())
            return out

        torch.rand_like = rec
        return self

    def __exit__(self, *a):
        torch.rand_like = self._orig


def synth_molecules(gen, num_graphs, nmin, nmax, elements):
    Z, pos, batch, natoms = [], [], [], []
    for g in range(num_graphs):
        n = int(torch.randint(nmin, nmax + 1, (1,), generator=gen))
        side = (9.0 * n) ** (1.0 / 3.0)
        while True:
            p = torch.rand(n, 3, generator=gen) * side
            d = torch.cdist(p, p) + torch.eye(n) * 10
            if d.min() > 0.7:
                break
        Z.append(torch.tensor(elements)[torch.randint(0, len(elements), (n,), generator=gen)])
        pos.append(p)
        batch.append(torch.full((n,), g, dtype=torch.long))
        natoms.append(n)
    return torch.cat(Z), torch.cat(pos), torch.cat(batch), torch.tensor(natoms)


def synth_cells(gen, num_graphs, n_atoms, vol_per_atom=15.0, zmax=89):
    Z, pos, batch, natoms, cells = [], [], [], [], []
    for g in range(num_graphs):
        side = (vol_per_atom * n_atoms) ** (1.0 / 3.0)
        cell = side * torch.eye(3) + 0.05 * side * torch.randn(3, 3, generator=gen)
        while True:
            frac = torch.rand(n_atoms, 3, generator=gen)
            p = frac @ cell
            ok = True
            for a in (-1, 0, 1):
                for b in (-1, 0, 1):
                    for c in (-1, 0, 1):
                        off = torch.tensor([a, b, c], dtype=torch.float32) @ cell
                        d = torch.cdist(p, p + off)
                        if (a, b, c) == (0, 0, 0):
                            d = d + torch.eye(n_atoms) * 10
                        ok = ok and bool(d.min() > 0.7)
            if ok:
                break
        Z.append(torch.randint(1, zmax + 1, (n_atoms,), generator=gen))
        pos.append(p)
        batch.append(torch.full((n_atoms,), g, dtype=torch.long))
        natoms.append(n_atoms)
        cells.append(cell)
    return torch.cat(Z), torch.cat(pos), torch.cat(batch), torch.tensor(natoms), torch.stack(cells)


def params_of(model):
    return {k: v.detach().clone() for k, v in model.named_parameters()}


def golden_oc20():
    from equiformerv2_oc20 import EquiformerV2_OC20
    import fairchem.core.graph.compute as fc
    hp = dict(lmax=3, mmax=2, C=16, H=8, heads=2, alpha_ch=8, value_ch=4, ffn_hidden=16, edge_ch=16,
              num_layers=2, norm_type="rms_norm_sh", grid_res=18, num_rbf=600, cutoff=6.0,
              max_elements=90, max_neighbors=8)
    for norm_type in ("rms_norm_sh", "layer_norm_sh", "layer_norm"):
        torch.manual_seed(11)
        model = EquiformerV2_OC20(
            max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=90,
            num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"],
            num_heads=hp["heads"], attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"],
            ffn_hidden_channels=hp["ffn_hidden"], norm_type=norm_type, lmax_list=[hp["lmax"]],
            mmax_list=[hp["mmax"]], grid_resolution=18, edge_channels=hp["edge_ch"],
            alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
        # embeddings are ~1e-3 at init; perturb all params so every path carries signal
        gen = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=gen))
        Z, pos, batch, natoms, cell = synth_cells(gen, 2, 6, zmax=89)
        data = dict(atomic_numbers=Z, pos=pos, batch=batch, natoms=natoms, cell=cell)
        captured = {}
        orig = fc.generate_graph
        import equiformerv2_oc20 as mod

        def spy(**kw):
            out = orig(**kw)
            captured.update(out)
            return out

        mod.generate_graph = spy
        with RandRecorder() as rr:
            energy, forces = model(data)
        mod.generate_graph = orig
        loss = energy.sum() + (forces * torch.linspace(-1, 1, forces.numel()).view_as(forces)).sum()
        loss.backward()
        fx = dict(hyper=dict(hp, norm_type=norm_type), params=params_of(model),
                  grads={k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
                  inputs=data, edge_index=captured["edge_index"], edge_distance=captured["edge_distance"],
                  edge_vec=captured["edge_distance_vec"], rand_vec=rr.draws[0] - 0.5,
                  energy=energy.detach(), forces=forces.detach())
        torch.save(fx, os.path.join(OUT, f"oc20_small_{norm_type}.pt"))
        print("oc20", norm_type, "E", captured["edge_index"].shape[1], energy.detach())


def golden_qm9():
    from equiformerv2_qm9 import EquiformerV2_QM9
    hp = dict(lmax=2, mmax=2, C=16, H=8, heads=2, alpha_ch=8, value_ch=4, ffn_hidden=16, edge_ch=16,
              num_layers=2, norm_type="rms_norm_sh", grid_res=18, num_rbf=600, cutoff=5.0,
              max_elements=10, max_neighbors=6, num_targets=3)
    torch.manual_seed(3)
    model = EquiformerV2_QM9(
        num_targets=3, max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=10,
        num_layers=2, sphere_channels=16, attn_hidden_channels=8, num_heads=2, attn_alpha_channels=8,
        attn_value_channels=4, ffn_hidden_channels=16, lmax_list=[2], mmax_list=[2], grid_resolution=18,
        edge_channels=16, alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    Z, pos, batch, natoms = synth_molecules(gen, 3, 5, 9, [1, 6, 7, 8, 9])
    data = dict(atomic_numbers=Z, pos=pos, batch=batch, natoms=natoms)
    with RandRecorder() as rr:
        ei, dist, vec, *_ = model.generate_graph(data)
        pred = model(data)
    loss = (pred * torch.linspace(-1, 1, pred.numel()).view_as(pred)).sum()
    loss.backward()
    fx = dict(hyper=hp, params=params_of(model),
              grads={k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
              inputs=data, edge_index=ei, edge_distance=dist, edge_vec=vec, rand_vec=rr.draws[0] - 0.5,
              pred=pred.detach())
    torch.save(fx, os.path.join(OUT, "qm9_small.pt"))
    print("qm9 E", ei.shape[1], pred.detach())


def golden_components():
    """Component-level vectors from the reference classes (third-party shims underneath)."""
    from EquiformerV2Functions.so3 import SO3_Rotation, SO3_Grid
    from EquiformerV2Functions.edge_rot_mat import init_edge_rot_mat
    gen = torch.Generator().manual_seed(21)
    vec = torch.randn(40, 3, generator=gen) * 2.0
    torch.manual_seed(99)
    with RandRecorder() as rr:
        R = init_edge_rot_mat(vec)
    out = dict(edge_vec=vec, rand_vec=rr.draws[0] - 0.5, rot=R)
    for lmax in (2, 4, 6):
        rot = SO3_Rotation(lmax)
        rot.set_wigner(R)
        out[f"wigner_l{lmax}"] = rot.wigner.clone()
    for (l, m) in ((4, 2), (4, 4), (6, 2), (6, 6), (2, 2), (3, 2), (3, 3)):
        g = SO3_Grid(l, m, resolution=18, normalization="component")
        out[f"to_grid_{l}_{m}"] = g.to_grid_mat.clone()
        out[f"from_grid_{l}_{m}"] = g.from_grid_mat.clone()
    torch.save(out, os.path.join(OUT, "components.pt"))
    print("components ok")


def golden_matpes():
    """MatPES v1/v2 graphs and the v2 model's train-step quantities (train_MatPES_GATAWandB.py:67-91 pattern:
    forces by autograd.grad(create_graph=True), loss on energy + forces, loss.backward())."""
    import importlib
    v1 = importlib.import_module("equiformerv2_MatPES")
    v2 = importlib.import_module("equiformerv2_MatPESv2")
    gen = torch.Generator().manual_seed(13)
    Z, pos, batch, natoms, cell = synth_cells(gen, 2, 7, vol_per_atom=14.0, zmax=89)
    data = dict(atomic_numbers=Z, pos=pos, batch=batch, natoms=natoms, cell=cell)
    hp = dict(lmax=3, mmax=2, C=16, H=8, heads=2, alpha_ch=8, value_ch=4, ffn_hidden=16, edge_ch=16, num_layers=2,
              norm_type="rms_norm_sh", grid_res=18, num_rbf=600, cutoff=4.5, max_elements=100, max_neighbors=8)
    kw = dict(max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=100, num_layers=2,
              sphere_channels=16, attn_hidden_channels=8, num_heads=2, attn_alpha_channels=8, attn_value_channels=4,
              ffn_hidden_channels=16, lmax_list=[3], mmax_list=[2], grid_resolution=18, edge_channels=16,
              alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    torch.manual_seed(4)
    m1 = v1.EquiformerV2_MatPES(regress_forces=False, regress_stress=False, **kw)
    ei1, d1, vec1, *_ = m1.generate_graph(data)
    torch.manual_seed(4)
    model = v2.EquiformerV2_MatPES(**kw)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    ei2, d2, vec2 = model.generate_graph(pos, batch, cell)
    posg = pos.clone().requires_grad_(True)
    out = model(dict(data, pos=posg))
    forces = -torch.autograd.grad(out["energy_total"].sum(), posg, create_graph=True, retain_graph=True)[0]
    wf = torch.linspace(-1, 1, forces.numel()).view_as(forces)
    we = torch.linspace(0.5, 1.5, out["energy"].numel()).view_as(out["energy"])
    loss = (out["energy"] * we).sum() + (forces * wf).sum()
    loss.backward()
    fx = dict(hyper=hp, params=params_of(model),
              grads={k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
              inputs=data, v1_edge_index=ei1, v1_edge_distance=d1.detach(), v1_edge_vec=vec1.detach(),
              edge_index=ei2, edge_distance=d2.detach(), edge_vec=vec2.detach(),
              energy=out["energy"].detach(), energy_total=out["energy_total"].detach(), forces=forces.detach())
    torch.save(fx, os.path.join(OUT, "matpes_v2_small.pt"))
    print("matpes E1", ei1.shape[1], "E2", ei2.shape[1], out["energy"].detach().view(-1), forces.abs().max().item(),
          "self edges", int((ei2[0] == ei2[1]).sum()))


def golden_matpes_v1():
    """BASELINE config 3 as named: equiformerv2_MatPES.py (v1), forces pass and stress pass separately (the combined
    default crashes in the reference)."""
    import importlib
    v1 = importlib.import_module("equiformerv2_MatPES")
    gen = torch.Generator().manual_seed(23)
    Z, pos, batch, natoms, cell = synth_cells(gen, 2, 6, vol_per_atom=14.0, zmax=89)
    data = dict(atomic_numbers=Z, pos=pos, batch=batch, natoms=natoms, cell=cell)
    hp = dict(lmax=3, mmax=2, C=16, H=8, heads=2, alpha_ch=8, value_ch=4, ffn_hidden=16, edge_ch=16, num_layers=2,
              norm_type="rms_norm_sh", grid_res=18, num_rbf=600, cutoff=4.5, max_elements=100, max_neighbors=8)
    kw = dict(max_neighbors=8, max_radius=4.5, max_num_elements=100, num_layers=2, sphere_channels=16,
              attn_hidden_channels=8, num_heads=2, attn_alpha_channels=8, attn_value_channels=4, ffn_hidden_channels=16,
              lmax_list=[3], mmax_list=[2], grid_resolution=18, edge_channels=16, alpha_drop=0.0, drop_path_rate=0.0,
              proj_drop=0.0)
    torch.manual_seed(6)
    mf = v1.EquiformerV2_MatPES(regress_forces=True, regress_stress=False, **kw)
    with torch.no_grad():
        for p in mf.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    ms = v1.EquiformerV2_MatPES(regress_forces=False, regress_stress=True, **kw)
    ms.load_state_dict(mf.state_dict())
    ms.eval()
    torch.manual_seed(77)
    with RandRecorder() as rr:
        of = mf(dict(data, pos=pos.clone()))
    torch.manual_seed(77)
    os_ = ms(dict(data, pos=pos.clone()))
    ei, d, vec, *_ = mf.generate_graph(data)
    fx = dict(hyper=hp, params=params_of(mf), inputs=data, edge_index=ei, edge_distance=d.detach(), edge_vec=vec.detach(),
              rand_vec=rr.draws[0] - 0.5, energy=of["energy"].detach(), forces=of["forces"].detach(),
              energy_stress_pass=os_["energy"].detach(), stress=os_["stress"].detach())
    # The same reference evaluated in float64 (its fp32 result carries ~1e-5 of rounding noise in the autograd forces and
    # stress, so parity with the fp32 numbers is only meaningful down to that floor).  The reference hard-codes
    # dtype=torch.float32 in one torch.arange (equiformerv2_MatPES.py:275); that single dtype is redirected here.
    real_arange, real_rand_like = torch.arange, torch.rand_like
    torch.arange = lambda *a, **k: real_arange(*a, **{**k, "dtype": torch.float64 if k.get("dtype") == torch.float32 else k.get("dtype")})
    torch.rand_like = lambda t, *a, **k: rr.draws[0].to(t.dtype)
    torch.set_default_dtype(torch.float64)
    try:
        data64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()}
        of64 = mf.double()(dict(data64, pos=data64["pos"].clone()))
        os64 = ms.double()(dict(data64, pos=data64["pos"].clone()))
    finally:
        torch.arange, torch.rand_like = real_arange, real_rand_like
        torch.set_default_dtype(torch.float32)
    fx.update(energy_f64=of64["energy"].detach(), forces_f64=of64["forces"].detach(), stress_f64=os64["stress"].detach())
    torch.save(fx, os.path.join(OUT, "matpes_v1_small.pt"))
    print("matpes v1 E", ei.shape[1], of["energy"].detach().view(-1), os_["stress"].detach()[0])


def golden_gata(modname="equiformerv2_MatPES_GATAV2", out_name="matpes_gatav2_small.pt"):
    """BASELINE config 4 family: equiformerv2_MatPES_GATAV2.py (HTR + GATA value activation) and its
    phi-at-every-iteration twin, train-step pattern."""
    import importlib
    mod = importlib.import_module(modname)
    gen = torch.Generator().manual_seed(17)
    Z, pos, batch, natoms, cell = synth_cells(gen, 2, 6, vol_per_atom=14.0, zmax=89)
    data = dict(atomic_numbers=Z, pos=pos, batch=batch, natoms=natoms, cell=cell)
    hp = dict(lmax=3, mmax=3, C=16, H=8, heads=2, alpha_ch=8, value_ch=4, ffn_hidden=16, edge_ch=16, num_layers=2,
              norm_type="rms_norm_sh", grid_res=18, num_rbf=600, cutoff=4.5, max_elements=100, max_neighbors=8)
    torch.manual_seed(8)
    model = mod.EquiformerV2_MatPES(
        max_neighbors=8, max_radius=4.5, max_num_elements=100, num_layers=2, sphere_channels=16, attn_hidden_channels=8,
        num_heads=2, attn_alpha_channels=8, attn_value_channels=4, ffn_hidden_channels=16, lmax_list=[3], mmax_list=[3],
        grid_resolution=18, edge_channels=16, alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    ei, d, vec = model.generate_graph(pos, batch, cell)
    posg = pos.clone().requires_grad_(True)
    out = model(dict(data, pos=posg))
    forces = -torch.autograd.grad(out["energy_total"].sum(), posg, create_graph=True, retain_graph=True)[0]
    wf = torch.linspace(-1, 1, forces.numel()).view_as(forces)
    we = torch.linspace(0.5, 1.5, out["energy"].numel()).view_as(out["energy"])
    ((out["energy"] * we).sum() + (forces * wf).sum()).backward()
    fx = dict(hyper=hp, params=params_of(model),
              grads={k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
              inputs=data, edge_index=ei, edge_distance=d.detach(), edge_vec=vec.detach(),
              energy=out["energy"].detach(), energy_total=out["energy_total"].detach(), forces=forces.detach())
    torch.save(fx, os.path.join(OUT, out_name))
    nograd = [k for k, p in model.named_parameters() if p.grad is None]
    print("gata E", ei.shape[1], out["energy"].detach().view(-1), "params without grad:", len(nograd))


if __name__ == "__main__":
    ref_loader.install()
    os.makedirs(OUT, exist_ok=True)
    golden_components()
    golden_oc20()
    golden_qm9()
    golden_matpes()
    golden_matpes_v1()
    golden_gata()
    golden_gata("equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata", "matpes_gatav2_phi_small.pt")
    golden_gata("equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE",
                "matpes_gatav2_global_small.pt")
