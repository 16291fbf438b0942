"""Input side: pinned collate in the reference schema (data_loader_matpes.py:290-314) and padding to
(atoms, edges) buckets with a masked ghost structure -- which must leave real outputs and parameter gradients unchanged."""
import pytest
import torch

from conftest import golden
from helpers import build_matpes_v2, build_oc20, fixed_rand_like, load_params, pkg, rel_err


def test_collate_matches_reference_schema():
    batching = pkg("batching")
    gen = torch.Generator().manual_seed(0)
    structs = []
    for n in (3, 5, 2):
        structs.append(dict(atomic_numbers=torch.randint(1, 90, (n,), generator=gen), pos=torch.randn(n, 3, generator=gen),
                            cell=torch.eye(3) * (4.0 + n), energy=torch.randn(1, generator=gen),
                            forces=torch.randn(n, 3, generator=gen), stress=torch.randn(6, generator=gen)))
    out = batching.collate(structs, pin=False)
    # the reference collate (data_loader_matpes.py:290-314), restated
    assert torch.equal(out["atomic_numbers"], torch.cat([s["atomic_numbers"] for s in structs]))
    assert torch.equal(out["pos"], torch.cat([s["pos"] for s in structs]))
    assert torch.equal(out["forces"], torch.cat([s["forces"] for s in structs]))
    assert torch.equal(out["energy"], torch.stack([s["energy"] for s in structs]))
    assert torch.equal(out["natoms"], torch.tensor([3, 5, 2]))
    assert torch.equal(out["batch"], torch.tensor([0, 0, 0, 1, 1, 1, 1, 1, 2, 2]))
    assert torch.equal(out["cell"], torch.stack([s["cell"] for s in structs]))
    assert out["pbc"].shape == (3, 3) and bool(out["pbc"].all())
    assert torch.equal(out["stress"], torch.stack([s["stress"] for s in structs]))
    assert out["energy"].shape == (3, 1) and out["batch"].dtype == torch.long


def test_bucket_sizes():
    batching = pkg("batching")
    assert batching.bucket_sizes(640, 13489, 64, 512) == (704, 13824)
    assert batching.bucket_sizes(62, 512, 64, 512) == (64, 512)
    assert batching.bucket_sizes(63, 513, 64, 512) == (128, 1024)


def test_padded_oc20_batch_gives_the_same_outputs_and_gradients(backend):
    batching = pkg("batching")
    fx = golden("oc20_small_rms_norm_sh.pt")
    data = backend.to(dict(fx["inputs"], edge_index=fx["edge_index"], edge_distance=fx["edge_distance"],
                           edge_distance_vec=fx["edge_vec"]))

    def run(pad):
        model = build_oc20(fx["hyper"], backend.device)
        load_params(model, fx["params"])
        with fixed_rand_like(fx["rand_vec"] + 0.5), torch.no_grad():
            full = dict(data)
            full.update(model.prepare(data))
        if pad:
            full = batching.pad_to_bucket(full, 16, 64)
            N, E = data["pos"].shape[0], fx["edge_index"].shape[1]
            assert full["pos"].shape[0] % 16 == 0 and full["edge_index"].shape[1] % 64 == 0
            assert full["pos"].shape[0] >= N + 2 and len(full["natoms"]) == len(data["natoms"]) + 1
            assert bool((full["edge_index"][1][1:] >= full["edge_index"][1][:-1]).all())       # still destination-sorted
            assert float(full["atom_mask"].sum()) == N and float(full["structure_mask"].sum()) == len(data["natoms"])
        energy, forces = model(full)
        B, N = len(data["natoms"]), data["pos"].shape[0]
        w = torch.linspace(-1, 1, 3 * N, device=forces.device).view(N, 3)
        (energy[:B].sum() + (forces[:N] * w).sum()).backward()
        return energy[:B].detach(), forces[:N].detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    e0, f0, g0 = run(False)
    e1, f1, g1 = run(True)
    assert rel_err(e0, fx["energy"]) < 1e-5 and rel_err(f0, fx["forces"]) < 1e-5
    assert rel_err(e1, e0) < 1e-6 and rel_err(f1, f0) < 1e-6
    scale = max(float(g.abs().max()) for g in g0.values())
    for k in g0:
        assert float((g1[k] - g0[k]).abs().max()) <= 2e-6 * max(float(g0[k].abs().max()), 1e-7 * scale), k


def test_padded_matpes_batch_double_backward_is_unchanged(backend):
    """MatPES pattern: forces by autograd through the padded batch (the ghost atoms' edge vectors are recomputed from
    their positions inside the model), loss masked, double backward."""
    if backend.name == "emu":
        pytest.skip("CPU suite budget (2 minutes under the emulator): the first-order padded-batch test runs on the "
                    "emulator, this double-backward one on the GPU")
    batching = pkg("batching")
    fx = golden("matpes_v2_small.pt")
    data = backend.to(dict(fx["inputs"]))

    def run(pad):
        model = build_matpes_v2(fx["hyper"], backend.device)
        load_params(model, fx["params"])
        full = dict(data)
        with torch.no_grad():
            full.update(model.prepare(data))
        if pad:
            full = batching.pad_to_bucket(full, 8, 32)
        N, B = data["pos"].shape[0], len(data["natoms"])
        pos = full["pos"].clone().requires_grad_(True)
        out = model(dict(full, pos=pos))
        forces = -torch.autograd.grad(out["energy_total"][:B].sum(), pos, create_graph=True, retain_graph=True)[0]
        wf = torch.linspace(-1, 1, 3 * N, device=forces.device).view(N, 3)
        we = torch.linspace(0.5, 1.5, B, device=forces.device).view(B, 1)
        ((out["energy"][:B] * we).sum() + (forces[:N] * wf).sum()).backward()
        return (out["energy"][:B].detach(), forces[:N].detach(),
                {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})

    e0, f0, g0 = run(False)
    e1, f1, g1 = run(True)
    assert rel_err(e0, fx["energy"]) < 1e-5 and rel_err(f0, fx["forces"]) < 1e-5
    assert rel_err(e1, e0) < 1e-6 and rel_err(f1, f0) < 2e-6
    assert set(g0) == set(g1)
    scale = max(float(g.abs().max()) for g in g0.values())
    for k in g0:
        assert float((g1[k] - g0[k]).abs().max()) <= 5e-6 * max(float(g0[k].abs().max()), 1e-7 * scale), k


@pytest.mark.gpu
def test_bucketed_graph_replay_serves_varying_batches():
    """40 training steps over batches whose (atoms, edges) differ every step: with buckets, >= 90 % of the steps replay
    one of <= 6 captured graphs; the loss of every step equals the eager, unpadded step on the same weights."""
    import copy
    graphs, batching, syn = pkg("graphs"), pkg("batching"), pkg("synthetic")
    dev = torch.device("cuda:0")
    fx = golden("oc20_small_rms_norm_sh.pt")
    hp = dict(fx["hyper"], max_neighbors=12)
    model = build_oc20(hp, dev)
    load_params(model, fx["params"])
    ref_model = copy.deepcopy(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)            # weights stay put: every step is comparable
    ref_opt = torch.optim.SGD(ref_model.parameters(), lr=0.0)

    def loss_fn(m, d):
        energy, forces = m(d)
        if "atom_mask" in d:
            return (batching.masked_mean((energy - d["energy"]).abs(), d["structure_mask"]) +
                    batching.masked_mean((forces - d["forces"]).abs(), d["atom_mask"]))
        return (energy - d["energy"]).abs().mean() + (forces - d["forces"]).abs().mean()

    stepper = graphs.GraphedTrainStep(model, None, opt, forward_loss=lambda d: loss_fn(model, d), bucket=(32, 256),
                                      max_graphs=6)
    gen = torch.Generator().manual_seed(3)
    sigs = set()
    for it in range(40):
        batch = syn.matpes_batch(2, seed=100 + it, n_atoms=int(torch.randint(18, 30, (1,), generator=gen)))
        data = {k: v.to(dev) for k, v in batch.items()}
        data["energy"] = data["energy"].view(-1)
        torch.manual_seed(1000 + it)                      # same edge-frame draw on both paths
        loss = stepper(data)
        torch.manual_seed(1000 + it)
        ref = loss_fn(ref_model, data)
        assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref)), (it, float(loss), float(ref))
        sigs.add((data["pos"].shape[0],))
    assert len(sigs) >= 6                                  # the batches really differed
    assert len(stepper.graphs) <= 6 and stepper.eager_steps == 0, (len(stepper.graphs), stepper.eager_steps)
    assert stepper.replays >= 36
