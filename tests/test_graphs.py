"""GPU only: a train step replayed from a CUDA graph (graphs.GraphedTrainStep: forward + loss + backward captured,
neighbour list / edge frames / optimizer eager) must follow the same trajectory as the eager step -- losses and
parameters after several updates, including a change of batch signature (re-capture) and a return to the first one."""
import pytest
import torch

from helpers import pkg


def _loss(out, d):
    return (out[0] - d["energy"]).abs().mean() + (out[1] - d["forces"]).abs().mean()


def _run(graphed, batches, mode=None):
    oc20, graphs, ops = pkg("models.equiformerv2_oc20"), pkg("graphs"), pkg("ops")
    ops.set_gemm_mode(mode or ops.DEFAULT_GEMM_MODE)
    torch.manual_seed(0)
    model = oc20.EquiformerV2_OC20(num_layers=2, sphere_channels=32, attn_hidden_channels=16, num_heads=2,
                                   attn_alpha_channels=16, attn_value_channels=8, ffn_hidden_channels=32, lmax_list=[3],
                                   mmax_list=[2], edge_channels=32, alpha_drop=0.0, drop_path_rate=0.0, max_radius=8.0).cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    sync = None
    if graphed == "captured_sync":      # the data-parallel exchange object, captured with the backward pass (world size 1:
        sync = pkg("parallel").OverlappedGradientAllReducer(model.parameters(), bucket_mb=0.25)    # no collective issued)
    stepper = graphs.GraphedTrainStep(model, _loss, opt, grad_sync=sync) if graphed else None
    torch.manual_seed(3)
    out = []
    for d in batches:
        if graphed:
            out.append(float(stepper(d)))
        else:
            loss = _loss(model(d), d)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            out.append(float(loss))
    if graphed:
        assert stepper.replays == len(batches) and len(stepper.graphs) == 2
    if sync is not None:                # the gradients the optimizer read LIVE in the flat buckets, sent from inside backward
        assert len(sync.buckets) >= 3
        for p in model.parameters():
            assert p.grad is not None and p.grad.data_ptr() == sync.view[id(p)].data_ptr()
    return out, [p.detach().clone() for p in model.parameters()]


@pytest.mark.gpu
def test_graphed_step_follows_eager_trajectory():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg("_lib")._state["lib"] = None
    syn = pkg("synthetic")
    a = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.oc20_batch(2, seed=5).items()}
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.oc20_batch(3, seed=6).items()}
    batches = [a, a, b, a, b]
    l0, p0 = _run(False, batches)
    l1, p1 = _run(True, batches)
    assert max(abs(x - y) / abs(x) for x, y in zip(l0, l1)) < 1e-5, (l0, l1)
    worst = max(float((x - y).abs().max() / (x.abs().max() + 1e-12)) for x, y in zip(p0, p1))
    assert worst < 1e-4, worst
    # with the bucketed gradient exchange captured inside the replayed backward pass (parallel.OverlappedGradientAllReducer)
    l3, p3 = _run("captured_sync", batches)
    assert max(abs(x - y) / abs(x) for x, y in zip(l0, l3)) < 1e-5, (l0, l3)
    assert max(float((x - y).abs().max() / (x.abs().max() + 1e-12)) for x, y in zip(p0, p3)) < 1e-4
    # the same trajectory on the exact FFMA engine: catches host-side caches that go stale across optimizer updates
    # (fused AdamW does not bump Tensor._version) -- several steps, not just one
    try:
        l2, _ = _run(False, batches, mode="fp32")
    finally:
        pkg("ops").set_gemm_mode(pkg("ops").DEFAULT_GEMM_MODE)
    assert max(abs(x - y) / abs(x) for x, y in zip(l0, l2)) < 1e-4, (l0, l2)


def _matpes_forward_loss(model, d):
    pos = d["pos"].detach().requires_grad_(True)
    out = model(dict(d, pos=pos))
    forces = -torch.autograd.grad(out["energy_total"].sum(), pos, create_graph=True, retain_graph=True)[0]
    return (out["energy"] - d["energy"]).abs().mean() + (forces - d["forces"]).abs().mean()


@pytest.mark.gpu
def test_graphed_matpes_double_backward_step_follows_eager_trajectory():
    """MatPES pattern (train_MatPES_GATAWandB.py:67-91): forces by autograd.grad(create_graph=True) inside the loss, then
    loss.backward() -- captured as one graph (forward + force gradient + double backward) and replayed."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg("_lib")._state["lib"] = None
    syn, mp, graphs = pkg("synthetic"), pkg("models.equiformerv2_MatPESv2"), pkg("graphs")
    data = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.matpes_batch(2, seed=3).items()}
    res = []
    for graphed in (False, True):
        torch.manual_seed(0)
        model = mp.EquiformerV2_MatPES(num_layers=2, sphere_channels=32, attn_hidden_channels=32, num_heads=2,
                                       attn_alpha_channels=16, attn_value_channels=8, ffn_hidden_channels=64, lmax_list=[3],
                                       mmax_list=[2], edge_channels=32, alpha_drop=0.0, drop_path_rate=0.0).cuda()
        # plain SGD: Adam's g / sqrt(v) turns rounding-level differences of near-zero gradients (split-K atomics) into
        # lr-sized parameter differences, which says nothing about the replay
        opt = torch.optim.SGD(model.parameters(), lr=1e-2)
        stepper = graphs.GraphedTrainStep(model, None, opt, forward_loss=lambda d, m=model: _matpes_forward_loss(m, d))
        losses = []
        for _ in range(4):
            if graphed:
                losses.append(float(stepper(data).detach()))
            else:
                loss = _matpes_forward_loss(model, data)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
        res.append((losses, [p.detach().clone() for p in model.parameters()]))
    (l0, p0), (l1, p1) = res
    assert max(abs(x - y) / abs(x) for x, y in zip(l0, l1)) < 1e-5, (l0, l1)
    assert max(float((x - y).abs().max() / (x.abs().max() + 1e-12)) for x, y in zip(p0, p1)) < 1e-4
