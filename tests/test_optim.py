"""FusedAdamW (clip + AdamW + EMA in one multi-tensor pass, csrc/optim.cu) against the reference's optimizer-side step
(train_oc20v2_parallel.py:95-126,177-186): torch.nn.utils.clip_grad_norm_ -> torch.optim.AdamW(param groups from
add_weight_decay) -> ExponentialMovingAverage.update, run on the CPU in fp32."""
import pytest
import torch

from helpers import pkg


def _make(seed, device):
    gen = torch.Generator().manual_seed(seed)
    shapes = [(3,), (17, 5), (40000,), (7, 3, 129), (1,), (16385,)]       # below / across / above one 16 384-element chunk
    return [torch.randn(s, generator=gen).to(device).requires_grad_(True) for s in shapes]


def _reference_steps(params, grads_per_step, lr, wd, clip, ema_decay, betas, eps):
    decay, no_decay = params[1::2], params[0::2]
    opt = torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": wd}], lr=lr,
                            betas=betas, eps=eps)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    shadow = [p.data.clone() for p in params]
    norms = []
    for grads in grads_per_step:
        for p, g in zip(params, grads):
            p.grad = None if g is None else g.clone()
        if clip > 0:
            norms.append(float(torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], clip)))
        opt.step()
        sched.step()
        if ema_decay > 0:       # reference ExponentialMovingAverage.update
            shadow = [(1.0 - ema_decay) * p.data + ema_decay * s for p, s in zip(params, shadow)]
    return shadow, norms


@pytest.mark.parametrize("clip,ema_decay", [(0.0, 0.0), (0.5, 0.999), (100.0, 0.9)])
def test_fused_adamw_matches_torch_adamw_clip_and_reference_ema(backend, clip, ema_decay):
    optim = pkg("optim")
    lr, wd, betas, eps = 2e-3, 1e-2, (0.9, 0.99), 1e-8
    ref_params = _make(1, "cpu")
    mine = [p.detach().clone().to(backend.device).requires_grad_(True) for p in ref_params]
    gen = torch.Generator().manual_seed(2)
    steps = []
    for s in range(6):
        gs = [torch.randn(p.shape, generator=gen) * (3.0 if s == 2 else 0.3) for p in ref_params]
        if s in (1, 4):
            gs[0] = None            # a parameter without gradient in some steps (GATA family: SURVEY 0.11)
        steps.append(gs)
    shadow, norms = _reference_steps(ref_params, steps, lr, wd, clip, ema_decay, betas, eps)

    opt = optim.FusedAdamW([{"params": mine[0::2], "weight_decay": 0.0}, {"params": mine[1::2], "weight_decay": wd}],
                           lr=lr, betas=betas, eps=eps, max_grad_norm=clip, ema_decay=ema_decay)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    got_norms = []
    for gs in steps:
        for p, g in zip(mine, gs):
            p.grad = None if g is None else g.clone().to(backend.device)
        opt.step()
        sched.step()
        if clip > 0:
            got_norms.append(float(opt.grad_norm[0]))
    for a, b in zip(mine, ref_params):
        assert float((a.detach().cpu() - b.detach()).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max()))
    if clip > 0:
        assert all(abs(x - y) <= 1e-5 * y for x, y in zip(got_norms, norms))
    if ema_decay > 0:
        view = opt.ema([(str(i), p) for i, p in enumerate(mine)])
        for i, s in enumerate(shadow):
            assert float((view.shadow[str(i)].cpu() - s).abs().max()) <= 2e-6 * max(1.0, float(s.abs().max()))
        # store / copy_to / restore round trip, as the reference evaluates with EMA weights
        before = [p.detach().clone() for p in mine]
        view.store()
        view.copy_to()
        assert all(torch.equal(p.detach(), view.shadow[str(i)]) for i, p in enumerate(mine))
        view.restore()
        assert all(torch.equal(p.detach(), b) for p, b in zip(mine, before))
    sd = opt.state_dict()
    assert {"step", "exp_avg", "exp_avg_sq"} <= set(sd["state"][0])
