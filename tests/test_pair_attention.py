"""csrc/pair_attn.cu (all-pairs attention core of the global-attention classes, reference
NewFunctions/GATA_and_all2all/activation.py:1533-1547) against the per-structure torch restatement of the reference lines:
    attn = softmax(einsum('ihd,jhd->hij', q, k) * scale + bias);  out_l = einsum('hij,jmhd->imhd', attn, v_l)
values, first derivatives and the derivative of the backward pass (forces by autograd: create_graph=True)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import pkg


def _dense(q, k, values, counts, scale, bias):
    outs, start = [[] for _ in values], 0
    for g, n in enumerate(counts):
        sl = slice(start, start + n)
        attn = torch.einsum("ihd,jhd->hij", q[sl], k[sl]) * scale
        if bias is not None:
            attn = attn + bias[g]
        attn = F.softmax(attn, dim=-1)
        for o, v in zip(outs, values):
            o.append(torch.einsum("hij,jmhd->imhd", attn, v[sl]).reshape(n, v.shape[1], -1))
        start += n
    return [torch.cat(o) for o in outs]


@pytest.mark.parametrize("H,D,counts,ms", [(8, 16, [5, 1, 37, 12], [1, 3, 5]), (4, 8, [9, 2], [1]), (2, 64, [33], [7])])
def test_pair_attention_matches_dense_to_second_order(backend, H, D, counts, ms):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(sum(counts) + H)
    N = sum(counts)
    dd = torch.float64

    def mk(*shape):
        return torch.randn(*shape, generator=gen)

    q0, k0 = mk(N, H, D), mk(N, H, D)
    v0 = [mk(N, m, H, D) for m in ms]
    b0 = [0.3 * mk(H, n, n) for n in counts]
    go = [mk(N, m, H * D) for m in ms]
    w0 = mk(N, H, D)            # direction for the second-order check
    scale = D ** -0.5

    def run(dev, dtype, fn):
        q = q0.to(dev, dtype).requires_grad_(True)
        k = k0.to(dev, dtype).requires_grad_(True)
        vs = [v.to(dev, dtype).requires_grad_(True) for v in v0]
        bs = [b.to(dev, dtype).requires_grad_(True) for b in b0]
        outs = fn(q, k, vs, bs)
        loss = sum((o * g.to(dev, dtype)).sum() for o, g in zip(outs, go))
        first = torch.autograd.grad(loss, [q, k] + vs + bs, create_graph=True)
        # a "force-like" scalar of the first derivatives, differentiated again
        s2 = (first[0] * w0.to(dev, dtype)).sum() + sum((f * f).sum() for f in first[2:2 + len(vs)])
        second = torch.autograd.grad(s2, [q, k] + vs)
        return [o.detach().cpu().double() for o in outs], [f.detach().cpu().double() for f in first], \
            [s.detach().cpu().double() for s in second]

    mine = run(backend.device, torch.float32,
               lambda q, k, vs, bs: ops.pair_attention(q, k, vs, counts, scale, None, bs))
    ref = run(torch.device("cpu"), dd, lambda q, k, vs, bs: _dense(q, k, vs, counts, scale, bs))
    for name, a, b, tol in (("out", mine[0], ref[0], 2e-5), ("first", mine[1], ref[1], 5e-5), ("second", mine[2], ref[2], 2e-4)):
        for x, y in zip(a, b):
            err = float((x - y).abs().max()) / max(1.0, float(y.abs().max()))
            assert err < tol, (name, err)


def test_pair_attention_without_bias_and_single_value(backend):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(1)
    counts, H, D = [4, 6], 4, 4
    N = sum(counts)
    q, k, v = (torch.randn(N, H, D, generator=gen) for _ in range(3))
    (out,) = ops.pair_attention(backend.to(q), backend.to(k), [backend.to(v).reshape(N, 1, H, D)], counts, 0.5)
    (ref,) = _dense(q, k, [v.reshape(N, 1, H, D)], counts, 0.5, None)
    assert float((out.cpu() - ref).abs().max()) < 1e-5
