"""Global (all-to-all) node attention variants of reference NewFunctions/GATA_and_all2all/activation.py:419-1567:
`GlobalNodeAttention` (:419), `GlobalNodeAttentionFullEquivariant` (:686), `GlobalNodeAttentionHTR` (:1025),
`GlobalNodeAttentionHTR_with_distance` (:1217) and `GlobalNodeAttentionHTR_with_ROPE` (:1377, the one BASELINE config 5
instantiates, equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_..._with_DISTANCE.py:231-237).  Same constructor
arguments, parameter names / state_dict keys and forward signatures.

The reference works on [N_tot, N_tot] maps over ALL atoms of the batch with cross-graph pairs masked (the HTR variants
materialise score[N_tot, N_tot, C]: 5.2 GB at 16 x 200 atoms) or on tensors padded to [B, N_max]; here every variant
works per structure, and the HTR variants use only the two marginals of the score tensor they need
(`_HTRGlobalAttention`).  The Linear layers run on the GEMM kernels (`ops.linear`); the per-structure n x n attention
products are plain library matmuls -- a fused flash-style kernel with a twice-differentiable backward is the open item
of this file.  Parity with the unmodified reference classes: tests/test_global_attention.py."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..Gotennets_GATA_phi_refined_every_layer.activation import *  # noqa: F401,F403
from ..Gotennets_GATA_phi_refined_every_layer.activation import GATAValueActivation, HTR  # noqa: F401


def _lin(layer, x):
    shp = x.shape
    y = ops.linear(x.reshape(-1, shp[-1]), layer.weight, layer.bias)
    return y.reshape(*shp[:-1], layer.weight.shape[0])


def real_sh_integral(lmax, xyz):
    """Real SH of unit vectors, l = 0..lmax, 'integral' normalisation, basis of SURVEY App. B.1 (polar axis y,
    azimuth atan2(x, z), no Condon-Shortley phase), as polynomials in (x, y, z): [..., (lmax+1)^2], l-major."""
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    cm, sm = [torch.ones_like(x)], [torch.zeros_like(x)]          # Re / Im of (z + i x)^m
    for m in range(1, lmax + 1):
        cm.append(cm[m - 1] * z - sm[m - 1] * x)
        sm.append(sm[m - 1] * z + cm[m - 1] * x)
    q = {}
    for m in range(lmax + 1):
        q[(m, m)] = torch.full_like(y, float(math.prod(range(1, 2 * m, 2))))
        if m < lmax:
            q[(m + 1, m)] = (2 * m + 1) * y * q[(m, m)]
        for l in range(m + 2, lmax + 1):
            q[(l, m)] = ((2 * l - 1) * y * q[(l - 1, m)] - (l + m - 1) * q[(l - 2, m)]) / (l - m)
    cols = []
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            a = abs(m)
            n = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - a) / math.factorial(l + a))
            if m == 0:
                cols.append(n * q[(l, 0)])
            elif m > 0:
                cols.append(n * math.sqrt(2.0) * q[(l, a)] * cm[a])
            else:
                cols.append(n * math.sqrt(2.0) * q[(l, a)] * sm[a])
    return torch.stack(cols, dim=-1)


_SIZES = {"counts": None}


def set_structure_sizes(counts):
    """Atoms per structure of the batch the next forward works on (host ints), or None."""
    _SIZES["counts"] = None if counts is None else [int(c) for c in counts]


def _structure_counts(batch):
    """Atoms per structure: announced by the model wrapper (set_structure_sizes: no device read-back, capturable in a
    CUDA graph) or, as the reference does, read from `batch` (activation.py:1487 `bincount`)."""
    return list(_SIZES["counts"]) if _SIZES["counts"] is not None else torch.bincount(batch).tolist()


def _attend(q, k, values, counts, scale, dropout, bias_fn=None):
    """Per-structure multi-head attention shared by every variant: q, k [N, H, D]; values = list of [N, m, H, D] tensors
    that are mixed with the SAME attention weights; softmax over the atoms of the query's own structure (the reference
    masks cross-graph pairs of an [N_tot, N_tot] map or pads to [B, N_max]: identical weights).  bias_fn(g) -> [H, n, n]
    additive logit bias of structure g.  Returns the list of mixed values, [N, m, H*D] each."""
    H, D = q.shape[1], q.shape[2]
    if ops.pair_attention_available(H, D):
        # one scores / softmax / mix launch for the whole batch on the ragged pair layout (csrc/pair_attn.cu)
        blocks = None
        if bias_fn is not None:
            blocks = [bias_fn(g) for g in range(len(counts))]
            if any(b is None for b in blocks):
                assert all(b is None for b in blocks)
                blocks = None
        return ops.pair_attention(q, k, values, counts, scale, dropout, blocks)
    outs = [[] for _ in values]
    start = 0
    for g, n in enumerate(counts):
        sl = slice(start, start + n)
        attn = torch.einsum("ihd,jhd->hij", q[sl], k[sl]) * scale
        bias = bias_fn(g) if bias_fn is not None else None
        if bias is not None:
            attn = attn + bias
        attn = dropout(torch.nan_to_num(F.softmax(attn, dim=-1), nan=0.0))
        for o, v in zip(outs, values):
            o.append(torch.einsum("hij,jmhd->imhd", attn, v[sl]).reshape(n, v.shape[1], -1))
        start += n
    return [torch.cat(o, dim=0) for o in outs]


class _HTRGlobalAttention(nn.Module):
    """Shared body of the three HTR-scored variants (reference activation.py:1025-1216, :1217-1376, :1377-1567).

    The reference materialises score[N_tot, N_tot, C] over ALL atoms of the batch and masks cross-graph pairs.  Only the two
    marginals of that tensor are used (`.mean(dim=1)` -> queries, `.mean(dim=0)` -> keys), so this version works per
    structure and never forms it:
        q_in[i] = 1/N_tot  sum_l <X_i^l, sum_{j != i} Y_l(r_ij)> / (2l+1)  [+ rbf_proj(sum_{j != i} rbf(d_ij))]
        k_in[j] = 1/N_tot  sum_{i != j} sum_{l,m} X_i^{l,m} Y_l^m(r_ij) / (2l+1)  [+ the same distance term]
    Positions stay differentiable (to second order) through Y_l(r_ij) -- evaluated pole-free as Cartesian polynomials (an
    angle-based form gives NaN gradients at the poles) -- and through the distance term."""

    def _init_common(self, sphere_channels, lmax, num_heads, dropout):
        assert sphere_channels % num_heads == 0
        self.sphere_channels = sphere_channels
        self.lmax = lmax
        self.num_heads = num_heads
        self.head_dim = sphere_channels // num_heads
        self.scale = self.head_dim ** -0.5
        self.degree_sizes = [2 * l + 1 for l in range(lmax + 1)]

    def _init_projections(self, sphere_channels, lmax, dropout):
        self.q_proj = nn.Linear(sphere_channels, sphere_channels, bias=True)
        self.k_proj = nn.Linear(sphere_channels, sphere_channels, bias=True)
        self.v_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=(l == 0)) for l in range(lmax + 1)])
        self.out_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=False) for _ in range(lmax + 1)])
        self.norms = nn.ModuleList([nn.LayerNorm(sphere_channels) for _ in range(lmax + 1)])
        self.dropout = nn.Dropout(dropout)

    # hooks of the variants ---------------------------------------------------------------------------------
    def _distance_term(self, dist, valid):
        """[n, C] added to BOTH marginals (before the 1/N_tot), or None.  dist [n, n, 1] differentiable."""
        return None

    def _logit_bias(self, dist_detached):
        """[H, n, n] additive logit bias from the (detached) pair distances [n, n], or None."""
        return None

    def forward(self, x_emb, batch, pos):
        N, K, C = x_emb.shape
        H, D = self.num_heads, self.head_dim
        dev = x_emb.device
        key = (str(dev), x_emb.dtype)
        if getattr(self, "_wl_key", None) != key:          # built once per device: no host->device copy per call
            self._wl = torch.tensor([1.0 / (2 * l + 1) for l in range(self.lmax + 1) for _ in range(2 * l + 1)],
                                    dtype=x_emb.dtype, device=dev)
            self._wl_key = key
        wl = self._wl
        counts = _structure_counts(batch)
        q_parts, k_parts, geo = [], [], []
        start = 0
        for n in counts:                                     # structures are contiguous (`batch` is non-decreasing)
            sl = slice(start, start + n)
            p = pos[sl]
            diff = p.unsqueeze(1) - p.unsqueeze(0)           # diff[i, j] = r_i - r_j  (reference :1500)
            dist = diff.norm(dim=-1, keepdim=True).clamp(min=1e-8)
            rhat = F.normalize(diff / dist, dim=-1)
            valid = (~torch.eye(n, dtype=torch.bool, device=dev)).to(x_emb.dtype).unsqueeze(-1)
            Y = real_sh_integral(self.lmax, rhat) * wl * valid                       # [n, n, K]
            X = x_emb[sl]
            q_g = torch.einsum("ikc,ik->ic", X, Y.sum(dim=1))
            k_g = torch.matmul(Y.permute(1, 0, 2).reshape(n, n * K), X.reshape(n * K, C))
            extra = self._distance_term(dist, valid)
            if extra is not None:
                q_g, k_g = q_g + extra, k_g + extra
            q_parts.append(q_g / N)
            k_parts.append(k_g / N)
            geo.append(dist.detach().squeeze(-1))
            start += n
        q = _lin(self.q_proj, torch.cat(q_parts, dim=0)).view(N, H, D)
        k = _lin(self.k_proj, torch.cat(k_parts, dim=0)).view(N, H, D)
        values = [_lin(self.v_projs[l], x_emb[:, l * l:(l + 1) ** 2]).view(N, 2 * l + 1, H, D) for l in range(self.lmax + 1)]
        mixed = _attend(q, k, values, counts, self.scale, self.dropout, lambda g: self._logit_bias(geo[g]))
        res = []
        for l in range(self.lmax + 1):
            o = _lin(self.out_projs[l], mixed[l])
            feat = x_emb[:, l * l:(l + 1) ** 2]
            res.append(F.layer_norm(feat + o, (C,), self.norms[l].weight, self.norms[l].bias, self.norms[l].eps))
        return torch.cat(res, dim=1)


class GlobalNodeAttentionHTR(_HTRGlobalAttention):
    """reference activation.py:1025-1216."""

    def __init__(self, sphere_channels, lmax, num_heads=8, dropout=0.0):
        super().__init__()
        self._init_common(sphere_channels, lmax, num_heads, dropout)
        self._init_projections(sphere_channels, lmax, dropout)


class GlobalNodeAttentionHTR_with_distance(_HTRGlobalAttention):
    """reference activation.py:1217-1376: the pair score additionally carries rbf_proj(exp(-(d - mu)^2 / w)) -- linear in
    the Gaussian features, so its marginals are rbf_proj applied to the per-atom SUM of the features."""

    def __init__(self, sphere_channels, lmax, num_heads=8, dropout=0.0, num_rbf=16, rbf_cutoff=10.0):
        super().__init__()
        self._init_common(sphere_channels, lmax, num_heads, dropout)
        self.num_rbf = num_rbf
        self.rbf_cutoff = rbf_cutoff
        self.register_buffer("rbf_centers", torch.linspace(0.0, rbf_cutoff, num_rbf))
        self.rbf_width = (rbf_cutoff / num_rbf) ** 2
        self.rbf_proj = nn.Linear(num_rbf, sphere_channels, bias=False)
        self._init_projections(sphere_channels, lmax, dropout)

    def _distance_term(self, dist, valid):
        feats = torch.exp(-((dist - self.rbf_centers) ** 2) / self.rbf_width) * valid       # [n, n, num_rbf]
        return _lin(self.rbf_proj, feats.sum(dim=1))


class GlobalNodeAttentionHTR_with_ROPE(_HTRGlobalAttention):
    """reference activation.py:1377-1567 (BASELINE config 5; instantiated at
    equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_..._with_DISTANCE.py:231-237): HTR scores + a Fourier distance bias
    on the logits (positions detached there, :1534)."""

    def __init__(self, sphere_channels, lmax, num_heads=8, dropout=0.0, num_rbf=16, rbf_cutoff=10.0, use_rope=True,
                 rope_dim=16):
        super().__init__()
        self._init_common(sphere_channels, lmax, num_heads, dropout)
        self.use_rope = use_rope
        self.rope_dim = rope_dim
        # allocated but unused by this class in the reference too; kept for state_dict parity
        self.register_buffer("rbf_centers", torch.linspace(0.0, rbf_cutoff, num_rbf))
        self.rbf_width = (rbf_cutoff / num_rbf) ** 2
        self.rbf_proj = nn.Linear(num_rbf, sphere_channels, bias=False)
        self._init_projections(sphere_channels, lmax, dropout)
        if use_rope:
            self.rope_freqs = nn.Parameter(torch.randn(rope_dim) * 0.1)
            self.rope_proj = nn.Linear(rope_dim, num_heads, bias=False)

    def _logit_bias(self, dist_detached):
        if not self.use_rope:
            return None
        fourier = torch.cos(dist_detached.unsqueeze(-1) * self.rope_freqs.abs())
        return F.linear(fourier, self.rope_proj.weight).permute(2, 0, 1)


class GlobalNodeAttention(nn.Module):
    """reference activation.py:419-579: all-to-all attention on the invariant (l = 0) channel only, optional Euclidean-RoPE
    distance bias.  forward(x [N, C], batch, pos) -> [N, C].  Per structure instead of the reference's padded
    [B, N_max] tensors (same weights: padded keys are masked to -inf there)."""

    def __init__(self, d_model, num_heads=8, dropout=0.0, use_rope=True, rope_dim=16):
        super().__init__()
        assert d_model % num_heads == 0
        self.d_model = d_model
        self.num_heads = num_heads
        self.head_dim = d_model // num_heads
        self.scale = self.head_dim ** -0.5
        self.use_rope = use_rope
        self.rope_dim = rope_dim
        self.qkv_proj = nn.Linear(d_model, 3 * d_model, bias=False)
        self.out_proj = nn.Linear(d_model, d_model, bias=False)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        if use_rope:
            self.rope_freqs = nn.Parameter(torch.randn(rope_dim) * 0.1)
            self.rope_proj = nn.Linear(rope_dim, num_heads, bias=False)

    def forward(self, x, batch, pos):
        N, C = x.shape
        H, D = self.num_heads, self.head_dim
        counts = _structure_counts(batch)
        q, k, v = _lin(self.qkv_proj, x).chunk(3, dim=-1)
        bias_fn = None
        if self.use_rope:
            starts = [0]
            for c in counts:
                starts.append(starts[-1] + c)

            def bias_fn(g):       # the reference differentiates this bias w.r.t. pos (it detaches only afterwards, :566)
                p = pos[starts[g]:starts[g + 1]]
                d = (p.unsqueeze(1) - p.unsqueeze(0)).norm(dim=-1)
                return F.linear(torch.cos(d.unsqueeze(-1) * self.rope_freqs.abs()), self.rope_proj.weight).permute(2, 0, 1)
        (out,) = _attend(q.reshape(N, H, D), k.reshape(N, H, D), [v.reshape(N, 1, H, D)], counts, self.scale, self.dropout,
                         bias_fn)
        out = _lin(self.out_proj, out.reshape(N, C))
        return F.layer_norm(x + out, (C,), self.norm.weight, self.norm.bias, self.norm.eps)


class GlobalNodeAttentionFullEquivariant(nn.Module):
    """reference activation.py:686-920: one attention round PER DEGREE; queries / keys from the invariant norm of the
    degree-l features, values = the features themselves.  forward(x_emb [N, K, C], batch) -> [N, K, C]."""

    def __init__(self, sphere_channels, lmax, num_heads=8, dropout=0.0):
        super().__init__()
        assert sphere_channels % num_heads == 0
        self.sphere_channels = sphere_channels
        self.lmax = lmax
        self.num_heads = num_heads
        self.head_dim = sphere_channels // num_heads
        self.scale = self.head_dim ** -0.5
        self.degree_sizes = [2 * l + 1 for l in range(lmax + 1)]
        self.q_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=True) for _ in range(lmax + 1)])
        self.k_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=True) for _ in range(lmax + 1)])
        self.v_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=False) for _ in range(lmax + 1)])
        self.out_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=False) for _ in range(lmax + 1)])
        self.norms = nn.ModuleList([nn.LayerNorm(sphere_channels) for _ in range(lmax + 1)])
        self.dropout = nn.Dropout(dropout)

    def forward(self, x_emb, batch):
        N, K, C = x_emb.shape
        H, D = self.num_heads, self.head_dim
        counts = _structure_counts(batch)
        res = []
        for l in range(self.lmax + 1):
            feat = x_emb[:, l * l:(l + 1) ** 2]
            fnorm = feat.norm(dim=1)
            q = _lin(self.q_projs[l], fnorm).view(N, H, D)
            k = _lin(self.k_projs[l], fnorm).view(N, H, D)
            v = _lin(self.v_projs[l], feat).view(N, 2 * l + 1, H, D)
            (mixed,) = _attend(q, k, [v], counts, self.scale, self.dropout)
            o = _lin(self.out_projs[l], mixed)
            res.append(F.layer_norm(feat + o, (C,), self.norms[l].weight, self.norms[l].bias, self.norms[l].eps))
        return torch.cat(res, dim=1)
