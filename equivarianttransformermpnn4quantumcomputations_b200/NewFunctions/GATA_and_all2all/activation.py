"""Global node attention of BASELINE config 5 (reference NewFunctions/GATA_and_all2all/activation.py:1377-1567,
`GlobalNodeAttentionHTR_with_ROPE`, instantiated at equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_..._with_DISTANCE.py:231-237).

The reference materialises score[N_tot, N_tot, C] over ALL atoms of the batch and masks cross-graph pairs
(5.2 GB at 16 x 200 atoms).  Only the two marginals of that tensor are ever used (`.mean(dim=1)` -> queries,
`.mean(dim=0)` -> keys), so this version works per structure and never forms it:
    q_in[i] = 1/N_tot  sum_l <X_i^l, sum_{j != i} Y_l(r_ij)> / (2l+1)               (K-vector per atom)
    k_in[j] = 1/N_tot  sum_{i != j} sum_{l,m} X_i^{l,m} Y_l^m(r_ij) / (2l+1)        ([n, n K] x [n K, C] product)
Same parameters / state_dict keys; positions stay differentiable (to second order) through Y_l(r_ij), evaluated
pole-free as Cartesian polynomials (an angle-based form gives NaN gradients at the poles).
The projections run on the GEMM kernels; the per-structure n x n attention products are plain library matmuls.
Status: first functional version for parity; a fused flash-style kernel is the next step."""
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


class GlobalNodeAttentionHTR_with_ROPE(nn.Module):
    def __init__(self, sphere_channels, lmax, num_heads=8, dropout=0.0, num_rbf=16, rbf_cutoff=10.0, use_rope=True,
                 rope_dim=16):
        super().__init__()
        assert sphere_channels % num_heads == 0
        self.sphere_channels = sphere_channels
        self.lmax = lmax
        self.num_heads = num_heads
        self.head_dim = sphere_channels // num_heads
        self.scale = self.head_dim ** -0.5
        self.degree_sizes = [2 * l + 1 for l in range(lmax + 1)]
        self.use_rope = use_rope
        self.rope_dim = rope_dim
        # allocated but unused by this class in the reference too; kept for state_dict parity
        self.register_buffer("rbf_centers", torch.linspace(0.0, rbf_cutoff, num_rbf))
        self.rbf_width = (rbf_cutoff / num_rbf) ** 2
        self.rbf_proj = nn.Linear(num_rbf, sphere_channels, bias=False)
        self.q_proj = nn.Linear(sphere_channels, sphere_channels, bias=True)
        self.k_proj = nn.Linear(sphere_channels, sphere_channels, bias=True)
        self.v_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=(l == 0)) for l in range(lmax + 1)])
        self.out_projs = nn.ModuleList([nn.Linear(sphere_channels, sphere_channels, bias=False) for _ in range(lmax + 1)])
        self.norms = nn.ModuleList([nn.LayerNorm(sphere_channels) for _ in range(lmax + 1)])
        self.dropout = nn.Dropout(dropout)
        if use_rope:
            self.rope_freqs = nn.Parameter(torch.randn(rope_dim) * 0.1)
            self.rope_proj = nn.Linear(rope_dim, num_heads, bias=False)

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
        # structure sizes: announced by the model wrapper when it already has them on the host (set_structure_sizes:
        # keeps this module free of device read-backs, hence capturable in a CUDA graph), else as the reference (:1487)
        counts = list(_SIZES["counts"]) if _SIZES["counts"] is not None else torch.bincount(batch).tolist()
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        q_in = x_emb.new_zeros(N, C)
        k_in = x_emb.new_zeros(N, C)
        geo = []
        for g, n in enumerate(counts):                      # structures are contiguous (`batch` is non-decreasing)
            sl = slice(starts[g], starts[g + 1])
            p = pos[sl]
            diff = p.unsqueeze(1) - p.unsqueeze(0)           # diff[i, j] = r_i - r_j  (reference :1500)
            dist = diff.norm(dim=-1, keepdim=True).clamp(min=1e-8)
            rhat = F.normalize(diff / dist, dim=-1)
            valid = (~torch.eye(n, dtype=torch.bool, device=dev)).to(x_emb.dtype).unsqueeze(-1)
            Y = real_sh_integral(self.lmax, rhat) * wl * valid                       # [n, n, K]
            X = x_emb[sl]
            q_g = torch.einsum("ikc,ik->ic", X, Y.sum(dim=1)) / N
            k_g = torch.matmul(Y.permute(1, 0, 2).reshape(n, n * K), X.reshape(n * K, C)) / N
            q_in = q_in.index_add(0, torch.arange(sl.start, sl.stop, device=dev), q_g)
            k_in = k_in.index_add(0, torch.arange(sl.start, sl.stop, device=dev), k_g)
            geo.append(dist.detach().squeeze(-1))
        q = _lin(self.q_proj, q_in).view(N, H, D)
        k = _lin(self.k_proj, k_in).view(N, H, D)
        v = [_lin(self.v_projs[l], x_emb[:, l * l:(l + 1) ** 2]) for l in range(self.lmax + 1)]
        outs = [[] for _ in range(self.lmax + 1)]
        for g, n in enumerate(counts):
            sl = slice(starts[g], starts[g + 1])
            attn = torch.einsum("ihd,jhd->hij", q[sl], k[sl]) * self.scale
            if self.use_rope:
                fourier = torch.cos(geo[g].unsqueeze(-1) * self.rope_freqs.abs())    # positions detached (:1534)
                attn = attn + F.linear(fourier, self.rope_proj.weight).permute(2, 0, 1)
            attn = torch.nan_to_num(F.softmax(attn, dim=-1), nan=0.0)
            attn = self.dropout(attn)
            for l in range(self.lmax + 1):
                m = 2 * l + 1
                vl = v[l][sl].view(n, m, H, D)
                outs[l].append(torch.einsum("hij,jmhd->imhd", attn, vl).reshape(n, m, C))
        res = []
        for l in range(self.lmax + 1):
            o = _lin(self.out_projs[l], torch.cat(outs[l], dim=0))
            feat = x_emb[:, l * l:(l + 1) ** 2]
            res.append(F.layer_norm(feat + o, (C,), self.norms[l].weight, self.norms[l].bias, self.norms[l].eps))
        return torch.cat(res, dim=1)


def _not_built(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            raise NotImplementedError(f"{name}: not used by any BASELINE config (SURVEY §8f-4); "
                                      "only GlobalNodeAttentionHTR_with_ROPE is built")
    _Stub.__name__ = name
    return _Stub


GlobalNodeAttention = _not_built("GlobalNodeAttention")
GlobalNodeAttentionHTR = _not_built("GlobalNodeAttentionHTR")
GlobalNodeAttentionFullEquivariant = _not_built("GlobalNodeAttentionFullEquivariant")
GlobalNodeAttentionHTR_with_distance = _not_built("GlobalNodeAttentionHTR_with_distance")
