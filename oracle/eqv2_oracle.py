"""ORACLE -- CPU restatement of the reference's EquiformerV2 SO(2) graph-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs do.

It is a functional (parameter-dict driven) plain-PyTorch restatement of the algorithm in
/root/reference/models/EquiformerV2Functions/*.py and the model wrappers
equiformerv2_{qm9,oc20}.py; every function cites the reference file:line it follows.
Parameter names are the reference's own `state_dict` keys, so the same dict drives the
reference, this oracle and the CUDA product.

Pinning status: the reference has no tests or golden vectors.  This oracle is
pinned against outputs of the UNMODIFIED reference files imported in the build container
through oracle/refshim (third-party stand-ins) -- see oracle/make_golden.py and
tests/golden/*.pt, checked by tests/test_oracle_golden.py.  At the e3nn / torch_geometric /
fairchem boundaries (un-vendored, un-pinned upstream) parity is UNPINNED and replaced by the
mathematical invariants of tests/test_oracle_invariants.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import sh_basis

# --------------------------------------------------------------------------------------
# coefficient bookkeeping  (so3.py:45-199 CoefficientMappingModule, single resolution)
# --------------------------------------------------------------------------------------


class Layout:
    """Index tables for one (lmax, mmax) pair.

    l-major full index: l*l + l + m (so3.py:76-88).  Reduced set keeps |m| <= mmax
    (so3.py:155-170).  m-major order: [m=0: l=0..L][m=1 +: l=1..L][m=1 -: l=1..L]...
    (so3.py:97-109).
    """

    def __init__(self, lmax, mmax):
        self.lmax, self.mmax = lmax, mmax
        self.K = (lmax + 1) ** 2
        ls, ms = [], []
        for l in range(lmax + 1):
            mm = min(mmax, l)
            for m in range(-mm, mm + 1):
                ls.append(l)
                ms.append(m)
        self.l_red = torch.tensor(ls)
        self.m_red = torch.tensor(ms)
        self.Kr = len(ls)
        # position (in the full K vector) of each reduced coefficient
        self.mask = torch.tensor([l * l + l + m for l, m in zip(ls, ms)])
        perm = []
        self.m_size = []
        for m in range(mmax + 1):
            plus = [i for i in range(self.Kr) if ms[i] == m]
            perm += plus
            self.m_size.append(len(plus))
            if m > 0:
                perm += [i for i in range(self.Kr) if ms[i] == -m]
        self.to_m = torch.tensor(perm)           # m_primary[j] = l_primary[to_m[j]]
        inv = torch.empty(self.Kr, dtype=torch.long)
        inv[self.to_m] = torch.arange(self.Kr)
        self.to_l = inv
        # rotate_inv rescale per full row (so3.py:175-195): rows of degree l > mmax
        resc = torch.ones(self.K)
        for l in range(lmax + 1):
            if l > mmax:
                resc[l * l:(l + 1) ** 2] = math.sqrt((2 * l + 1) / (2 * mmax + 1))
        self.rescale = resc
        self.expand_l = torch.tensor([l for l in range(lmax + 1) for _ in range(2 * l + 1)])


# --------------------------------------------------------------------------------------
# Wigner-D  (wigner.py:17-39, so3.py:499-545)
# --------------------------------------------------------------------------------------
_JD_CACHE = {}


def jd(l, dtype=torch.float32):
    if l not in _JD_CACHE:
        _JD_CACHE[l] = torch.tensor(sh_basis.make_jd(l)[l], dtype=torch.float64)
    return _JD_CACHE[l].to(dtype)


def z_rot(angle, l):
    """wigner.py:31-39: M[i,i] = cos(f_i a), M[i,2l-i] = sin(f_i a), f = l..-l (diag written last)."""
    n = 2 * l + 1
    M = angle.new_zeros(angle.shape + (n, n))
    idx = torch.arange(n)
    freq = torch.arange(l, -l - 1, -1, dtype=angle.dtype)
    M[..., idx, n - 1 - idx] = torch.sin(freq * angle[..., None])
    M[..., idx, idx] = torch.cos(freq * angle[..., None])
    return M


def wigner_d(l, alpha, beta, gamma):
    J = jd(l, alpha.dtype)
    return z_rot(alpha, l) @ J @ z_rot(beta, l) @ J @ z_rot(gamma, l)


def _mat_y(a):
    c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def _mat_x(a):
    c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
    return torch.stack([torch.stack([o, z, z], -1), torch.stack([z, c, -s], -1), torch.stack([z, s, c], -1)], -2)


def edge_angles(rot):
    """so3.py:525-534: Euler angles (alpha, beta, gamma) of an edge frame [E,3,3]."""
    x = rot @ rot.new_tensor([0.0, 1.0, 0.0])
    x = F.normalize(x, p=2, dim=-1).clamp(-1, 1)
    beta = torch.acos(x[..., 1])
    alpha = torch.atan2(x[..., 0], x[..., 2])
    R = (_mat_y(alpha) @ _mat_x(beta)).transpose(-1, -2) @ rot
    gamma = torch.atan2(R[..., 0, 2], R[..., 0, 0])
    return alpha, beta, gamma


def rotation_to_wigner(rot, lmax):
    """so3.py:525-545: dense block-diagonal [E,K,K] (fp32), detached."""
    alpha, beta, gamma = edge_angles(rot)
    K = (lmax + 1) ** 2
    W = rot.new_zeros(len(rot), K, K)
    for l in range(lmax + 1):
        W[:, l * l:(l + 1) ** 2, l * l:(l + 1) ** 2] = wigner_d(l, alpha, beta, gamma)
    return W.detach()


def edge_rot_mat(edge_vec, rand_vec):
    """edge_rot_mat.py:13-80 with the `torch.rand_like(edge_vec) - 0.5` draw (line 28) passed
    in explicitly as `rand_vec` (parity needs the same draw on both paths)."""
    d = edge_vec.pow(2).sum(1).sqrt()
    nx = edge_vec / d.view(-1, 1)
    v = rand_vec / rand_vec.pow(2).sum(1).sqrt().view(-1, 1)
    vb = v.clone()
    vb[:, 0], vb[:, 1] = -v[:, 1], v[:, 0]
    vc = v.clone()
    vc[:, 1], vc[:, 2] = -v[:, 2], v[:, 1]
    dot_b = (vb * nx).sum(1).abs().view(-1, 1)
    dot_c = (vc * nx).sum(1).abs().view(-1, 1)
    dot = (v * nx).sum(1).abs().view(-1, 1)
    v = torch.where(dot > dot_b, vb, v)
    dot = (v * nx).sum(1).abs().view(-1, 1)
    v = torch.where(dot > dot_c, vc, v)
    assert (v * nx).sum(1).abs().max() < 0.99
    nz = torch.cross(nx, v, dim=1)
    nz = nz / nz.pow(2).sum(1, keepdim=True).sqrt()
    nz = nz / nz.pow(2).sum(1).sqrt().view(-1, 1)
    ny = torch.cross(nx, nz, dim=1)
    ny = ny / ny.pow(2).sum(1, keepdim=True).sqrt()
    inv = torch.stack([nz, nx, -ny], dim=2)          # columns (z, x_edge, -y)
    return inv.transpose(1, 2).detach()


# --------------------------------------------------------------------------------------
# S2 grid matrices  (so3.py:552-646 SO3_Grid)
# --------------------------------------------------------------------------------------
_GRID_CACHE = {}


def s2_grid_mats(lmax, mmax, res):
    """(to_grid_mat, from_grid_mat) [res, res, Kr] for SO3_Grid(lmax, mmax, resolution=res,
    normalization='component') -- so3.py:576-618."""
    key = (lmax, mmax, res)
    if key not in _GRID_CACHE:
        lay = Layout(lmax, mmax)
        shb, sha = sh_basis.to_s2grid_tensors(lmax, res, res)
        to_g = torch.einsum("mbi,am->bai", torch.tensor(shb, dtype=torch.float32), torch.tensor(sha, dtype=torch.float32))
        shb, sha = sh_basis.from_s2grid_tensors(res, res, lmax)
        fr_g = torch.einsum("am,mbi->bai", torch.tensor(sha, dtype=torch.float32), torch.tensor(shb, dtype=torch.float32))
        if lmax != mmax:
            for l in range(lmax + 1):
                if l > mmax:
                    f = math.sqrt((2 * l + 1) / (2 * mmax + 1))
                    to_g[:, :, l * l:(l + 1) ** 2] = to_g[:, :, l * l:(l + 1) ** 2] * f
                    fr_g[:, :, l * l:(l + 1) ** 2] = fr_g[:, :, l * l:(l + 1) ** 2] * f
        _GRID_CACHE[key] = (to_g[:, :, lay.mask].contiguous(), fr_g[:, :, lay.mask].contiguous())
    return _GRID_CACHE[key]


# --------------------------------------------------------------------------------------
# elementary layers
# --------------------------------------------------------------------------------------

def gaussian_smearing(dist, cutoff, num_basis=600, width=2.0, start=0.0):
    """equiformerv2_oc20.py:43-60."""
    offset = torch.linspace(start, cutoff, num_basis)
    coeff = -0.5 / (width * (offset[1] - offset[0])).item() ** 2
    return torch.exp(coeff * (dist.view(-1, 1) - offset.view(1, -1)).pow(2))


def radial_function(P, pre, x):
    """radial_function.py:5-30: Linear -> LN -> SiLU -> Linear -> LN -> SiLU -> Linear."""
    x = F.linear(x, P[pre + "net.0.weight"], P[pre + "net.0.bias"])
    x = F.silu(F.layer_norm(x, x.shape[-1:], P[pre + "net.1.weight"], P[pre + "net.1.bias"]))
    x = F.linear(x, P[pre + "net.3.weight"], P[pre + "net.3.bias"])
    x = F.silu(F.layer_norm(x, x.shape[-1:], P[pre + "net.4.weight"], P[pre + "net.4.bias"]))
    return F.linear(x, P[pre + "net.6.weight"], P[pre + "net.6.bias"])


def so3_linear(P, pre, x, lay_full):
    """so3.py:698-743 SO3_LinearV2: per-l weight, bias on l=0."""
    w = P[pre + "weight"][lay_full.expand_l]                 # [K, Cout, Cin]
    out = torch.einsum("bmi,moi->bmo", x, w)
    out = torch.cat([out[:, :1] + P[pre + "bias"].view(1, 1, -1), out[:, 1:]], dim=1)
    return out


def equivariant_norm(P, pre, x, norm_type, lmax, eps=1e-5):
    """layer_norm.py:38-108 ('layer_norm'), :112-201 ('layer_norm_sh'), :265-351 ('rms_norm_sh')."""
    N, K, C = x.shape
    if norm_type == "layer_norm":
        outs = []
        for l in range(lmax + 1):
            f = x[:, l * l:(l + 1) ** 2]
            if l == 0:
                f = f - f.mean(dim=2, keepdim=True)
            s = f.pow(2).mean(dim=1, keepdim=True).mean(dim=2, keepdim=True)
            f = f * ((s + eps).pow(-0.5) * P[pre + "affine_weight"][l].view(1, 1, -1))
            if l == 0:
                f = f + P[pre + "affine_bias"].view(1, 1, -1)
            outs.append(f)
        return torch.cat(outs, dim=1)
    if norm_type == "layer_norm_sh":
        outs = [F.layer_norm(x[:, :1], (C,), P[pre + "norm_l0.weight"], P[pre + "norm_l0.bias"], eps)]
        if lmax > 0:
            bw = torch.cat([torch.full((2 * l + 1,), 1.0 / (2 * l + 1)) for l in range(1, lmax + 1)]) / lmax
            f = x[:, 1:]
            s = torch.einsum("nic,i->nc", f.pow(2), bw).mean(dim=1).view(N, 1, 1)
            s = (s + eps).pow(-0.5)
            for l in range(1, lmax + 1):
                outs.append(x[:, l * l:(l + 1) ** 2] * (s * P[pre + "affine_weight"][l - 1].view(1, 1, -1)))
        return torch.cat(outs, dim=1)
    if norm_type == "rms_norm_sh":
        f0 = x[:, :1] - x[:, :1].mean(dim=2, keepdim=True)
        f = torch.cat([f0, x[:, 1:]], dim=1)
        bw = torch.cat([torch.full((2 * l + 1,), 1.0 / (2 * l + 1)) for l in range(lmax + 1)]) / (lmax + 1)
        s = torch.einsum("nic,i->nc", f.pow(2), bw).mean(dim=1).view(N, 1, 1)
        s = (s + eps).pow(-0.5)
        expand = torch.tensor([l for l in range(lmax + 1) for _ in range(2 * l + 1)])
        out = f * (s * P[pre + "affine_weight"][expand].view(1, K, C))
        out = torch.cat([out[:, :1] + P[pre + "affine_bias"].view(1, 1, C), out[:, 1:]], dim=1)
        return out
    raise ValueError(norm_type)


def so2_convolution(P, pre, x_l, x_edge, lay, c_in, c_out, extra=None, radial=True):
    """so2_ops.py:136-204 (+ SO2_m_Convolution :53-61).  x_l: [E,Kr,c_in] l-primary reduced."""
    E = x_l.shape[0]
    L, M = lay.lmax, lay.mmax
    x = x_l[:, lay.to_m]                                           # _m_primary (so3.py:322-334)
    rad = radial_function(P, pre + "rad_func.", x_edge) if radial else None
    outs = []
    n0 = (L + 1) * c_in
    x0 = x[:, :L + 1].reshape(E, -1)
    if rad is not None:
        x0 = x0 * rad[:, :n0]
    x0 = F.linear(x0, P[pre + "fc_m0.weight"], P[pre + "fc_m0.bias"])
    x_extra = None
    if extra is not None:
        x_extra, x0 = x0[:, :extra], x0[:, extra:]
    outs.append(x0.reshape(E, -1, c_out))
    off, off_rad = L + 1, n0
    for m in range(1, M + 1):
        nm = L - m + 1
        xm = x[:, off:off + 2 * nm].reshape(E, 2, -1)
        if rad is not None:
            xm = xm * rad[:, off_rad:off_rad + nm * c_in].unsqueeze(1)
        y = F.linear(xm, P[pre + f"so2_m_conv.{m - 1}.fc.weight"])
        half = y.shape[-1] // 2
        yr, yi = y[..., :half], y[..., half:]
        o_p = yr[:, 0:1] - yi[:, 1:2]
        o_m = yr[:, 1:2] + yi[:, 0:1]
        outs.append(torch.cat([o_p, o_m], dim=1).reshape(E, -1, c_out))
        off += 2 * nm
        off_rad += nm * c_in
    out = torch.cat(outs, dim=1)[:, lay.to_l]                       # _l_primary
    return (out, x_extra) if extra is not None else out


def s2_activation(x, to_g, fr_g):
    """activation.py:153-170."""
    g = torch.einsum("bai,zic->zbac", to_g, x)
    return torch.einsum("bai,zbac->zic", fr_g, F.silu(g))


def separable_s2_activation(scalars, x, to_g, fr_g):
    """activation.py:173-192."""
    out = s2_activation(x, to_g, fr_g)
    return torch.cat([F.silu(scalars).unsqueeze(1), out[:, 1:]], dim=1)


def smooth_leaky_relu(x, a=0.2):
    """activation.py:66-75."""
    return ((1 + a) / 2) * x + ((1 - a) / 2) * x * (2 * torch.sigmoid(x) - 1)


def segment_softmax(src, index, n):
    """torch_geometric.utils.softmax as used at transformer_block.py:315."""
    idx = index.view(-1, 1).expand_as(src)
    smax = torch.full((n, src.shape[1]), float("-inf"), dtype=src.dtype)
    smax = smax.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    e = (src - smax.gather(0, idx)).exp()
    ssum = torch.zeros(n, src.shape[1], dtype=src.dtype).scatter_add(0, idx, e)
    return e / (ssum.gather(0, idx) + 1e-16)


class Hyper:
    """Hyper-parameters of one model family (SURVEY §8 size table)."""

    def __init__(self, lmax, mmax, C, H, heads, alpha_ch, value_ch, ffn_hidden, edge_ch=128,
                 num_layers=12, norm_type="rms_norm_sh", grid_res=18, num_rbf=600, cutoff=12.0,
                 max_elements=90, avg_degree=23.395238876342773, avg_nodes=77.81317,
                 num_targets=1, max_neighbors=20):
        self.__dict__.update(locals())
        del self.__dict__["self"]
        self.lay = Layout(lmax, mmax)
        self.lay_full = Layout(lmax, lmax)


def graph_attention(P, pre, hp, x, Z, rbf, edge_index, W, out_channels=None):
    """transformer_block.py:231-336 SO2EquivariantGraphAttention.forward (sep-S2 path)."""
    lay, C, H = hp.lay, hp.C, hp.H
    src, dst = edge_index[0], edge_index[1]
    x_edge = torch.cat([rbf, P[pre + "source_embedding.weight"][Z[src]], P[pre + "target_embedding.weight"][Z[dst]]], dim=1)
    msg = torch.cat([x[src], x[dst]], dim=2)                        # :250-264
    msg = torch.bmm(W[:, lay.mask, :], msg)                         # so3.py:509-512
    n_alpha = hp.heads * hp.alpha_ch
    msg, extra = so2_convolution(P, pre + "so2_conv_1.", msg, x_edge, lay, 2 * C, H, extra=n_alpha + H)
    x_alpha, gate = extra[:, :n_alpha], extra[:, n_alpha:]
    to_g, fr_g = s2_grid_mats(hp.lmax, hp.mmax, hp.grid_res)
    msg = separable_s2_activation(gate, msg, to_g, fr_g)            # :292-294
    msg = so2_convolution(P, pre + "so2_conv_2.", msg, None, lay, H, hp.heads * hp.value_ch, radial=False)
    a = x_alpha.reshape(-1, hp.heads, hp.alpha_ch)                  # :311-315
    a = F.layer_norm(a, (hp.alpha_ch,), P[pre + "alpha_norm.weight"], P[pre + "alpha_norm.bias"])
    a = smooth_leaky_relu(a)
    alpha = torch.einsum("bik,ik->bi", a, P[pre + "alpha_dot"])
    alpha = segment_softmax(alpha, dst, x.shape[0])
    E, Kr = msg.shape[:2]
    msg = (msg.reshape(E, Kr, hp.heads, hp.value_ch) * alpha.view(E, 1, hp.heads, 1)).reshape(E, Kr, -1)
    Winv = W.transpose(1, 2)[:, :, lay.mask] * lay.rescale.view(1, -1, 1)   # so3.py:516-521
    msg = torch.bmm(Winv, msg)
    out = torch.zeros(x.shape[0], lay.K, msg.shape[2], dtype=msg.dtype).index_add_(0, dst, msg)
    return so3_linear(P, pre + "proj.", out, hp.lay_full)


def feed_forward(P, pre, hp, x):
    """transformer_block.py:417-453 (sep-S2 branch)."""
    gate = F.linear(x[:, 0], P[pre + "gating_linear.weight"], P[pre + "gating_linear.bias"])
    h = so3_linear(P, pre + "so3_linear_1.", x, hp.lay_full)
    to_g, fr_g = s2_grid_mats(hp.lmax, hp.lmax, hp.grid_res)
    h = separable_s2_activation(gate, h, to_g, fr_g)
    return so3_linear(P, pre + "so3_linear_2.", h, hp.lay_full)


def trans_block(P, pre, hp, x, Z, rbf, edge_index, W):
    """transformer_block.py:585-634 (drop rates 0 / eval)."""
    h = equivariant_norm(P, pre + "norm_1.", x, hp.norm_type, hp.lmax)
    x = x + graph_attention(P, pre + "ga.", hp, h, Z, rbf, edge_index, W)
    h = equivariant_norm(P, pre + "norm_2.", x, hp.norm_type, hp.lmax)
    return x + feed_forward(P, pre + "ffn.", hp, h)


def edge_degree_embedding(P, pre, hp, Z, rbf, edge_index, W, num_nodes, rescale):
    """input_block.py:86-131."""
    lay = hp.lay
    src, dst = edge_index[0], edge_index[1]
    x_edge = torch.cat([rbf, P[pre + "source_embedding.weight"][Z[src]], P[pre + "target_embedding.weight"][Z[dst]]], dim=1)
    m0 = radial_function(P, pre + "rad_func.", x_edge).reshape(-1, hp.lmax + 1, hp.C)
    pad = m0.new_zeros(m0.shape[0], lay.Kr - (hp.lmax + 1), hp.C)
    x = torch.cat([m0, pad], dim=1)[:, lay.to_l]
    Winv = W.transpose(1, 2)[:, :, lay.mask] * lay.rescale.view(1, -1, 1)
    x = torch.bmm(Winv, x)
    out = torch.zeros(num_nodes, lay.K, hp.C, dtype=x.dtype).index_add_(0, dst, x)
    return out / rescale


def backbone(P, hp, Z, edge_index, edge_dist, edge_vec, rand_vec):
    """Shared trunk of equiformerv2_oc20.py:236-275 / equiformerv2_qm9.py:571-637."""
    R = edge_rot_mat(edge_vec.detach(), rand_vec)
    W = rotation_to_wigner(R, hp.lmax)
    N = Z.shape[0]
    x = torch.zeros(N, hp.lay.K, hp.C)
    x = torch.cat([P["sphere_embedding.weight"][Z].unsqueeze(1), x[:, 1:]], dim=1)
    rbf = gaussian_smearing(edge_dist, hp.cutoff, hp.num_rbf, 2.0)
    x = x + edge_degree_embedding(P, "edge_degree_embedding.", hp, Z, rbf, edge_index, W, N, hp.avg_degree)
    for i in range(hp.num_layers):
        x = trans_block(P, f"blocks.{i}.", hp, x, Z, rbf, edge_index, W)
    x = equivariant_norm(P, "norm.", x, hp.norm_type, hp.lmax)
    return x, rbf, W


def oc20_forward(P, hp, Z, batch, num_graphs, edge_index, edge_dist, edge_vec, rand_vec):
    """equiformerv2_oc20.py:205-289: energy + direct forces."""
    x, rbf, W = backbone(P, hp, Z, edge_index, edge_dist, edge_vec, rand_vec)
    node_e = feed_forward(P, "energy_block.", hp, x)[:, 0, 0]
    energy = torch.zeros(num_graphs, dtype=node_e.dtype).index_add_(0, batch, node_e) / hp.avg_nodes
    f = graph_attention(P, "force_block.", hp, x, Z, rbf, edge_index, W)
    return energy, f[:, 1:4, 0]


def qm9_forward(P, hp, Z, batch, num_graphs, edge_index, edge_dist, edge_vec, rand_vec):
    """equiformerv2_qm9.py:526-696: num_targets FFN heads, per-graph sums."""
    x, _, _ = backbone(P, hp, Z, edge_index, edge_dist, edge_vec, rand_vec)
    preds = []
    for t in range(hp.num_targets):
        node = feed_forward(P, f"output_blocks.{t}.", hp, x)[:, 0, 0]
        preds.append(torch.zeros(num_graphs, dtype=node.dtype).index_add_(0, batch, node))
    return torch.stack(preds, dim=1)


# --------------------------------------------------------------------------------------
# neighbour lists
# --------------------------------------------------------------------------------------

def radius_graph_qm9(pos, batch, cutoff, max_neighbors):
    """equiformerv2_qm9.py:423-525: per-molecule dense search, strict 0<d<cutoff, per-dst
    nearest `max_neighbors`; edge order = row-major (src, dst) within each molecule."""
    ei, dd, vv = [], [], []
    for g in torch.unique(batch):
        nodes = torch.where(batch == g)[0]
        p = pos[nodes]
        diff = p.unsqueeze(0) - p.unsqueeze(1)          # diff[i, j] = p[j] - p[i]
        dist = torch.norm(diff, dim=2)
        src, dst = torch.where((dist < cutoff) & (dist > 0))
        if max_neighbors is not None and len(src) > 0:
            keep = torch.zeros(len(src), dtype=torch.bool)
            for d in torch.unique(dst):
                ids = torch.where(dst == d)[0]
                if len(ids) > max_neighbors:
                    order = torch.sort(dist[src[ids], d])[1]
                    keep[ids[order[:max_neighbors]]] = True
                else:
                    keep[ids] = True
            src, dst = src[keep], dst[keep]
        ei.append(torch.stack([nodes[src], nodes[dst]]))
        dd.append(dist[src, dst])
        vv.append(diff[src, dst])
    return torch.cat(ei, dim=1), torch.cat(dd), torch.cat(vv)


def radius_graph_pbc_fairchem(pos, cell, batch, natoms, cutoff, max_neighbors, strict=False):
    """Brute-force stand-in for fairchem radius_graph_pbc + get_pbc_distances as called at
    equiformerv2_oc20.py:223-234 (fairchem-core is un-vendored and un-pinned by the reference; PARITY UNPINNED,
    SURVEY §8f-1).  Semantics restated from fairchem-core's published `radius_graph_pbc` / `get_max_neighbors_mask`
    (graph/compute.py, 1.x / 2.x series): candidates are all periodic images with SQUARED distance d2 <= r_c^2 and
    d2 > 1e-4; a centre with more than `max_neighbors` candidates keeps, when not strict, those with
    d2 <= d2_sorted[max_neighbors] + 0.01 (the (max_neighbors+1)-th smallest squared distance plus the degeneracy
    tolerance -- so a truncated centre keeps AT LEAST max_neighbors + 1 edges, which is what makes the reference's
    _AVG_DEGREE 23.4 at max_neighbors 20), and exactly the max_neighbors nearest when strict.
    edge_index[0] = neighbour j, edge_index[1] = centre i, vec = pos[j] + offset - pos[i]."""
    ei, dd, vv = [], [], []
    start = 0
    for g, n in enumerate(natoms.tolist()):
        p = pos[start:start + n].double()
        c = cell[g].double()
        # number of repeats needed along each lattice vector
        vol = torch.det(c).abs()
        reps = []
        for k in range(3):
            a, b = c[(k + 1) % 3], c[(k + 2) % 3]
            h = vol / torch.linalg.norm(torch.linalg.cross(a, b))
            reps.append(int(math.ceil(cutoff / h.item())))
        rng = [torch.arange(-r, r + 1, dtype=torch.float64) for r in reps]
        cells = torch.stack(torch.meshgrid(*rng, indexing="ij"), dim=-1).reshape(-1, 3)
        offs = cells @ c                                                       # [S,3]
        # vec[i, j, s] = p[j] + off[s] - p[i]
        vec = p.unsqueeze(0).unsqueeze(2) + offs.view(1, 1, -1, 3) - p.view(n, 1, 1, 3)
        d2 = (vec * vec).sum(-1)
        d = d2.sqrt()
        ok = (d2 <= cutoff * cutoff) & (d2 > 1e-4)
        ci, cj, cs = torch.where(ok)
        dsel = d2[ci, cj, cs]
        keep = torch.ones(len(ci), dtype=torch.bool)
        for i in range(n):
            ids = torch.where(ci == i)[0]
            if len(ids) > max_neighbors:
                ds, order = torch.sort(dsel[ids])
                thr = ds[max_neighbors] + 0.01       # non-strict: index max_neighbors of the sorted squared distances
                drop = ids[dsel[ids] > thr] if not strict else ids[order[max_neighbors:]]
                keep[drop] = False
        ci, cj, cs = ci[keep], cj[keep], cs[keep]
        ei.append(torch.stack([cj + start, ci + start]))
        dd.append(d[ci, cj, cs].to(pos.dtype))
        vv.append(vec[ci, cj, cs].to(pos.dtype))
        start += n
    return torch.cat(ei, dim=1), torch.cat(dd), torch.cat(vv)


def canonical_edge_order(edge_index, num_nodes, tiebreak=None):
    """Permutation sorting edges by (dst, src[, tiebreak]) -- the 'canonical sort' under which
    neighbour indices are compared bit-exactly."""
    key = edge_index[1].double() * num_nodes + edge_index[0].double()
    if tiebreak is not None:
        order = np.lexsort((tiebreak.numpy(), key.numpy()))
        return torch.from_numpy(order)
    return torch.argsort(key, stable=True)


def _pbc27_candidates(pos_b, cell_b, cutoff):
    """27-image candidate search shared by the MatPES builders (equiformerv2_MatPES.py:270-290,
    equiformerv2_MatPESv2.py:186-201): image order = meshgrid(-1..1, indexing='ij'); for image offset o,
    diff[i, j] = (pos[j] + o) - pos[i]; keep dist < cutoff (and > 1e-6 for the zero image).
    Returns src (=i), dst (=j), image index, and the true image vectors."""
    rng = torch.arange(-1, 2, dtype=torch.float32)
    gx, gy, gz = torch.meshgrid(rng, rng, rng, indexing="ij")
    offs = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1) @ cell_b
    src_l, dst_l, img_l, vec_l = [], [], [], []
    for s, o in enumerate(offs):
        diff = (pos_b + o.unsqueeze(0)).unsqueeze(0) - pos_b.unsqueeze(1)
        dist = torch.norm(diff, dim=2)
        within = (dist < cutoff) & (dist > 1e-6) if o.abs().sum() < 1e-6 else (dist < cutoff)
        src, dst = torch.where(within)
        src_l.append(src)
        dst_l.append(dst)
        img_l.append(torch.full_like(src, s))
        vec_l.append(diff[src, dst])
    return torch.cat(src_l), torch.cat(dst_l), torch.cat(img_l), torch.cat(vec_l)


def radius_graph_matpes(pos, cell, batch, cutoff, max_neighbors, version):
    """version 1: equiformerv2_MatPES.py:258-340 -- vectors keep the periodic image offset, per-destination
    ranking by the true image distance.
    version 2: equiformerv2_MatPESv2.py:177-240 (also equiformerv2_MatPES_GATAV2.py:285-349) -- candidates as
    in v1, but the ranking distance and the returned vector are pos[dst] - pos[src] WITHOUT the image offset
    (duplicates and zero-length self-image edges are kept).
    Returns edge_index [2,E] (row 0 = src, row 1 = dst), distance, vector, image index."""
    ei, dd, vv, im = [], [], [], []
    for g in range(cell.shape[0]):
        nodes = torch.where(batch == g)[0]
        pb = pos[nodes]
        src, dst, img, vec = _pbc27_candidates(pb, cell[g], cutoff)
        if len(src) == 0:
            continue
        if version == 2:
            vec = pb[dst] - pb[src]
        dist = torch.norm(vec, dim=1)
        if max_neighbors is not None:
            keep = torch.zeros(len(src), dtype=torch.bool)
            for d in torch.unique(dst):
                ids = torch.where(dst == d)[0]
                if len(ids) > max_neighbors:
                    order = torch.sort(dist[ids], stable=True)[1]
                    keep[ids[order[:max_neighbors]]] = True
                else:
                    keep[ids] = True
            src, dst, img, vec, dist = src[keep], dst[keep], img[keep], vec[keep], dist[keep]
        ei.append(torch.stack([nodes[src], nodes[dst]]))
        dd.append(dist)
        vv.append(vec)
        im.append(img)
    return torch.cat(ei, dim=1), torch.cat(dd), torch.cat(vv), torch.cat(im)


def edge_rot_mat_deterministic(edge_vec):
    """equiformerv2_MatPESv2.py:41-66: helper axis = cardinal axis of the smallest |component|."""
    ev = edge_vec.detach()
    dist = torch.sqrt(torch.sum(ev ** 2, dim=1, keepdim=True)).clamp(min=1e-8)
    nx = ev / dist
    ref = torch.eye(3, dtype=ev.dtype)[torch.argmin(torch.abs(nx), dim=1)]
    nz = torch.cross(nx, ref, dim=1)
    nz = nz / torch.sqrt(torch.sum(nz ** 2, dim=1, keepdim=True)).clamp(min=1e-8)
    ny = torch.cross(nx, nz, dim=1)
    ny = ny / torch.sqrt(torch.sum(ny ** 2, dim=1, keepdim=True)).clamp(min=1e-8)
    inv = torch.stack([nz, nx, -ny], dim=2)
    return inv.transpose(1, 2)


def matpes_v2_forward(P, hp, Z, batch, natoms, pos, edge_index):
    """equiformerv2_MatPESv2.py:242-300: energy only; differentiable w.r.t. `pos` through
    dvec = pos[dst] - pos[src] -> distance -> RBF -> radial MLPs (frames / Wigner-D detached).
    Returns energy_total [B]."""
    dvec = pos[edge_index[1]] - pos[edge_index[0]]
    dist = torch.norm(dvec, dim=1)
    R = edge_rot_mat_deterministic(dvec)
    W = rotation_to_wigner(R, hp.lmax)
    N = Z.shape[0]
    x = torch.zeros(N, hp.lay.K, hp.C)
    x = torch.cat([P["sphere_embedding.weight"][Z].unsqueeze(1), x[:, 1:]], dim=1)
    rbf = gaussian_smearing(dist, hp.cutoff, hp.num_rbf, 2.0)
    x = x + edge_degree_embedding(P, "edge_degree_embedding.", hp, Z, rbf, edge_index, W, N, hp.avg_degree)
    for i in range(hp.num_layers):
        x = trans_block(P, f"blocks.{i}.", hp, x, Z, rbf, edge_index, W)
    x = equivariant_norm(P, "norm.", x, hp.norm_type, hp.lmax)
    node_e = feed_forward(P, "energy_block.", hp, x)[:, 0, 0]
    return torch.zeros(len(natoms), dtype=node_e.dtype).index_add_(0, batch, node_e)
