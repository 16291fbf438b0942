"""ORACLE shim for the `e3nn.o3` symbols the reference touches.

Call sites in the reference: so3.py:527-533 (xyz_to_angles / angles_to_matrix),
so3.py:584-608 (ToS2Grid / FromS2Grid, attributes .shb [m,b,i] and .sha [a,m]),
drop.py:13,79 (import only), equiformerv2_MatPES_GATAV2.py:137-140 (SphericalHarmonics).
"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import sh_basis  # noqa: E402


def xyz_to_angles(xyz):
    xyz = torch.nn.functional.normalize(xyz, p=2, dim=-1)
    xyz = xyz.clamp(-1, 1)
    beta = torch.acos(xyz[..., 1])
    alpha = torch.atan2(xyz[..., 0], xyz[..., 2])
    return alpha, beta


def matrix_x(angle):
    c, s = angle.cos(), angle.sin()
    o, z = torch.ones_like(angle), torch.zeros_like(angle)
    return torch.stack([torch.stack([o, z, z], -1), torch.stack([z, c, -s], -1), torch.stack([z, s, c], -1)], -2)


def matrix_y(angle):
    c, s = angle.cos(), angle.sin()
    o, z = torch.ones_like(angle), torch.zeros_like(angle)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def angles_to_matrix(alpha, beta, gamma):
    alpha, beta, gamma = torch.broadcast_tensors(alpha, beta, gamma)
    return matrix_y(alpha) @ matrix_x(beta) @ matrix_y(gamma)


class ToS2Grid(torch.nn.Module):
    def __init__(self, lmax=None, res=None, normalization="component", dtype=None, device=None):
        super().__init__()
        res_beta, res_alpha = res
        shb, sha = sh_basis.to_s2grid_tensors(lmax, res_beta, res_alpha, normalization)
        self.lmax, self.res_beta, self.res_alpha = lmax, res_beta, res_alpha
        self.register_buffer("shb", torch.tensor(shb, dtype=torch.get_default_dtype()))
        self.register_buffer("sha", torch.tensor(sha, dtype=torch.get_default_dtype()))


class FromS2Grid(torch.nn.Module):
    def __init__(self, res=None, lmax=None, normalization="component", lmax_in=None, dtype=None, device=None):
        super().__init__()
        res_beta, res_alpha = res
        shb, sha = sh_basis.from_s2grid_tensors(res_beta, res_alpha, lmax, normalization)
        self.lmax, self.res_beta, self.res_alpha = lmax, res_beta, res_alpha
        self.register_buffer("shb", torch.tensor(shb, dtype=torch.get_default_dtype()))
        self.register_buffer("sha", torch.tensor(sha, dtype=torch.get_default_dtype()))


def _cart_sh(lmax, xyz):
    """Pole-free Cartesian evaluation of the App. B.1 basis on unit vectors (differentiable)."""
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    # A_m = Re (z + i x)^m = s^m cos(m a), B_m = Im = s^m sin(m a)   (alpha = atan2(x, z))
    A = [torch.ones_like(x)]
    B = [torch.zeros_like(x)]
    for m in range(1, lmax + 1):
        A.append(A[-1] * z - B[-1] * x)
        B.append(B[m - 1] * z + A[m - 1] * x)
    # Q_l^m(y) = P_l^m(y) / s^m (no CS phase), polynomial in y
    Q = {}
    for m in range(lmax + 1):
        dfact = 1.0
        for k in range(1, 2 * m, 2):
            dfact *= k
        Q[(m, m)] = torch.full_like(y, dfact)
        if m + 1 <= lmax:
            Q[(m + 1, m)] = (2 * m + 1) * y * Q[(m, m)]
        for l in range(m + 2, lmax + 1):
            Q[(l, m)] = ((2 * l - 1) * y * Q[(l - 1, m)] - (l + m - 1) * Q[(l - 2, m)]) / (l - m)
    cols = []
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            n = sh_basis.sh_norm(l, m)
            if m < 0:
                cols.append(n * math.sqrt(2.0) * Q[(l, -m)] * B[-m])
            elif m == 0:
                cols.append(n * Q[(l, 0)])
            else:
                cols.append(n * math.sqrt(2.0) * Q[(l, m)] * A[m])
    return torch.stack(cols, dim=-1)


def spherical_harmonics(l, x, normalize, normalization="integral"):
    ls = [l] if isinstance(l, int) else list(l)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)
    full = _cart_sh(max(ls), x)
    outs = []
    for li in ls:
        blk = full[..., li * li:(li + 1) * (li + 1)]
        if normalization == "norm":
            blk = blk * math.sqrt(4 * math.pi / (2 * li + 1))
        elif normalization == "component":
            blk = blk * math.sqrt(4 * math.pi)
        outs.append(blk)
    return torch.cat(outs, dim=-1)


class _Irrep:
    def __init__(self, l, p=1):
        self.l, self.p, self.dim = l, p, 2 * l + 1

    def is_scalar(self):
        return self.l == 0


class Irreps(list):
    def __init__(self, spec=None):
        super().__init__()
        if isinstance(spec, str):
            for tok in spec.split("+"):
                tok = tok.strip()
                mul, ir = tok.split("x") if "x" in tok else ("1", tok)
                self.append((int(mul), _Irrep(int(ir[:-1]))))
        elif spec is not None:
            for mul, ir in spec:
                self.append((mul, ir))

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, _Irrep(l)) for l in range(lmax + 1)])

    @property
    def num_irreps(self):
        return sum(m for m, _ in self)

    @property
    def dim(self):
        return sum(m * ir.dim for m, ir in self)

    @property
    def ls(self):
        return [ir.l for m, ir in self for _ in range(m)]


class SphericalHarmonics(torch.nn.Module):
    def __init__(self, irreps_out, normalize, normalization="integral", irreps_in=None):
        super().__init__()
        self.ls = Irreps(irreps_out).ls if not isinstance(irreps_out, Irreps) else irreps_out.ls
        self.normalize, self.normalization = normalize, normalization

    def forward(self, x):
        return spherical_harmonics(self.ls, x, self.normalize, self.normalization)


class ElementwiseTensorProduct(torch.nn.Module):  # drop.py:79 (dead code upstream)
    def __init__(self, *a, **k):
        super().__init__()
