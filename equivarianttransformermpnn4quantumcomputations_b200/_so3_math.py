"""Host-side constants of the SO(3) machinery, computed once in fp64 (numpy) at module build
time: the J matrices of the Z-J-Z-J-Z Wigner-D factorisation (reference loads them from the
git-ignored `Jd.pt`, wigner.py:9) and the S2 grid projection matrices (reference takes them
from e3nn ToS2Grid/FromS2Grid, so3.py:584-608).

Real spherical harmonics Y_{l,m}: polar axis y, azimuth alpha = atan2(x, z), no Condon-Shortley
phase, 'integral' normalisation.  Everything here is evaluated with
recurrences / exact quadrature -- it shares no code with oracle/sh_basis.py, so the two act as
a cross-check of each other (tests/test_host_constants.py).
"""
import functools
import math

import numpy as np


def _legendre_no_cs(lmax, c, s):
    """P_l^m(c) * (no CS phase), m >= 0, via upward recurrences. Returns dict[(l,m)]."""
    P = {}
    for m in range(lmax + 1):
        pmm = np.ones_like(c)
        for k in range(1, m + 1):
            pmm = pmm * (2 * k - 1) * s
        P[(m, m)] = pmm
        if m + 1 <= lmax:
            P[(m + 1, m)] = (2 * m + 1) * c * pmm
        for l in range(m + 2, lmax + 1):
            P[(l, m)] = ((2 * l - 1) * c * P[(l - 1, m)] - (l + m - 1) * P[(l - 2, m)]) / (l - m)
    return P


def _norm(l, m):
    return math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))


def real_sh(lmax, cos_b, sin_b, alpha):
    """[..., (lmax+1)^2] in l-major order, m = -l..l."""
    P = _legendre_no_cs(lmax, cos_b, sin_b)
    cols = []
    r2 = math.sqrt(2.0)
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            base = _norm(l, abs(m)) * P[(l, abs(m))]
            if m < 0:
                cols.append(base * r2 * np.sin(-m * alpha))
            elif m == 0:
                cols.append(base)
            else:
                cols.append(base * r2 * np.cos(m * alpha))
    return np.stack(cols, axis=-1)


def _sh_of_xyz(lmax, xyz):
    y = np.clip(xyz[..., 1], -1.0, 1.0)
    s = np.sqrt(np.maximum(0.0, 1.0 - y * y))
    alpha = np.arctan2(xyz[..., 0], xyz[..., 2])
    return real_sh(lmax, y, s, alpha)


@functools.lru_cache(maxsize=None)
def jd_blocks(lmax):
    """J_l = D_l(S), S = [[0,1,0],[1,0,0],[0,0,-1]], from the exact projection
    D_ij = integral Y_i(S p) Y_j(p) dOmega (Gauss-Legendre x uniform azimuth quadrature)."""
    n = 2 * lmax + 4
    xs, ws = np.polynomial.legendre.leggauss(n)
    na = 2 * n
    al = np.arange(na) * (2 * math.pi / na)
    C, A = np.meshgrid(xs, al, indexing="ij")
    Wq = np.repeat(ws[:, None], na, axis=1) * (2 * math.pi / na)
    S_ = np.sqrt(1 - C * C)
    p = np.stack([S_ * np.sin(A), C, S_ * np.cos(A)], axis=-1)          # (x, y, z)
    Sp = np.stack([p[..., 1], p[..., 0], -p[..., 2]], axis=-1)
    Yp = _sh_of_xyz(lmax, p)
    Ys = _sh_of_xyz(lmax, Sp)
    out = []
    for l in range(lmax + 1):
        sl = slice(l * l, (l + 1) ** 2)
        J = np.einsum("bai,baj,ba->ij", Ys[..., sl], Yp[..., sl], Wq)
        J[np.abs(J) < 1e-13] = 0.0
        out.append(J)
    return out


def jd_packed(lmax):
    return np.concatenate([J.reshape(-1) for J in jd_blocks(lmax)]).astype(np.float32)


def _dh_weights(b):
    k = np.arange(b)
    w = np.array([
        (2.0 / b) * math.sin(math.pi * (2 * j + 1) / (4.0 * b))
        * np.sum(np.sin((2 * j + 1) * (2 * k + 1) * math.pi / (4.0 * b)) / (2 * k + 1))
        for j in range(2 * b)])
    return w / (2.0 * (2 * b) ** 2)


@functools.lru_cache(maxsize=None)
def s2_grid_matrices(lmax, mmax, res_beta, res_alpha):
    """(to_grid, from_grid), each [res_beta, res_alpha, Kr] fp32, reduced to |m| <= mmax with the
    sqrt((2l+1)/(2 mmax+1)) rescale for l > mmax -- the buffers `to_grid_mat` / `from_grid_mat`
    of SO3_Grid (so3.py:584-622), normalization='component'."""
    betas = (np.arange(res_beta) + 0.5) / res_beta * math.pi
    alphas = np.arange(res_alpha) / res_alpha * 2 * math.pi
    Bm, Am = np.meshgrid(betas, alphas, indexing="ij")
    Y = real_sh(lmax, np.cos(Bm), np.sin(Bm), Am)                        # [b, a, K]
    assert res_beta % 2 == 0
    qw = _dh_weights(res_beta // 2) * res_beta ** 2 / res_alpha           # [b]
    to_g = np.empty_like(Y)
    fr_g = np.empty_like(Y)
    for l in range(lmax + 1):
        sl = slice(l * l, (l + 1) ** 2)
        nt = math.sqrt(4 * math.pi) / math.sqrt(2 * l + 1) / math.sqrt(lmax + 1)
        nf = math.sqrt(4 * math.pi) * math.sqrt(2 * l + 1) * math.sqrt(lmax + 1)
        to_g[..., sl] = Y[..., sl] * nt
        fr_g[..., sl] = Y[..., sl] * nf * qw[:, None, None]
    # the reference builds the matrices in fp32 (e3nn default dtype) and rescales afterwards
    to_g = to_g.astype(np.float32)
    fr_g = fr_g.astype(np.float32)
    keep = []
    for l in range(lmax + 1):
        if lmax != mmax and l > mmax:
            f = np.float32(math.sqrt((2 * l + 1) / (2 * mmax + 1)))
            to_g[..., l * l:(l + 1) ** 2] *= f
            fr_g[..., l * l:(l + 1) ** 2] *= f
        mm = min(l, mmax)
        keep += [l * l + l + m for m in range(-mm, mm + 1)]
    return np.ascontiguousarray(to_g[..., keep]), np.ascontiguousarray(fr_g[..., keep])


@functools.lru_cache(maxsize=None)
def s2_grid_factors(lmax, mmax, res):
    """Latitude / longitude factors of `s2_grid_matrices(lmax, mmax, res, res)`:
        to_grid[b,a,(l,m)]   = Pt[|m|,b,l] * trig[a,m]        from_grid[b,a,(l,m)] = Pf[|m|,b,l] * trig[a,m]
    trig[a,m] = 1 (m=0) | sqrt2 cos(m alpha_a) (m>0) | sqrt2 sin(|m| alpha_a) (m<0).
    Returns fp32 arrays Pt, Pf [7,res,7] (zero padded) and ct, st [res,7]."""
    MAXL = 6
    assert lmax <= MAXL and res % 2 == 0
    betas = (np.arange(res) + 0.5) / res * math.pi
    alphas = np.arange(res) / res * 2 * math.pi
    P = _legendre_no_cs(lmax, np.cos(betas), np.sin(betas))
    qw = _dh_weights(res // 2) * res ** 2 / res
    Pt = np.zeros((MAXL + 1, res, MAXL + 1))
    Pf = np.zeros((MAXL + 1, res, MAXL + 1))
    for l in range(lmax + 1):
        nt = math.sqrt(4 * math.pi) / math.sqrt(2 * l + 1) / math.sqrt(lmax + 1)
        nf = math.sqrt(4 * math.pi) * math.sqrt(2 * l + 1) * math.sqrt(lmax + 1)
        f = math.sqrt((2 * l + 1) / (2 * mmax + 1)) if (lmax != mmax and l > mmax) else 1.0
        for m in range(0, min(l, mmax) + 1):
            base = _norm(l, m) * P[(l, m)]
            Pt[m, :, l] = base * nt * f
            Pf[m, :, l] = base * nf * qw * f
    ct = np.zeros((res, MAXL + 1))
    st = np.zeros((res, MAXL + 1))
    ct[:, 0] = 1.0
    for m in range(1, mmax + 1):
        ct[:, m] = math.sqrt(2.0) * np.cos(m * alphas)
        st[:, m] = math.sqrt(2.0) * np.sin(m * alphas)
    return tuple(np.ascontiguousarray(a.astype(np.float32)) for a in (Pt, Pf, ct, st))
