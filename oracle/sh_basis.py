"""ORACLE (test infrastructure, never shipped): fp64 real spherical-harmonic basis.

This file restates, with scipy's associated Legendre functions, the conventions of the
third-party `e3nn` package that the reference relies on but does not vendor
(reference call sites: models/EquiformerV2Functions/wigner.py:9-39 (`_Jd`, `_z_rot_mat`),
so3.py:527-533 (`o3.xyz_to_angles`, `o3.angles_to_matrix`), so3.py:584-608 (`ToS2Grid`,
`FromS2Grid` with `.shb`/`.sha`)).  e3nn is un-pinned in env/requirements.txt; wigner.py:5
says "borrowed from e3nn 0.4.0".  PARITY UNPINNED at this boundary: no golden vectors
exist in the reference; the basis below is pinned instead by the mathematical invariants in
tests/test_oracle_invariants.py (Wigner l=1 block == rotation matrix, orthogonality,
from_grid∘to_grid = identity, ...).

Convention: polar axis y, azimuth alpha = atan2(x, z), beta = acos(y);
  Y_{l,m} = N_{l|m|} P_l^{|m|}(cos beta) * { sqrt2 sin(|m| alpha)  m<0 ; 1  m=0 ; sqrt2 cos(m alpha)  m>0 }
with no Condon-Shortley phase and N = sqrt((2l+1)/(4 pi) (l-|m|)!/(l+|m|)!).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import math

import numpy as np
from scipy.special import lpmv


def sh_norm(l, m):
    m = abs(m)
    return math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))


def legendre_table(lmax, cos_beta):
    """N_{l,|m|} P_l^{|m|}(cos beta) without Condon-Shortley phase -> dict[(l, |m|)] -> array."""
    cos_beta = np.asarray(cos_beta, dtype=np.float64)
    out = {}
    for l in range(lmax + 1):
        for m in range(l + 1):
            out[(l, m)] = ((-1.0) ** m) * lpmv(m, l, cos_beta) * sh_norm(l, m)
    return out


def real_sh_angles(lmax, alpha, beta):
    """[..., (lmax+1)^2] real SH ('integral' normalisation) at azimuth alpha / polar beta."""
    alpha = np.asarray(alpha, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    leg = legendre_table(lmax, np.cos(beta))
    cols = []
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            if m < 0:
                trig = math.sqrt(2.0) * np.sin(-m * alpha)
            elif m == 0:
                trig = np.ones_like(alpha)
            else:
                trig = math.sqrt(2.0) * np.cos(m * alpha)
            cols.append(leg[(l, abs(m))] * trig)
    return np.stack(cols, axis=-1)


def real_sh_xyz(lmax, xyz):
    xyz = np.asarray(xyz, dtype=np.float64)
    n = xyz / np.linalg.norm(xyz, axis=-1, keepdims=True)
    beta = np.arccos(np.clip(n[..., 1], -1.0, 1.0))
    alpha = np.arctan2(n[..., 0], n[..., 2])
    return real_sh_angles(lmax, alpha, beta)


def rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def wigner_from_matrix(l, R, rng=None, npts=None):
    """D_l(R) defined by Y_l(R p) = D_l(R) Y_l(p); least squares over random unit vectors."""
    rng = rng or np.random.default_rng(1234 + l)
    npts = npts or 8 * (2 * l + 1) + 16
    p = rng.normal(size=(npts, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    sl = slice(l * l, (l + 1) * (l + 1))
    Yp = real_sh_xyz(l, p)[:, sl]                 # [P, 2l+1]
    Yq = real_sh_xyz(l, p @ R.T)[:, sl]           # Y(R p)
    # Yq^T = D Yp^T  ->  Yp D^T = Yq
    Dt, *_ = np.linalg.lstsq(Yp, Yq, rcond=None)
    return Dt.T


_S_SWAP = np.array([[0, 1, 0], [1, 0, 0], [0, 0, -1]], dtype=np.float64)


def make_jd(lmax):
    """The `_Jd` list of wigner.py:9 (file Jd.pt is git-ignored upstream): J_l = D_l(S)."""
    out = []
    for l in range(lmax + 1):
        J = wigner_from_matrix(l, _S_SWAP)
        J[np.abs(J) < 1e-14] = 0.0
        out.append(J)
    return out


def quadrature_weights(b):
    """Driscoll-Healy weights."""
    k = np.arange(b)
    w = np.array([
        (2.0 / b) * math.sin(math.pi * (2.0 * j + 1.0) / (4.0 * b))
        * np.sum(np.sin((2 * j + 1) * (2 * k + 1) * math.pi / (4.0 * b)) / (2 * k + 1))
        for j in range(2 * b)
    ])
    return w / (2.0 * (2 * b) ** 2)


def s2_grid_angles(res_beta, res_alpha):
    betas = (np.arange(res_beta) + 0.5) / res_beta * math.pi
    alphas = np.arange(res_alpha) / res_alpha * 2 * math.pi
    return betas, alphas


def s2_sha(lmax, alphas):
    """[a, 2 lmax + 1] for m = -lmax..lmax."""
    cols = []
    for m in range(-lmax, lmax + 1):
        if m < 0:
            cols.append(math.sqrt(2.0) * np.sin(-m * alphas))
        elif m == 0:
            cols.append(np.ones_like(alphas))
        else:
            cols.append(math.sqrt(2.0) * np.cos(m * alphas))
    return np.stack(cols, axis=1)


def s2_shb(lmax, betas, norm_l, extra_b=None):
    """[m (2lmax+1), b, i ((lmax+1)^2)]: n_l N P_l^{|m_i|}(cos beta_b) on the matching m row."""
    leg = legendre_table(lmax, np.cos(betas))
    out = np.zeros((2 * lmax + 1, len(betas), (lmax + 1) ** 2))
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            v = norm_l[l] * leg[(l, abs(m))]
            if extra_b is not None:
                v = v * extra_b
            out[m + lmax, :, l * l + l + m] = v
    return out


def to_s2grid_tensors(lmax, res_beta, res_alpha, normalization="component"):
    assert normalization == "component"
    betas, alphas = s2_grid_angles(res_beta, res_alpha)
    n = [math.sqrt(4 * math.pi) / math.sqrt(2 * l + 1) / math.sqrt(lmax + 1) for l in range(lmax + 1)]
    return s2_shb(lmax, betas, n), s2_sha(lmax, alphas)


def from_s2grid_tensors(res_beta, res_alpha, lmax, normalization="component"):
    assert normalization == "component"
    assert res_beta % 2 == 0
    betas, alphas = s2_grid_angles(res_beta, res_alpha)
    n = [math.sqrt(4 * math.pi) * math.sqrt(2 * l + 1) * math.sqrt(lmax + 1) for l in range(lmax + 1)]
    qw = quadrature_weights(res_beta // 2) * res_beta ** 2 / res_alpha
    return s2_shb(lmax, betas, n, extra_b=qw), s2_sha(lmax, alphas)
