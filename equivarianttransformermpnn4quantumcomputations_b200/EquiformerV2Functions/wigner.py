"""wigner_D / _z_rot_mat (reference wigner.py:17-39) for API parity.  The hot path never calls
these: SO3_Rotation.set_wigner launches eqv2_wigner_from_rot.  `_Jd` is regenerated from first
principles (the reference loads a git-ignored Jd.pt)."""
import torch

from .. import _so3_math

_Jd = [torch.from_numpy(J.copy()) for J in _so3_math.jd_blocks(8)]


def _z_rot_mat(angle, l):
    n = 2 * l + 1
    M = angle.new_zeros((*angle.shape, n, n))
    idx = torch.arange(n, device=angle.device)
    freq = torch.arange(l, -l - 1, -1, dtype=angle.dtype, device=angle.device)
    M[..., idx, n - 1 - idx] = torch.sin(freq * angle[..., None])
    M[..., idx, idx] = torch.cos(freq * angle[..., None])
    return M


def wigner_D(l, alpha, beta, gamma):
    if not l < len(_Jd):
        raise NotImplementedError(f"wigner D maximum l implemented is {len(_Jd) - 1}")
    alpha, beta, gamma = torch.broadcast_tensors(alpha, beta, gamma)
    J = _Jd[l].to(dtype=alpha.dtype, device=alpha.device)
    return _z_rot_mat(alpha, l) @ J @ _z_rot_mat(beta, l) @ J @ _z_rot_mat(gamma, l)
