"""HTR and GATAValueActivation with the reference's parameter names
(reference NewFunctions/Gotennet_morethaninspired/activation.py:166-264, :270-414); the base activations are
re-exported as the reference fork does.

All Linear layers run on the grouped-GEMM kernels (`ops.linear`, differentiable to any order); the per-degree vector
rejections + inner products of HTR and the output assembly of GATAValueActivation are kernels of their own
(`ops.htr_inner`, `ops.gata_value`, csrc/gata.cu) whose backward passes are kernels again, so the double backward of the
force loss (they carry position gradients through `t_ij`) never leaves the library.  What stays in torch are the three
gate products (`gamma_w * gamma_t`, `W_rs(t) * SiLU(gamma_s(h))`, `* phi_proj`)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...EquiformerV2Functions.activation import (GateActivation, S2Activation, ScaledSigmoid, ScaledSiLU,  # noqa: F401
                                                  ScaledSmoothLeakyReLU, ScaledSwiGLU, SeparableS2Activation,
                                                  SmoothLeakyReLU, SwiGLU)


def _lin(layer, x):
    """nn.Linear parameters applied by the GEMM kernels to the last dimension of x."""
    shp = x.shape
    y = ops.linear(x.reshape(-1, shp[-1]), layer.weight, layer.bias)
    return y.reshape(*shp[:-1], layer.weight.shape[0])


class HTR(nn.Module):
    """t_ij <- t_ij + gamma_w(w_ij) * gamma_t(t_ij),  w_ij = sum_l <reject(W_vq X_i^l, r^l), reject(W_vk^l X_j^l, -r^l)> / (2l+1)."""

    def __init__(self, sphere_channels, edge_channels, lmax, hidden_channels=None):
        super().__init__()
        self.lmax = lmax
        self.edge_channels = edge_channels
        hidden_channels = hidden_channels or edge_channels
        self.degree_sizes = [2 * l + 1 for l in range(1, lmax + 1)]
        self.W_vq = nn.Linear(sphere_channels, hidden_channels, bias=False)
        self.W_vk = nn.ModuleList([nn.Linear(sphere_channels, hidden_channels, bias=False) for _ in range(lmax)])
        self.gamma_w = nn.Sequential(nn.Linear(edge_channels, edge_channels), nn.SiLU())
        self.gamma_t = nn.Sequential(nn.Linear(edge_channels, edge_channels), nn.SiLU(),
                                     nn.Linear(edge_channels, edge_channels), nn.SiLU())
        for module in self.gamma_w:
            if isinstance(module, nn.Linear):
                nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.zeros_(module.bias)

    @staticmethod
    def vector_rejection(rep, rl):
        rl_u = rl.unsqueeze(-1)
        return rep - (rep * rl_u).sum(dim=1, keepdim=True) * rl_u

    def forward(self, t_ij, X_i, X_j, rl_ij):
        q_all = _lin(self.W_vq, X_i)                                  # one GEMM for all degrees
        k_parts, off = [], 0
        for l_idx, n in enumerate(self.degree_sizes):
            k_parts.append(_lin(self.W_vk[l_idx], X_j[:, off:off + n]))
            off += n
        # sum_l <reject(q^l, r^l), reject(k^l, -r^l)> / (2l+1) in ONE kernel, differentiable to any order through its
        # gradient-map kernel (ops.HtrInnerFn / HtrGradFn) -- formerly ~10 element-wise / reduction launches per degree
        w_ij = ops.htr_inner(q_all, torch.cat(k_parts, dim=1), rl_ij.detach(), self.lmax)
        gw = F.silu(_lin(self.gamma_w[0], w_ij))
        gt = F.silu(_lin(self.gamma_t[2], F.silu(_lin(self.gamma_t[0], t_ij))))
        return t_ij + gw * gt


class GATAValueActivation(nn.Module):
    """combined = attn_output + W_rs(t_ij) * gamma_s(h_j) -> (o_s | o_d^l | o_t^l);
    out_0 = SiLU(o_s); out_l = o_d^l * r^l + o_t^l * (xj_proj X_j)^l on the first min(2l+1, 2 mmax+1) rows."""

    def __init__(self, sphere_channels, hidden_channels, edge_channels, lmax, mmax, num_rbf=None):
        super().__init__()
        self.lmax = lmax
        self.mmax = mmax
        self.hidden_channels = hidden_channels
        self.S = 1 + 2 * lmax
        self.W_rs = nn.Linear(edge_channels, self.S * hidden_channels)
        if num_rbf is not None:
            # `Gotennets_GATA_phi_refined_every_layer` variant (activation.py:314,352): extra factor phi_proj(phi_r)
            self.phi_proj = nn.Linear(num_rbf, self.S * hidden_channels, bias=False)
        self.gamma_s = nn.Sequential(nn.Linear(sphere_channels, self.S * hidden_channels), nn.SiLU())
        self.xj_proj = nn.Linear(sphere_channels, hidden_channels, bias=False)
        self.scalar_act = nn.SiLU()
        self.full_degree_sizes = [2 * l + 1 for l in range(1, lmax + 1)]
        self.reduced_degree_sizes = [min(2 * l + 1, 2 * mmax + 1) for l in range(1, lmax + 1)]

    def forward(self, attn_output, t_ij, h_j, X_j, rl_ij, phi_r=None):
        C = self.hidden_channels
        bias = _lin(self.W_rs, t_ij) * F.silu(_lin(self.gamma_s[0], h_j))
        if phi_r is not None:
            bias = bias * _lin(self.phi_proj, phi_r)
        combined = attn_output + bias
        Xp = _lin(self.xj_proj, X_j)
        # SiLU(o_s) | o_d^l r^l + o_t^l Xp^l on the first min(2l+1, 2 mmax+1) rows of each degree: one kernel per derivative
        # order (ops.GataValueFn) instead of split / broadcast-multiply / add / cat per degree
        return ops.gata_value(combined, Xp, rl_ij.detach(), self.lmax, self.mmax)
