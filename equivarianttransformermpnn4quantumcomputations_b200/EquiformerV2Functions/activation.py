"""Activation modules with the reference's names (reference activation.py).

S2Activation / SeparableS2Activation (activation.py:153-192) launch the fused S2 kernel
(to_grid -> SiLU -> from_grid with the [.,18,18,C] grid tensor kept in registers); the small
point-wise activations are parameter-free helpers used by the attention-logit kernel
(SmoothLeakyReLU, activation.py:66-75, is evaluated inside `eqv2_attn_alpha_*`) and are kept as
modules for API parity.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class ScaledSiLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace
        self.scale_factor = 1.6791767923989418

    def forward(self, inputs):
        return F.silu(inputs, inplace=self.inplace) * self.scale_factor

    def extra_repr(self):
        return "scale_factor={}".format(self.scale_factor) + (", inplace=True" if self.inplace else "")


class _GLU(nn.Module):
    def __init__(self, in_channels, out_channels, bias, act):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.w = nn.Linear(in_channels, 2 * out_channels, bias=bias)
        self.act = act

    def forward(self, inputs):
        a, b = self.w(inputs).split(self.out_channels, dim=-1)
        return self.act(a) * b


class ScaledSwiGLU(_GLU):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(in_channels, out_channels, bias, ScaledSiLU())


class SwiGLU(_GLU):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(in_channels, out_channels, bias, nn.SiLU())


class SmoothLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.alpha = negative_slope

    def forward(self, x):
        return 0.5 * (1 + self.alpha) * x + 0.5 * (1 - self.alpha) * x * (2 * torch.sigmoid(x) - 1)

    def extra_repr(self):
        return "negative_slope={}".format(self.alpha)


class ScaledSmoothLeakyReLU(nn.Module):
    def __init__(self):
        super().__init__()
        self.act = SmoothLeakyReLU(0.2)
        self.scale_factor = 1.531320475574866

    def forward(self, x):
        return self.act(x) * self.scale_factor

    def extra_repr(self):
        return "negative_slope={}, scale_factor={}".format(self.act.alpha, self.scale_factor)


class ScaledSigmoid(nn.Module):
    def __init__(self):
        super().__init__()
        self.scale_factor = 1.8467055342154763

    def forward(self, x):
        return torch.sigmoid(x) * self.scale_factor


class GateActivation(nn.Module):
    """Gate activation (activation.py:96-150).  No reference config selects it (`use_gate_act=False`
    everywhere); the module exists so `use_gate_act=True` fails loudly instead of silently."""

    def __init__(self, lmax, mmax, num_channels):
        super().__init__()
        raise NotImplementedError("GateActivation: no reference config uses use_gate_act=True; "
                                  "only the (separable) S2 activation has a kernel")


class S2Activation(nn.Module):
    """x [R, Kr, C] (l-primary reduced) -> from_grid(SiLU(to_grid(x)))."""

    def __init__(self, lmax, mmax):
        super().__init__()
        self.lmax = lmax
        self.mmax = mmax
        self.act = nn.SiLU()

    def forward(self, inputs, SO3_grid):
        mats = SO3_grid[self.lmax][self.mmax].kernel_mats("l")
        return ops.s2_act(inputs, None, mats)


class SeparableS2Activation(nn.Module):
    """scalars [R, C] -> SiLU on the l=0 row; S2 activation on the rest (activation.py:173-192)."""

    def __init__(self, lmax, mmax):
        super().__init__()
        self.lmax = lmax
        self.mmax = mmax
        self.scalar_act = nn.SiLU()
        self.s2_act = S2Activation(lmax, mmax)

    def forward(self, input_scalars, input_tensors, SO3_grid):
        mats = SO3_grid[self.lmax][self.mmax].kernel_mats("l")
        scalars = input_scalars.reshape(input_scalars.shape[0], input_scalars.shape[-1])
        return ops.s2_act(input_tensors, scalars, mats)
