"""RadialFunction (reference radial_function.py:5-30): Linear -> [LayerNorm -> SiLU -> Linear]*.
Same `net.<i>` parameter names; the forward runs the grouped-GEMM engine and the fused
LayerNorm+SiLU kernel instead of nn.Sequential."""
import torch.nn as nn

from .. import ops


class RadialFunction(nn.Module):
    def __init__(self, channels_list):
        super().__init__()
        mods = []
        for i in range(1, len(channels_list)):
            mods.append(nn.Linear(channels_list[i - 1], channels_list[i], bias=True))
            if i < len(channels_list) - 1:
                mods.append(nn.LayerNorm(channels_list[i]))
                mods.append(nn.SiLU())
        self.net = nn.Sequential(*mods)

    def hidden_and_last(self, inputs):
        """(activations entering the last Linear, that Linear): lets the graph-attention block fuse the output layer with
        the kernels that consume the radial weights (ops.GatherRotateConvFn)."""
        mods = list(self.net)
        return self._run(inputs, mods[:-1]), mods[-1]

    def forward(self, inputs):
        return self._run(inputs, list(self.net))

    def _run(self, inputs, mods):
        x = inputs
        i = 0
        while i < len(mods):
            lin = mods[i]
            if isinstance(x, ops.FusedEdgeFeatures):        # x_edge as a description: fused GaussianSmearing + layer 1
                x = ops.rbf_linear(x, lin.weight, lin.bias) if lin.out_features % 4 == 0 else \
                    ops.linear(x.dense(), lin.weight, lin.bias)
            else:
                x = ops.linear(x, lin.weight, lin.bias)
            i += 1
            if i < len(mods):
                ln = mods[i]
                x = ops.ln_silu(x, ln.weight, ln.bias, ln.eps)
                i += 2
        return x
