"""MLP — mirror of reference src/model/model_utils/utils.py:6-17 (same `network` Sequential so the
state_dict keys `network.{0,2,4,...}.{weight,bias}` match), executed as ONE fused bf16 tcgen05 kernel
(nrx_tower_fwd / nrx_tower_bwd) instead of one cuBLAS launch per layer."""
import torch.nn as nn

from ... import ops


class MLP(nn.Module):
    def __init__(self, dims=(16, 32, 32, 1), negative_slope=None):
        super().__init__()
        layers = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2:  # no activation after the last layer
                layers.append(nn.ReLU() if negative_slope is None else nn.LeakyReLU(negative_slope))
        self.network = nn.Sequential(*layers)
        self.negative_slope = negative_slope

    def linears(self):
        return [m for m in self.network if isinstance(m, nn.Linear)]

    def forward(self, x):
        lin = self.linears()
        ws = [m.weight for m in lin]
        bs = [m.bias for m in lin]
        return ops.TowerFn.apply(x, self.negative_slope, len(lin), *ws, *bs)
