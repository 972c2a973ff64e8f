"""DCN layers — mirror of reference src/model/sort/dcn/dcn_arch.py.  `DCNLayer`/`DCNNet` keep the
reference parameter shapes (w, b: [d, 1]; state_dict keys `cross_net.{i}.{w,b}`); the stack runs as one
fused kernel (nrx_dcn_cross_fwd) using x0 * (xl . w) + b + xl instead of the reference's [B, d, d]
outer product (:25).  `DCNv2Layer`/`DCNv2Net` (:33-50, :73-91; defined but unused by the reference DCN)
run their d x d Linear on the tcgen05 tower kernel."""
import torch
import torch.nn as nn

from .... import ops


class DCNLayer(nn.Module):
    def __init__(self, dim=32):
        super().__init__()
        self.w = nn.Parameter(torch.empty(dim, 1))
        self.b = nn.Parameter(torch.zeros(dim, 1))
        nn.init.xavier_uniform_(self.w)

    def forward(self, x_l, x_0):
        """Single layer with the reference signature (:14-30)."""
        d = x_0.shape[1]
        # one layer of the recurrence: x0 * (xl . w) + b + xl
        s = ops.FieldLogitFn.apply((x_l * self.w.view(1, -1)).contiguous(), [0], [d], ops.L.FIELD_SUM)
        return x_0 * s.view(-1, 1) + self.b.view(1, -1) + x_l


class DCNNet(nn.Module):
    def __init__(self, input_dim, num_layers=3):
        super().__init__()
        self.cross_net = nn.ModuleList([DCNLayer(input_dim) for _ in range(num_layers)])

    def cat_forward(self, x):
        """cat[x, x_L] in one kernel (what DCNModel.forward needs, dcn/model.py:29)."""
        ws = [l.w for l in self.cross_net]
        bs = [l.b for l in self.cross_net]
        return ops.CrossFn.apply(x, len(ws), *ws, *bs)

    def forward(self, x):
        d = x.shape[1]
        return self.cat_forward(x)[:, d:]


class DCNv2Layer(nn.Module):
    def __init__(self, dim=32):
        super().__init__()
        self.linear = nn.Linear(dim, dim, bias=True)

    def forward(self, x_l, x_0):
        lin = ops.TowerFn.apply(x_l, None, 1, self.linear.weight, self.linear.bias)
        return x_0 * lin + x_l


class DCNv2Net(nn.Module):
    def __init__(self, input_dim, num_layers=3):
        super().__init__()
        layers = []
        for _ in range(num_layers):
            layers.append(DCNv2Layer(input_dim))
            layers.append(nn.ReLU())
        self.cross_net = nn.ModuleList(layers)

    def forward(self, x):
        x_0 = x
        for layer in self.cross_net:
            x = layer(x, x_0) if isinstance(layer, DCNv2Layer) else layer(x)
        return x
