import sys, torch
sys.path.insert(0, ".")
from oracle import ref_path as R
from news_recsys_b200 import ops
DEV = "cuda"
def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
def mk(dims, seed):
    g = torch.Generator().manual_seed(seed)
    ws = [torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5 for i in range(len(dims) - 1)]
    bs = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(len(dims) - 1)]
    return ws, bs
for dims, slope, B in [([112, 128, 128, 128, 64, 1], None, 300), ([48, 128, 128, 64, 16], 0.2, 513), ([20, 16, 8, 1], None, 77), ([224,128,128,128,64,1], None, 4096)]:
    ws, bs = mk(dims, B)
    x = torch.randn(B, dims[0], generator=torch.Generator().manual_seed(1))
    gy = torch.randn(B, dims[-1], generator=torch.Generator().manual_seed(2))
    xr = x.clone().requires_grad_(True); wr = [w.clone().requires_grad_(True) for w in ws]; br = [b.clone().requires_grad_(True) for b in bs]
    yr = R.mlp(xr, wr, br, negative_slope=slope); yr.backward(gy)
    xd = x.to(DEV).requires_grad_(True); wd = [w.to(DEV).requires_grad_(True) for w in ws]; bd = [b.to(DEV).requires_grad_(True) for b in bs]
    y = ops.TowerFn.apply(xd, slope, len(wd), *wd, *bd)
    y.backward(gy.to(DEV)); torch.cuda.synchronize()
    print(dims, "B", B, "fwd", f"{rel(y, yr):.2e}", "gx", f"{rel(xd.grad, xr.grad):.2e}",
          "gw", [f"{rel(a.grad, b.grad):.2e}" for a, b in zip(wd, wr)], "gb", [f"{rel(a.grad, b.grad):.2e}" for a, b in zip(bd, br)])
