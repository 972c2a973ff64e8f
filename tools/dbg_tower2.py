import sys, torch
sys.path.insert(0, ".")
from news_recsys_b200 import ops
DEV = "cuda"
def bf(x): return x.to(torch.bfloat16).to(torch.float64)
def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
def r16(x): return (x + 15) & ~15
def al(x): return (x + 255) & ~255
def layout(dims, B, n_sm=148):
    L = len(dims) - 1
    tiny = 1 if (dims[-1] <= 4 and L >= 2) else 0
    n_mma = L - tiny
    K = dims[:-1]; N = dims[1:]; Kp = [r16(k) for k in K]; Np = [r16(n) for n in N]
    nt = (B + 127) // 128
    wb = sum(Kp[l] * Np[l] * 2 for l in range(n_mma))
    o = al(wb) + al(wb)
    act_off, dz_off = [], []
    for l in range(L): act_off.append(o); o += al(nt * Kp[l] * 128 * 2)
    for l in range(L): dz_off.append(o); o += al(nt * Np[l] * 128 * 2)
    return dict(L=L, tiny=tiny, n_mma=n_mma, Kp=Kp, Np=Np, nt=nt, act_off=act_off, dz_off=dz_off)
def image(ws, off, width, nt, B):
    n = nt * width * 128
    raw = ws[off: off + n * 2].view(torch.bfloat16).view(nt, width // 8, 128, 8)   # [tile][chunk][row][8]
    return raw.permute(0, 2, 1, 3).reshape(nt * 128, width)[:B].double().cpu()
for dims, slope, B in [([112, 128, 128, 128, 64, 1], None, 300), ([48, 128, 128, 64, 16], 0.2, 513)]:
    g = torch.Generator().manual_seed(B)
    ws_ = [torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5 for i in range(len(dims) - 1)]
    bs_ = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(len(dims) - 1)]
    x = torch.randn(B, dims[0], generator=torch.Generator().manual_seed(1))
    gy = torch.randn(B, dims[-1], generator=torch.Generator().manual_seed(2))
    lay = layout(dims, B)
    sl = 0.0 if slope is None else slope
    # emulation (fp64 math on bf16-rounded operands)
    a = [bf(x)]
    for l in range(lay["n_mma"]):
        z = a[l] @ bf(ws_[l]).T + bs_[l].double()
        if l == lay["L"] - 1: y_em = z
        else: a.append(bf(torch.where(z > 0, z, z * sl).float()))
    if lay["tiny"]:
        y_em = a[-1] @ ws_[-1].double().T + bs_[-1].double()
    dz = [None] * lay["L"]
    if lay["tiny"]:
        dz[lay["L"] - 1] = bf(gy)
        da = gy.double() @ ws_[-1].double()
        dz[lay["n_mma"] - 1] = bf((da * torch.where(a[lay["n_mma"]] > 0, 1.0, sl)).float())
    else:
        dz[lay["L"] - 1] = bf(gy)
    for l in range(lay["n_mma"] - 1, 0, -1):
        da = dz[l] @ bf(ws_[l])
        dz[l - 1] = bf((da * torch.where(a[l] > 0, 1.0, sl)).float())
    gx_em = dz[0] @ bf(ws_[0])
    xd = x.to(DEV).requires_grad_(True); wd = [w.to(DEV).requires_grad_(True) for w in ws_]; bd = [b.to(DEV).requires_grad_(True) for b in bs_]
    y, tctx = ops.tower_fwd(xd.detach(), [w.detach() for w in wd], [b.detach() for b in bd], slope, training=True)
    gx, gws, gbs = ops.tower_bwd(tctx, gy.to(DEV)); torch.cuda.synchronize()
    ws = tctx[2]
    print(dims, "fwd vs emu", f"{rel(y, y_em):.2e}")
    for l in range(lay["L"]):
        im = image(ws, lay["act_off"][l], lay["Kp"][l], lay["nt"], B)[:, :dims[l]]
        print("  a[%d] image vs emu %.2e" % (l, rel(im, a[l])))
    for l in range(lay["L"] - 1, -1, -1):
        im = image(ws, lay["dz_off"][l], lay["Np"][l], lay["nt"], B)[:, :dims[l + 1]]
        print("  dz[%d] image vs emu %.2e" % (l, rel(im, dz[l])), " maxabs", float(dz[l].abs().max()))
    print("  gx vs emu %.2e" % rel(gx, gx_em))
    for l in range(lay["L"]):
        gw_em = dz[l].T @ a[l]; gb_em = dz[l].sum(0)
        print("  gw[%d] %.2e gb[%d] %.2e" % (l, rel(gws[l], gw_em), l, rel(gbs[l], gb_em)))
    # ---- stepwise check: each backward step recomputed from the GPU's own images (no mask-flip noise) ----
    A = [image(ws, lay["act_off"][l], lay["Kp"][l], lay["nt"], B)[:, :dims[l]] for l in range(lay["L"])]
    DZ = [image(ws, lay["dz_off"][l], lay["Np"][l], lay["nt"], B)[:, :dims[l + 1]] for l in range(lay["L"])]
    for l in range(lay["n_mma"] - 1, 0, -1):
        step = bf(((DZ[l] @ bf(ws_[l])) * torch.where(A[l] > 0, 1.0, sl)).float())
        print("  step dz[%d] <- dz[%d]: %.2e" % (l - 1, l, rel(DZ[l - 1], step)))
    print("  step gx <- dz[0]: %.2e" % rel(gx, DZ[0] @ bf(ws_[0])))
    for l in range(lay["L"]):
        print("  step gw[%d] %.2e gb[%d] %.2e" % (l, rel(gws[l], DZ[l].T @ A[l]), l, rel(gbs[l], DZ[l].sum(0))))
    for l in range(lay["n_mma"]):
        z = A[l] @ bf(ws_[l]).T + bs_[l].double()
        if l < lay["L"] - 1:
            print("  step a[%d] <- a[%d]: %.2e" % (l + 1, l, rel(A[l + 1], bf(torch.where(z > 0, z, z * sl).float()))))
