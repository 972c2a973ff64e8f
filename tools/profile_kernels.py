"""Per-kernel sweep for profiles/: achieved GB/s (HBM-bound kernels) and TFLOP/s (tensor kernels) at several
batch sizes.  Each op is captured `reps` times in one CUDA graph and the graph replay is timed with CUDA
events (launch gaps amortised, L2 flushed between replays by a 256 MB memset).  Not a bench line.

    python tools/profile_kernels.py [--only k1,fm,tower,k3,topk] [--sizes 16384,65536,262144,1048576] [--once]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from news_recsys_b200 import ops  # noqa: E402
from news_recsys_b200 import _lib as L  # noqa: E402
from news_recsys_b200.synthetic import MIND_SMALL_ROWS, CFG1_ROWS  # noqa: E402

DEV = "cuda"
flush = None


def timeit(fn, reps=10, iters=5, once=False):
    global flush
    if once:  # profiler mode: a single plain launch
        fn()
        torch.cuda.synchronize()
        return 0.0
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    best = 1e30
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3  # us


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j["bf16_tflops"])
    return 6650.0, 1590.0


def mk_tables(rows, dims):
    t = {k: torch.randn(n, dims[k], device=DEV) for k, n in rows.items()}
    for w in t.values():
        w[0] = 0
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="k1,fm,tower,k3,topk")
    ap.add_argument("--sizes", default="16384,65536,262144,1048576")
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    only = set(a.only.split(","))
    sizes = [int(x) for x in a.sizes.split(",")]
    hbm, tf = peaks()
    rows_out = []

    def rec(name, B, us, qty, bound):
        ach = (qty / (us * 1e-6) / (1e9 if bound == "hbm" else 1e12)) if us > 0 else 0.0
        pk = hbm if bound == "hbm" else tf
        rows_out.append((name, B, us, ach, "GB/s" if bound == "hbm" else "TFLOP/s", ach / pk))
        print(f"{name:34s} B={B:8d} {us:9.2f} us  {ach:9.1f} {'GB/s' if bound == 'hbm' else 'TFLOP/s'}  frac {ach / pk:6.3f}", flush=True)

    for B in sizes:
        if "k1" in only:
            # cfg1 schema: 5 sparse (32/32/16/16/16) + user_history L=50 (D=32, shares item_id)
            dims = {"user_id": 32, "item_id": 32, "category": 16, "subcategory": 16, "user_click_category": 16}
            tables = mk_tables(CFG1_ROWS, dims)
            names = ["category", "item_id", "subcategory", "user_click_category", "user_history", "user_id"]
            Lh = 50
            batch = {k: torch.randint(1, n, (B,), device=DEV) for k, n in CFG1_ROWS.items()}
            lens = torch.randint(0, Lh + 1, (B,), device=DEV)
            mask = (torch.arange(Lh, device=DEV)[None] < lens[:, None]).float()
            batch["user_history"] = torch.randint(1, CFG1_ROWS["item_id"], (B, Lh), device=DEV) * mask.long()
            batch["user_history_mask"] = mask
            specs, col = [], 0
            tid = {k: i for i, k in enumerate(dims)}
            for nme in names:
                t = "item_id" if nme == "user_history" else nme
                specs.append(ops.FeatSpec(nme, t, tid[t], dims[t], Lh if nme == "user_history" else 1, nme == "user_history", col))
                col += dims[t]
            fb = ops.FeatBinding(specs, tables, batch)
            nvalid = float(mask.sum() / B)
            bytes_fwd = B * (5 * 8 + 4 * 112 + Lh * 12 + nvalid * 128 + 4 * col)
            rec("K1 embed_pool_fwd (cfg1, L=50)", B, timeit(lambda: ops.embed_pool_fwd(fb, col), once=a.once), bytes_fwd, "hbm")
            if "k3" in only:
                gout = torch.randn(B, col, device=DEV)
                m_ = [None] * 16
                by_id = [None] * 16
                for k, i in tid.items():
                    by_id[i] = tables[k]
                mm = [None if w is None else torch.zeros_like(w) for w in by_id]
                vv = [None if w is None else torch.zeros_like(w) for w in by_id]
                n_occ = B * (5 + Lh)
                rec("K3 plan (keys + radix sort)", B, timeit(lambda: ops.BwdPlan(fb), reps=3, once=a.once), n_occ * 64, "hbm")
                plan = ops.BwdPlan(fb)
                byt = B * (4 * col + 8 * (5 + Lh)) + B * (112 + nvalid * 32) * 4 * 7
                rec("K3 apply (fused row AdamW)", B, timeit(lambda: ops.embed_bwd_rowopt(plan, gout, by_id, L.BWD_ADAMW, 1e-3, 1, weight_decay=0.01, m_by_id=mm, v_by_id=vv), reps=3, once=a.once), byt, "hbm")
        if "fm" in only:
            dims = {k: 16 for k in MIND_SMALL_ROWS}
            tables = mk_tables(MIND_SMALL_ROWS, dims)
            names = sorted(MIND_SMALL_ROWS)
            batch = {k: torch.randint(1, n, (B,), device=DEV) for k, n in MIND_SMALL_ROWS.items()}
            specs = [ops.FeatSpec(k, k, i, 16, 1, False, i * 16) for i, k in enumerate(names)]
            fb = ops.FeatBinding(specs, tables, batch)
            bias = torch.zeros(1, device=DEV)
            label = (torch.rand(B, device=DEV) < 0.04).float()
            rec("K2 fm_fused_fwd (cfg2)", B, timeit(lambda: ops.fm_fused_fwd(fb, bias, label), once=a.once), B * (5 * 72 + 16), "hbm")
            dl = torch.randn(B, device=DEV)
            rec("K2 fm_fused_bwd (cfg2)", B, timeit(lambda: ops.fm_fused_bwd(fb, dl, 80), once=a.once), B * (5 * 72 + 4 + 320), "hbm")
        if "tower" in only:
            for nm, dims_ in (("Deep 112", [112, 128, 128, 128, 64, 1]), ("DCN 224", [224, 128, 128, 128, 64, 1])):
                ws_ = [torch.randn(dims_[i + 1], dims_[i], device=DEV) / dims_[i] ** 0.5 for i in range(5)]
                bs_ = [torch.zeros(dims_[i + 1], device=DEV) for i in range(5)]
                x = torch.randn(B, dims_[0], device=DEV)
                fl = 2 * sum(dims_[i] * dims_[i + 1] for i in range(5)) * B
                rec(f"K4 tower_fwd inference ({nm})", B, timeit(lambda: ops.tower_fwd(x, ws_, bs_, None, training=False), once=a.once), fl, "tensor")
                rec(f"K4 tower_fwd training ({nm})", B, timeit(lambda: ops.tower_fwd(x, ws_, bs_, None, training=True), once=a.once), fl, "tensor")
                for tr_ in (False, True):   # the pipelined kernel alone: weights prepacked, bf16 input image in place
                    packed = ops.tower_prepack(B, ws_, bs_, None, training=tr_)
                    ops.tower_image_from_rows(x, ops.tower_input_image(packed, B))
                    us_ = timeit(lambda: ops.tower_fwd(None, ws_, bs_, None, training=tr_, packed=packed, ximg_rows=B), once=a.once)
                    rec(f"K4 tower_fwd3 from image, {'training' if tr_ else 'inference'} ({nm})", B, us_, fl, "tensor")
                    if tr_:   # the binding roofline of a training forward: input image in, every hidden image out
                        rec(f"K4 tower_fwd3 from image, training ({nm}) [HBM view]", B, us_, B * 2 * (dims_[0] + sum(dims_[1:5])), "hbm")
                y, ctx = ops.tower_fwd(x, ws_, bs_, None, training=True)
                gy = torch.randn(B, 1, device=DEV)
                us_ = timeit(lambda: ops.tower_bwd(ctx, gy), once=a.once)
                rec(f"K4 tower_bwd dx+dw+reduce ({nm})", B, us_, 2 * fl, "tensor")
                # binding roofline: dX reads the a images, writes the dz images + fp32 grad_x; dW reads a (incl. input) and dz
                rec(f"K4 tower_bwd dx+dw+reduce ({nm}) [HBM view]", B, us_, B * (4 * 2 * sum(dims_[1:5]) + 2 * dims_[0] + 4 * dims_[0]), "hbm")
    if "topk" in only:
        from news_recsys_b200.retrieval import TopkIndex
        N, D = 1_000_000, 128
        c = torch.nn.functional.normalize(torch.randn(N, D, device=DEV), dim=1)
        idx = TopkIndex(c)
        for Q in (1, 64, 256, 1024, 4096):
            q = torch.nn.functional.normalize(torch.randn(Q, D, device=DEV), dim=1)
            us = timeit(lambda: idx.search(q, 100), reps=2, iters=3, once=a.once)
            rec(f"K6 topk_search N=1M D=128 k=100 Q={Q}", Q, us, 2.0 * Q * N * D, "tensor")
    out = os.path.join(ROOT, "gpurun_out", "kernel_sweep.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump([dict(kernel=r[0], size=r[1], us=r[2], achieved=r[3], unit=r[4], frac=r[5]) for r in rows_out], open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
