"""Measures how far the bf16 tower's gradients are from the fp32 reference gradients (cosine similarity and relative L2
error per parameter) on the tower cases of tests/test_gpu_tower.py and the reference's golden fixtures.  The numbers
(profiles/r2_grad_cosines.json) are what the cosine gates in tests/test_gpu_tower.py / test_gpu_models.py are set from.

    python tests/tools/grad_cosines.py > gpurun_out/r2_grad_cosines.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_path as R  # noqa: E402
from tests._golden import load  # noqa: E402
from tests.test_gpu_tower import CASES, _mk  # noqa: E402

DEV = "cuda"


def cos(a, b):
    a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def rel2(a, b):
    a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    from news_recsys_b200 import ops
    out = {"tower": [], "models": {}}
    for dims, slope, B in CASES:
        ws, bs = _mk(dims, seed=B)
        x = torch.randn(B, dims[0], generator=torch.Generator().manual_seed(1))
        gy = torch.randn(B, dims[-1], generator=torch.Generator().manual_seed(2))
        xr = x.clone().requires_grad_(True)
        wr = [w.clone().requires_grad_(True) for w in ws]
        br = [b.clone().requires_grad_(True) for b in bs]
        R.mlp(xr, wr, br, negative_slope=slope).backward(gy)
        y, tctx = ops.tower_fwd(x.to(DEV), [w.to(DEV) for w in ws], [b.to(DEV) for b in bs], slope, training=True)
        gx, gws, gbs = ops.tower_bwd(tctx, gy.to(DEV))
        rec = {"dims": dims, "B": B, "gx": [cos(gx, xr.grad), rel2(gx, xr.grad)],
               "gw": [[cos(a, b.grad), rel2(a, b.grad)] for a, b in zip(gws, wr)],
               "gb": [[cos(a, b.grad), rel2(a, b.grad)] for a, b in zip(gbs, br)]}
        rec["min_cos"] = min([rec["gx"][0]] + [c for c, _ in rec["gw"]] + [c for c, _ in rec["gb"]])
        out["tower"].append(rec)
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.model.sort.dcn.model import DCN
    from news_recsys_b200.model.sort.widedeep.model import WideDeep
    for name, cls in (("deep", Deep), ("deep_hist", Deep), ("widedeep", WideDeep), ("widedeep_hist", WideDeep), ("dcn", DCN)):
        g = load(name)
        m = cls(g["cfg_path"])
        m.load_state_dict(g["sd"], strict=True)
        m = m.to(DEV)
        batch = {k: v.to(DEV) for k, v in g["batch"].items()}
        prob = m(batch)
        m.bceLoss(prob, batch["label"][:, 0]).backward()
        params = dict(m.named_parameters())
        rec = {k: [cos(params[k].grad, gr), rel2(params[k].grad, gr)] for k, gr in g["grads"].items() if float(gr.abs().max()) > 0}
        out["models"][name] = {"min_cos": min(c for c, _ in rec.values()), "max_rel_l2": max(r for _, r in rec.values()), "per_param": rec}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
