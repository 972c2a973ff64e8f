"""Helpers shared by the tests: load the reference-produced fixtures in tests/golden."""
import os

import numpy as np
import torch
import yaml

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFGS = os.path.join(GOLD, "configs")

MODEL_FIXTURES = ["lr", "fm", "fm_hist", "fm_soft", "fm_hist_soft", "deep", "deep_hist", "deep_hist_nomask",
                  "widedeep", "widedeep_hist", "dcn", "dcn_hist"]


def load(name):
    z = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=False)
    kind, cfg_name = str(z["kind"]), str(z["cfg"])
    cfg_path = os.path.join(CFGS, f"train_cf_{cfg_name}.yaml")
    cfg = yaml.safe_load(open(cfg_path))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    batch = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in__")}
    grads = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad__")}
    return dict(z=z, kind=kind, cfg=cfg, cfg_path=cfg_path, sd=sd, batch=batch, grads=grads)


def opt_batches(z):
    out, s = [], 0
    while f"opt{s}_loss" in z.files:
        pre = f"opt{s}_in__"
        out.append({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)})
        s += 1
    return out
