"""ORACLE / TEST INFRASTRUCTURE — golden vectors for the DSSM retrieval model.

`src/model/recall/DSSM/model.py` is not importable as shipped: it imports `BaseModel.base_model`, `model_utils`,
`DataReader` as TOP-LEVEL packages from a hard-coded path on the author's machine (:1-13), imports `faiss`
(not installed here), and calls `self.get_features_embedding` (:151,:168) where BaseModel defines
`get_feature_embedding` (base_model.py:262).  This script runs the UNMODIFIED class anyway:

  * import aliases only: the top-level names are bound in `sys.modules` to the very same reference modules that
    import fine under their package path (`src.model.BaseModel.base_model`, ...); `faiss` is an empty stub module
    (nothing here calls it); `oracle/refshim` stands in for lightning / omegaconf as for the sort models;
  * one attribute alias: `DSSM.get_features_embedding = BaseModel.get_feature_embedding` (the stale method name).

No arithmetic is supplied from outside the reference: towers (:26-44), gather/pool (:148-180), forward with in-batch
negatives and L2 normalisation (:51-73), `infoNCE_loss` (:92-110) and `triplet_loss` (:75-90) are the reference's.
The `torch.randperm` draws of forward (:63) are reproduced by re-seeding and recorded, as is the iteration order of
the feature-name SETS (:150,:167) which decides the column order of the tower inputs in that process.

    python oracle/make_golden_dssm.py        # build container only; writes tests/golden/dssm.npz
"""
import os as _os
import sys as _sys

# set iteration order inside the reference depends on the string hash seed: pin it so that regenerating reproduces the fixture
if _os.environ.get("PYTHONHASHSEED") != "0":
    _os.environ["PYTHONHASHSEED"] = "0"
    _os.execv(_sys.executable, [_sys.executable] + _sys.argv)

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NRX_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import yaml  # noqa: E402
from make_golden import synth_batch, np_sd, GOLD, CFGS  # noqa: E402


def import_reference_dssm():
    import src.model.BaseModel.base_model as bm
    import src.model.model_utils as mu
    import src.model.model_utils.lr_schedule as lrs
    import src.dataset.DataReader.data_reader as dr
    pkg = types.ModuleType("BaseModel"); pkg.base_model = bm
    sys.modules.update({"BaseModel": pkg, "BaseModel.base_model": bm, "model_utils": mu, "model_utils.lr_schedule": lrs,
                        "DataReader": types.ModuleType("DataReader"), "DataReader.data_reader": dr,
                        "faiss": types.ModuleType("faiss")})
    import src.model.recall.DSSM.model as dm
    dm.DSSM.get_features_embedding = bm.BaseModel.get_feature_embedding   # stale method name (:151,:168)
    return dm.DSSM


def main():
    DSSM = import_reference_dssm()
    cfg_name = "deep_hist"
    path = os.path.join(CFGS, f"train_cf_{cfg_name}.yaml")
    cfg = yaml.safe_load(open(path))
    hp = {"lr": 1e-3, "min_lr": 5e-6, "lr_milestones": [3, 9], "negative_sample_rate": 3}
    torch.manual_seed(42)
    model = DSSM(path, {}, hp)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    B = 24
    batch = synth_batch(cfg, B, torch.Generator().manual_seed(4321))
    out = {"kind": np.array("dssm"), "cfg": np.array(cfg_name), "neg_rate": np.array(hp["negative_sample_rate"])}
    out.update(np_sd(model.state_dict()))
    for k, v in batch.items():
        out["in__" + k] = v.numpy()
    out["user_order"] = np.array(list(model.user_feature_names))   # set iteration order of THIS process
    out["item_order"] = np.array(list(model.item_feature_names))
    uv, iv = model.get_user_embedding(batch), model.get_item_embedding(batch)
    out["user_vector"], out["item_vector"] = uv.detach().numpy(), iv.detach().numpy()
    out["user_tower"], out["item_tower"] = model.user_fc(uv).detach().numpy(), model.item_fc(iv).detach().numpy()
    torch.manual_seed(99)
    perms = [torch.randperm(B) for _ in range(hp["negative_sample_rate"])]
    torch.manual_seed(99)
    u, it, neg = model(batch)
    out["neg_perms"] = torch.stack(perms).numpy()
    out["user_emb"], out["item_emb"], out["neg_emb"] = u.detach().numpy(), it.detach().numpy(), neg.detach().numpy()
    mask = batch["label"][:, 1]
    loss = model.infoNCE_loss(u, it, neg, mask=mask)
    out["infonce"] = loss.detach().numpy()
    out["infonce_nomask"] = model.infoNCE_loss(u, it, neg).detach().numpy()
    out["triplet"] = model.triplet_loss(u, it, neg, mask=mask).detach().numpy()
    model.zero_grad()
    loss.backward()
    for n, p in model.named_parameters():
        out["grad__" + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(GOLD, "dssm.npz"), **out)
    print(f"wrote dssm.npz  infoNCE={float(loss):.6f}  user_order={list(model.user_feature_names)} item_order={list(model.item_feature_names)}")


if __name__ == "__main__":
    main()
