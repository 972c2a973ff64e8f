"""Synthetic MIND-shaped workloads: the feature schema of the shipped
train_cf_<model>.yaml files (user_id, item_id, category, subcategory, user_click_category,
optional user_history array aliased to item_id) with seeded uniform / Zipf ids.
There is no network in the build or bench environment, so no real MIND data is used anywhere."""
from __future__ import annotations

import copy
from typing import Dict, Optional

import numpy as np
import torch

MIND_SMALL_ROWS = {"user_id": 94058, "item_id": 65239, "category": 18, "subcategory": 270, "user_click_category": 18}
CFG1_ROWS = {"user_id": 50001, "item_id": 65001, "category": 18, "subcategory": 270, "user_click_category": 18}
MIND_LARGE_ROWS = {"user_id": 1000001, "item_id": 160001, "category": 10000000, "subcategory": 270, "user_click_category": 18}

_BASE = {
    "name": "x",
    "paths": {"out_basedir": "", "user_history_path": ""},
    "features": {
        "sparse_feature_names": ["user_id", "item_id", "category", "subcategory", "user_click_category"],
        "dense_feature_names": [], "array_feature_names": [],
        "item_feature_names": ["item_id", "category", "subcategory"],
        "user_feature_names": ["user_id", "user_click_category"],
        "array_max_length": {}},
    "embeddings": {"embedding_size": {}, "embedding_table_size": {}, "share_emb_table_features": {}},
    "dataset": {"batch_size": 512, "num_workers": 0, "pin_memory": True},
    "train_hparams": {"val_freq": 1, "max_epoch": 30, "lr": 1.0e-3, "min_lr": 5.0e-6, "lr_milestones": [40000, 200000],
                      "max_step": 300000, "device": "gpu", "gpus": [0]},
}


def mind_config(kind: str, rows: Optional[Dict[str, int]] = None, history_len: int = 0) -> dict:
    """Config dict with the reference's YAML schema for `kind` in {lr,fm,deep,widedeep,dcn,deepfm}."""
    c = copy.deepcopy(_BASE)
    c["name"] = kind
    rows = dict(rows or MIND_SMALL_ROWS)
    if kind in ("fm", "deepfm"):
        size = {k: 16 for k in rows}                     # sort/fm/train_cf_fm.yaml:31-37
    elif kind == "lr":
        size = {k: 1 for k in rows}
    elif kind == "widedeep":
        size = {"user_id": 32, "item_id": 32, "category": 17, "subcategory": 17, "user_click_category": 17}
        c["wide_and_deep_cfg"] = {"wide_feature_names": ["category", "subcategory", "user_click_category"]}
    else:
        size = {"user_id": 32, "item_id": 32, "category": 16, "subcategory": 16, "user_click_category": 16}
    if kind == "deepfm":
        c["deepfm_cfg"] = {"fm_feature_names": list(rows), "fm_dim": 15}
    c["embeddings"]["embedding_size"] = size
    c["embeddings"]["embedding_table_size"] = rows
    if history_len > 0:
        c["features"]["array_feature_names"] = ["user_history"]
        c["features"]["user_feature_names"] = ["user_id", "user_click_category", "user_history"]
        c["features"]["array_max_length"] = {"user_history": history_len}
        c["embeddings"]["share_emb_table_features"] = {"user_history": "item_id"}
        if kind == "deepfm":
            c["deepfm_cfg"]["fm_feature_names"] = list(rows) + ["user_history"]
    return c


def _draw(n_rows: int, shape, gen: torch.Generator, zipf: float):
    if zipf and zipf > 1.0:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen))
        x = np.random.default_rng(seed).zipf(zipf, size=shape).astype(np.int64)
        return torch.from_numpy(np.minimum(x, n_rows - 1))
    return torch.randint(1, n_rows, shape, generator=gen)


def synth_batch(cfg: dict, B: int, seed: int = 42, zipf: float = 0.0, label_p: float = 0.04,
                id_dtype=torch.int64) -> Dict[str, torch.Tensor]:
    """Default-collated DataReader layout (reference data_reader.py:54-114) on the CPU."""
    gen = torch.Generator().manual_seed(seed)
    feats, emb = cfg["features"], cfg["embeddings"]
    share = emb.get("share_emb_table_features", {}) or {}
    out = {}
    for f in feats["sparse_feature_names"]:
        out[f] = _draw(emb["embedding_table_size"][share.get(f, f)], (B,), gen, zipf).to(id_dtype)
    for f in feats.get("array_feature_names", []) or []:
        Lh = feats["array_max_length"][f]
        lens = torch.randint(0, Lh + 1, (B,), generator=gen)
        mask = (torch.arange(Lh)[None, :] < lens[:, None]).float()
        ids = _draw(emb["embedding_table_size"][share.get(f, f)], (B, Lh), gen, zipf) * mask.long()
        out[f] = ids.to(id_dtype)
        out[f + "_mask"] = mask
    lab = (torch.rand(B, generator=gen) < label_p).float()
    out["label"] = torch.stack([lab, 1.0 - lab], dim=1)
    return out
