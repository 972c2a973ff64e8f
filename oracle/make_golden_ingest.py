"""ORACLE / TEST INFRASTRUCTURE — golden vectors for batch ingestion.

Writes a small feature text file in the reference's wire format ("name:value ... \\t labels",
src/dataset/DataReader/data_reader.py:57-60) and the batches the reference's OWN `DataReader` + torch default collate
(pl_dataloader.py:77-95) produce from it, to tests/golden/ingest_features.txt / ingest.npz.

    python oracle/make_golden_ingest.py      # build container only
"""
import os as _os
import sys as _sys

# set iteration order inside the reference depends on the string hash seed: pin it so that regenerating reproduces the fixture
if _os.environ.get("PYTHONHASHSEED") != "0":
    _os.environ["PYTHONHASHSEED"] = "0"
    _os.execv(_sys.executable, [_sys.executable] + _sys.argv)

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NRX_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)

import yaml  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG = os.path.join(GOLD, "configs", "train_cf_deep_hist.yaml")


def write_text(path, n=200, seed=5):
    cfg = yaml.safe_load(open(CFG))
    emb = cfg["embeddings"]
    share = emb.get("share_emb_table_features", {}) or {}
    feats = cfg["features"]
    rng = np.random.default_rng(seed)
    L = feats["array_max_length"]["user_history"]
    lines = []
    for i in range(n):
        items = []
        for f in feats["sparse_feature_names"]:
            rows = emb["embedding_table_size"][share.get(f, f)]
            items.append(f"{f}:{int(rng.integers(0, rows))}")
        rows = emb["embedding_table_size"][share.get("user_history", "user_history")]
        k = int(rng.integers(0, 2 * L))            # some longer than max_len -> truncated to the first L
        if i == 0:
            k = 0                                     # empty history: "user_history:"
        if i == 1:
            k = L                                     # exactly full
        if i == 2:
            k = L + 5                                 # over-long
        hist = ",".join(str(int(x)) for x in rng.integers(1, rows, size=k))
        items.insert(int(rng.integers(0, len(items) + 1)), f"user_history:{hist}")   # position in the line varies
        items.append(f"not_in_config:{i}")            # unknown names are ignored (data_reader.py:70-105 falls through)
        click = float(rng.random() < 0.3)
        lines.append(" ".join(items) + "\t" + f"{click:g} {1 - click:g}")
        if i % 50 == 10:
            lines.append("")                          # blank lines are skipped (:49)
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")


def main():
    from torch.utils.data import default_collate
    from src.dataset.DataReader.data_reader import DataReader
    txt = os.path.join(GOLD, "ingest_features.txt")
    write_text(txt)
    ds = DataReader(CFG, txt)
    out = {"n_rows": np.array(len(ds))}
    rng = np.random.default_rng(11)
    cases = {"seq": list(range(0, 64)), "tail": list(range(len(ds) - 7, len(ds))), "shuf": rng.permutation(len(ds))[:48].tolist(),
             "dup": [3, 3, 0, 199, 3]}
    for name, rows in cases.items():
        b = default_collate([ds[i] for i in rows])
        out[f"{name}__rows"] = np.array(rows)
        for k, v in b.items():
            out[f"{name}__{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "ingest.npz"), **out)
    print("wrote ingest_features.txt,", len(ds), "rows; ingest.npz keys:", sorted(k for k in out if k.startswith("seq__")))


if __name__ == "__main__":
    main()
