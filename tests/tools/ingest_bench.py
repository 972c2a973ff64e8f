"""Host-side measurement of batch ingestion: rows/s of FeatureFile.pack() into the trainer's blob vs the
CPU port of the reference loader (oracle.ref_path.datareader_batch == DataReader.__getitem__ + default collate) on the
same synthetic feature file.  Pure host code: runs anywhere.

    python tests/tools/ingest_bench.py [--rows 200000] [--batch 16384] > profiles/r1_ingest_bench.json
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from news_recsys_b200.ingest import FeatureFile, compile_feature_file  # noqa: E402
from news_recsys_b200.trainer import BatchLayout  # noqa: E402
from oracle import ref_path as R  # noqa: E402  (CPU baseline leg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--batch", type=int, default=16384)
    a = ap.parse_args()
    cfg_path = os.path.join(ROOT, "tests", "golden", "configs", "train_cf_deep_hist.yaml")
    cfg = yaml.safe_load(open(cfg_path))
    emb, feats = cfg["embeddings"], cfg["features"]
    share = emb.get("share_emb_table_features", {}) or {}
    L = feats["array_max_length"]["user_history"]
    rng = np.random.default_rng(0)
    d = tempfile.mkdtemp()
    txt, binf = os.path.join(d, "f.txt"), os.path.join(d, "f.nrxf")
    cols = {f: rng.integers(1, emb["embedding_table_size"][share.get(f, f)], size=a.rows) for f in feats["sparse_feature_names"]}
    lens = rng.integers(0, L + 1, size=a.rows)
    hist_rows = emb["embedding_table_size"][share.get("user_history", "user_history")]
    with open(txt, "w") as f:
        for i in range(a.rows):
            h = ",".join(map(str, rng.integers(1, hist_rows, size=lens[i])))
            f.write(" ".join(f"{k}:{v[i]}" for k, v in cols.items()) + f" user_history:{h}\t{i % 2} {1 - i % 2}\n")
    t0 = time.perf_counter()
    stats = compile_feature_file(cfg_path, txt, binf)
    t_compile = time.perf_counter() - t0
    ff = FeatureFile(binf)
    from news_recsys_b200.model.sort.deep.model import Deep
    layout = BatchLayout(Deep(cfg_path), a.batch, torch.int32)
    blob = torch.zeros(layout.nbytes, dtype=torch.uint8)
    n_b = a.rows // a.batch
    perm = rng.permutation(a.rows)
    res = {}
    for name, kw in (("sequential", lambda i: dict(start=i * a.batch)), ("shuffled", lambda i: dict(rows=perm[i * a.batch:(i + 1) * a.batch]))):
        ff.pack(layout, blob, **kw(0))
        t0 = time.perf_counter()
        for rep in range(3):
            for i in range(n_b):
                ff.pack(layout, blob, **kw(i))
        dt = (time.perf_counter() - t0) / (3 * n_b)
        res[name] = {"ms_per_batch": dt * 1e3, "rows_per_s": a.batch / dt}
    from news_recsys_b200.ingest import BlobPrefetcher
    for workers in (1, 4):
        order = [perm[i * a.batch:(i + 1) * a.batch] for i in range(n_b)] * 3
        t0 = time.perf_counter()
        n = sum(1 for _ in BlobPrefetcher(ff, layout, order, depth=6, workers=workers, pin=False))
        dt = (time.perf_counter() - t0) / n
        res[f"prefetcher_shuffled_{workers}_workers"] = {"ms_per_batch": dt * 1e3, "rows_per_s": a.batch / dt}
    lines = [l.strip() for l in open(txt)]
    sample = min(a.batch, 4096)
    t0 = time.perf_counter()
    R.datareader_batch(lines, cfg, perm[:sample].tolist())
    t_ref = time.perf_counter() - t0
    out = {"workload": f"{a.rows} rows, schema train_cf_deep_hist (5 sparse ids + user_history L={L}), batch {a.batch}, int32 ids",
           "file_bytes": stats["bytes"], "text_bytes": os.path.getsize(txt), "compile_s_once": t_compile,
           "pack": res, "blob_bytes": layout.nbytes,
           "cpu_baseline": {"kind": "port", "what": "oracle.ref_path.datareader_batch (DataReader.__getitem__ + default collate), 1 core",
                            "sample": f"{sample} shuffled rows", "rows_per_s": sample / t_ref},
           "host": {"cores": os.cpu_count()}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
