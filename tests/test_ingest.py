"""CPU: batch ingestion — the columnar feature file + host-side batch assembly (`nrx_ingest_*`) against
batches produced by the reference's own DataReader + default collate (tests/golden/ingest.npz) and against the oracle
restatement on random inputs.  No GPU work: these entry points are host code inside libnrx."""
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import ref_path as R
from tests._golden import GOLD

CFG = os.path.join(GOLD, "configs", "train_cf_deep_hist.yaml")
TXT = os.path.join(GOLD, "ingest_features.txt")
CASES = ["seq", "tail", "shuf", "dup"]


def _lines():
    return [l.strip() for l in open(TXT, encoding="utf-8") if l.strip()]


@pytest.fixture(scope="module")
def feature_file(tmp_path_factory):
    from news_recsys_b200.ingest import FeatureFile, compile_feature_file
    out = str(tmp_path_factory.mktemp("ingest") / "features.nrxf")
    stats = compile_feature_file(CFG, TXT, out)
    assert stats["n_rows"] == 200 and stats["n_labels"] == 2
    return FeatureFile(out)


@pytest.mark.parametrize("case", CASES)
def test_oracle_datareader_matches_reference(case):
    z = np.load(os.path.join(GOLD, "ingest.npz"))
    cfg = yaml.safe_load(open(CFG))
    b = R.datareader_batch(_lines(), cfg, z[f"{case}__rows"].tolist())
    keys = {k[len(case) + 2:] for k in z.files if k.startswith(case + "__")} - {"rows"}
    assert set(b) == keys
    for k in keys:
        ref = torch.from_numpy(z[f"{case}__{k}"])
        assert b[k].dtype == ref.dtype and torch.equal(b[k], ref), k


@pytest.mark.parametrize("case", CASES)
def test_feature_file_batches_equal_reference_dataloader(feature_file, case):
    z = np.load(os.path.join(GOLD, "ingest.npz"))
    rows = z[f"{case}__rows"].tolist()
    b = feature_file.batch(rows=rows)
    keys = {k[len(case) + 2:] for k in z.files if k.startswith(case + "__")} - {"rows"}
    assert set(b) == keys
    for k in keys:
        ref = torch.from_numpy(z[f"{case}__{k}"])
        assert b[k].dtype == ref.dtype and torch.equal(b[k], ref), k
    if case == "seq":   # contiguous rows through the row0 path
        b2 = feature_file.batch(start=0, B=64)
        assert all(torch.equal(b2[k], b[k]) for k in b)


@pytest.mark.parametrize("id_dtype", [torch.int64, torch.int32])
def test_pack_writes_the_trainer_blob(feature_file, id_dtype):
    """pack() == BatchLayout.pack(batch()) byte for byte, for contiguous and shuffled rows, int64 and int32 ids."""
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.trainer import BatchLayout
    model = Deep(CFG)
    layout = BatchLayout(model, 48, id_dtype)
    rows = np.random.default_rng(3).permutation(200)[:48]
    for kw in (dict(rows=rows), dict(start=100)):
        want = torch.zeros(layout.nbytes, dtype=torch.uint8)
        b = feature_file.batch(rows=kw.get("rows"), start=kw.get("start", 0), B=48)
        layout.pack(b, want)
        got = torch.zeros(layout.nbytes, dtype=torch.uint8)
        feature_file.pack(layout, got, **kw)
        assert torch.equal(got, want)
    with pytest.raises(Exception):
        feature_file.pack(layout, torch.zeros(8, dtype=torch.uint8), start=0)
    with pytest.raises(Exception):
        feature_file.pack(layout, torch.zeros(layout.nbytes, dtype=torch.uint8), start=180)   # runs past the file


def test_prefetcher_yields_every_batch_in_order(feature_file):
    """Worker threads pack ahead into a small ring; every yielded blob equals pack() of that batch at yield time."""
    from news_recsys_b200.ingest import BlobPrefetcher
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.trainer import BatchLayout
    layout = BatchLayout(Deep(CFG), 16, torch.int32)
    rng = np.random.default_rng(9)
    batches = [rng.permutation(200)[:16] for _ in range(23)] + [0, 16, 184]
    n = 0
    for i, blob in enumerate(BlobPrefetcher(feature_file, layout, batches, depth=3, workers=3, pin=False)):
        want = torch.zeros(layout.nbytes, dtype=torch.uint8)
        b = batches[i]
        feature_file.pack(layout, want, **(dict(start=int(b)) if isinstance(b, int) else dict(rows=b)))
        assert torch.equal(blob, want), i
        n += 1
    assert n == len(batches)
    with pytest.raises(Exception):   # a bad batch surfaces in the consumer instead of hanging it
        for _ in BlobPrefetcher(feature_file, layout, [0, 195], depth=3, workers=1, pin=False):
            pass


def test_random_files_equal_oracle(tmp_path):
    """Random schema-conforming files (ragged arrays incl. empty and over-long, ids up to 2^31-1) vs the oracle."""
    from news_recsys_b200.ingest import FeatureFile, compile_feature_file
    cfg = yaml.safe_load(open(CFG))
    rng = np.random.default_rng(0)
    L = cfg["features"]["array_max_length"]["user_history"]
    lines = []
    for i in range(300):
        k = int(rng.integers(0, 3 * L))
        hist = ",".join(str(int(x)) for x in rng.integers(0, 2**31 - 1, size=k))
        lines.append(f"user_history:{hist} user_id:{int(rng.integers(0, 2**31 - 1))} item_id:{i} category:1 subcategory:0 "
                     f"user_click_category:17\t{float(rng.random()):.6f} {int(rng.integers(0, 2))} 0.5")
    txt = tmp_path / "r.txt"
    txt.write_text("\n".join(lines) + "\n")
    compile_feature_file(CFG, str(txt), str(tmp_path / "r.nrxf"))
    ff = FeatureFile(str(tmp_path / "r.nrxf"))
    assert len(ff) == 300 and ff.n_labels == 3
    rows = rng.permutation(300)[:77].tolist()
    got, want = ff.batch(rows=rows), R.datareader_batch(lines, cfg, rows)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k


def test_errors_follow_the_reference(tmp_path):
    from news_recsys_b200.ingest import compile_feature_file
    cases = {"user_id:1 item_id:2 1 0": "missing tab separator", "user_id:1 item_id\t1 0": "does not contain ':'"}
    for text, msg in cases.items():
        p = tmp_path / "bad.txt"
        p.write_text(text + "\n")
        with pytest.raises(ValueError, match=msg):
            compile_feature_file(CFG, str(p), str(tmp_path / "bad.nrxf"))
        with pytest.raises(ValueError, match=msg):
            R.datareader_getitem(text, yaml.safe_load(open(CFG)))
    with pytest.raises(FileNotFoundError):
        compile_feature_file(CFG, str(tmp_path / "nope.txt"), str(tmp_path / "x.nrxf"))
    p = tmp_path / "ragged.txt"
    p.write_text("user_id:1 item_id:2\t1 0\nuser_id:1\t1 0\n")
    with pytest.raises(ValueError, match="differ from line 0"):
        compile_feature_file(CFG, str(p), str(tmp_path / "x.nrxf"))


def test_large_batches_split_over_threads_equal_small_ones(tmp_path):
    """With NRX_INGEST_THREADS > 1 a batch of >= 8192 rows is assembled by several host threads (software prefetch on
    shuffled rows); the result equals the single-thread path used for small batches.  Run in a fresh process: the
    thread count is read once."""
    import subprocess, sys, textwrap
    code = textwrap.dedent(f"""
        import os, sys
        os.environ["NRX_INGEST_THREADS"] = "4"
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        import tests.test_ingest as T, pathlib
        T._threads_case(pathlib.Path({str(tmp_path)!r}))
        print("ok")
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
    _threads_case(tmp_path)   # and with the default single thread


def _threads_case(tmp_path):
    from news_recsys_b200.ingest import FeatureFile, compile_feature_file
    rng = np.random.default_rng(2)
    lines = []
    for i in range(12000):
        hist = ",".join(str(int(x)) for x in rng.integers(1, 299, size=int(rng.integers(0, 20))))
        lines.append(f"user_id:{i % 399} item_id:{(7 * i) % 299} category:{i % 18} subcategory:2 user_click_category:3 "
                     f"user_history:{hist}\t{i % 2} {1 - i % 2}")
    (tmp_path / "t.txt").write_text("\n".join(lines) + "\n")
    compile_feature_file(CFG, str(tmp_path / "t.txt"), str(tmp_path / "t.nrxf"))
    ff = FeatureFile(str(tmp_path / "t.nrxf"))
    rows = rng.permutation(12000)[:10000]
    big = ff.batch(rows=rows)
    small = [ff.batch(rows=rows[i:i + 2500]) for i in range(0, 10000, 2500)]
    for k in big:
        assert torch.equal(big[k], torch.cat([s[k] for s in small])), k
    seq = ff.batch(start=1000, B=9000)
    seq_small = [ff.batch(start=1000 + i, B=3000) for i in range(0, 9000, 3000)]
    for k in seq:
        assert torch.equal(seq[k], torch.cat([s[k] for s in seq_small])), k
