"""CPU-only property tests (hypothesis) of the host-side logic: shard ranges, the batch blob layout, and batch ingestion
against the oracle restatement of the reference DataReader on generated feature files."""
import os

import numpy as np
import torch
import yaml
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import ref_path as R
from tests._golden import GOLD

CFG = os.path.join(GOLD, "configs", "train_cf_deep_hist.yaml")
_cfg = yaml.safe_load(open(CFG))
_L = _cfg["features"]["array_max_length"]["user_history"]


@given(n=st.integers(0, 10_000), world=st.integers(1, 16))
def test_shard_ranges_partition_the_rows(n, world):
    from news_recsys_b200.parallel import shard_range
    prev = 0
    sizes = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        assert lo == prev and hi >= lo
        sizes.append(hi - lo)
        prev = hi
    assert prev == n and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


line = st.builds(
    lambda u, i, c, s, k, hist, lab, order: (u, i, c, s, k, hist, lab, order),
    st.integers(0, 2**31 - 1), st.integers(0, 2**31 - 1), st.integers(0, 17), st.integers(0, 269), st.integers(0, 17),
    st.lists(st.integers(0, 2**31 - 1), min_size=0, max_size=3 * _L), st.lists(st.sampled_from(["0", "1", "0.5", "1e-3"]), min_size=2, max_size=2),
    st.integers(0, 5))


def _render(t):
    u, i, c, s, k, hist, lab, order = t
    items = [f"user_id:{u}", f"item_id:{i}", f"category:{c}", f"subcategory:{s}", f"user_click_category:{k}"]
    items.insert(order, "user_history:" + ",".join(map(str, hist)))
    return " ".join(items) + "\t" + " ".join(lab)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(rows=st.lists(line, min_size=1, max_size=40), data=st.data())
def test_ingest_equals_oracle_datareader(tmp_path_factory, rows, data):
    from news_recsys_b200.ingest import FeatureFile, compile_feature_file
    d = tmp_path_factory.mktemp("prop")
    lines = [_render(t) for t in rows]
    (d / "f.txt").write_text("\n".join(lines) + "\n")
    compile_feature_file(CFG, str(d / "f.txt"), str(d / "f.nrxf"))
    ff = FeatureFile(str(d / "f.nrxf"))
    pick = data.draw(st.lists(st.integers(0, len(rows) - 1), min_size=1, max_size=32))
    got, want = ff.batch(rows=pick), R.datareader_batch(lines, _cfg, pick)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k


@settings(max_examples=30, deadline=None)
@given(B=st.integers(1, 300), i32=st.booleans(), seed=st.integers(0, 1000))
def test_batch_layout_round_trip(B, i32, seed):
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.trainer import BatchLayout
    model = _model()
    layout = BatchLayout(model, B, torch.int32 if i32 else torch.int64)
    g = torch.Generator().manual_seed(seed)
    batch = {}
    for key, dt, shape, off in layout.fields:
        assert off % 256 == 0
        batch[key] = (torch.randint(0, 1000, shape, generator=g).to(dt) if dt in (torch.int32, torch.int64)
                      else torch.rand(shape, generator=g))
    blob = torch.zeros(layout.nbytes, dtype=torch.uint8)
    layout.pack(batch, blob)
    v = layout.views(blob)
    for key in batch:
        assert torch.equal(v[key], batch[key]), key
    ends = sorted((off, off + int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()) for _, dt, shape, off in layout.fields)
    assert all(a[1] <= b[0] for a, b in zip(ends, ends[1:])) and ends[-1][1] <= layout.nbytes


_MODEL = None


def _model():
    global _MODEL
    if _MODEL is None:
        from news_recsys_b200.model.sort.deep.model import Deep
        _MODEL = Deep(CFG)
    return _MODEL
