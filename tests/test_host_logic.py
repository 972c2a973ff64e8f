"""CPU-only: host-side mirror of the reference interface (config schema, column layout, state_dict keys,
LR schedule, batch blob layout) — no GPU work."""
import os

import pytest
import torch

from tests._golden import MODEL_FIXTURES, load


def _cls(kind):
    import importlib
    name = {"fm": "FM", "deep": "Deep", "widedeep": "WideDeep", "dcn": "DCN", "deepfm": "DeepFM", "lr": "LR"}[kind]
    return getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_state_dict_keys_and_shapes_match_reference(name):
    """Reference checkpoints must load with strict=True (base_model.py:531-536)."""
    g = load(name)
    m = _cls(g["kind"])(g["cfg_path"])
    sd = m.state_dict()
    assert set(sd) == set(g["sd"]), set(sd) ^ set(g["sd"])
    for k, v in g["sd"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    m.load_state_dict(g["sd"], strict=True)


@pytest.mark.parametrize("name", ["deep_hist", "widedeep_hist", "fm"])
def test_column_layout_is_sorted_by_feature_name(name):
    """base_model.py:286 — concat order is alphabetical; SURVEY §8g offsets for the cfg1 schema."""
    g = load(name)
    m = _cls(g["kind"])(g["cfg_path"])
    names = sorted(m.user_feature_names | m.item_feature_names)
    batch = {k: v for k, v in g["batch"].items()}
    specs, dims = m._specs_for(names, batch)
    assert [s.name for s in specs] == g["z"]["names"].tolist()
    assert dims == g["z"]["dims"].tolist()
    cols, c = [], 0
    for d in dims:
        cols.append(c)
        c += d
    assert [s.out_col for s in specs] == cols
    if name == "deep_hist":
        assert dict(zip(names, cols)) == {"category": 0, "item_id": 16, "subcategory": 48, "user_click_category": 64,
                                          "user_history": 80, "user_id": 112}
        assert specs[names.index("user_history")].table == "item_id"  # share_emb_table_features
        assert m.user_input_dim + m.item_input_dim == 144


def test_missing_config_raises_file_not_found():
    from news_recsys_b200.model.sort.deep.model import Deep
    with pytest.raises(FileNotFoundError):
        Deep("/nonexistent/train_cf_deep.yaml")


def test_missing_feature_is_skipped_like_the_reference():
    """base_model.py:291-293: a name absent from the batch is skipped (names list still has it)."""
    g = load("deep")
    m = _cls("deep")(g["cfg_path"])
    names = sorted(m.user_feature_names | m.item_feature_names)
    batch = {k: v for k, v in g["batch"].items() if k != "category"}
    specs, dims = m._specs_for(names, batch)
    assert "category" not in [s.name for s in specs] and len(dims) == len(names) - 1


def test_cosine_schedule_matches_reference_vector():
    import numpy as np
    from tests._golden import GOLD
    from news_recsys_b200.model.model_utils.lr_schedule import CosinDecayLR
    z = np.load(f"{GOLD}/units.npz")
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-3)
    sch = CosinDecayLR(opt, lrs=[1e-3, 5e-6], milestones=[3, 9])
    got = []
    for _ in range(12):
        got.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert got == pytest.approx(z["cos_lrs"].tolist(), rel=1e-12)


def test_batch_blob_roundtrip():
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import BatchLayout
    cfg = mind_config("deep", {"user_id": 50, "item_id": 40, "category": 18, "subcategory": 70, "user_click_category": 18}, history_len=5)
    m = _cls("deep")(cfg)
    lay = BatchLayout(m, 32)
    b = synth_batch(cfg, 32, seed=1)
    blob = torch.zeros(lay.nbytes, dtype=torch.uint8)
    lay.pack(b, blob)
    v = lay.views(blob)
    for k in b:
        assert torch.equal(v[k], b[k]), k


def test_synthetic_batch_layout_matches_data_reader_contract():
    """data_reader.py:54-114: sparse int64[B]; array int64[B,L] right-padded with 0 + float mask; label float[B,2]."""
    from news_recsys_b200.synthetic import mind_config, synth_batch
    cfg = mind_config("deep", history_len=50)
    b = synth_batch(cfg, 64, seed=3)
    assert b["user_id"].dtype == torch.int64 and b["user_id"].shape == (64,)
    assert b["user_history"].shape == (64, 50) and b["user_history_mask"].dtype == torch.float32
    assert torch.all((b["user_history"] == 0) | (b["user_history_mask"] == 1))
    lens = b["user_history_mask"].sum(1).long()
    for r in range(64):
        assert torch.all(b["user_history_mask"][r, :lens[r]] == 1) and torch.all(b["user_history_mask"][r, lens[r]:] == 0)
    assert b["label"].shape == (64, 2)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py keeps stdout for the single JSON line (library banners go to stderr); the CPU arm carries the
    keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_history_filter_matches_reference_hit_rate_loop():
    """DSSM.filter_history_hits (batched tensor form) == the reference's per-user Python loop (oracle.hit_rate_filtered),
    incl. empty histories, histories holding the target, duplicates, and candidates beyond the k + |hist| window."""
    import numpy as np
    from oracle import ref_path as R
    from news_recsys_b200.model.recall.DSSM.model import DSSM
    rng = np.random.default_rng(0)
    for k, H, N in ((10, 12, 60), (3, 5, 20), (1, 1, 5)):
        Q = 200
        ranked = np.stack([rng.permutation(N)[: k + H] for _ in range(Q)])
        lens = rng.integers(0, H + 1, size=Q)
        hist = np.full((Q, H), -1, dtype=np.int64)
        for q in range(Q):
            hist[q, :lens[q]] = rng.integers(0, N, size=lens[q])     # may contain duplicates / the target
        targets = rng.integers(0, N, size=Q)
        got = DSSM.filter_history_hits(torch.from_numpy(ranked), torch.from_numpy(hist), torch.from_numpy(targets), k)
        per_user = [R.hit_rate_filtered([ranked[q]], [set(hist[q][hist[q] >= 0].tolist())], [int(targets[q])], k) for q in range(Q)]
        assert got.tolist() == [bool(x) for x in per_user]
        assert float(got.float().mean()) == pytest.approx(
            R.hit_rate_filtered(ranked, [set(h[h >= 0].tolist()) for h in hist], targets.tolist(), k))
