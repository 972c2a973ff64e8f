"""-m gpu: the CUDA-graph FusedTrainer step vs the oracle (loss, and parameters after optimizer steps)."""
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cls(kind):
    import importlib
    name = {"fm": "FM", "deep": "Deep", "widedeep": "WideDeep", "dcn": "DCN", "deepfm": "DeepFM", "lr": "LR"}[kind]
    return getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)


def _oracle_steps(kind, sd, cfg, batches, lr):
    """Reference semantics restricted to what the fused trainer implements: AdamW (wd 0.01) on dense
    parameters every step, and on embedding rows only when the batch touches them (lazy rows)."""
    sd = {k: v.clone() for k, v in sd.items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    losses = []
    for s, b in enumerate(batches):
        _, loss, grads = R.loss_and_grads(kind, sd, cfg, b, dcn_materialise=False) if kind == "dcn" else R.loss_and_grads(kind, sd, cfg, b)
        losses.append(float(loss))
        for k in sd:
            p, mm, vv = R.adamw_step(sd[k], grads[k], m[k], v2[k], s + 1, lr)
            if k.startswith("embedding_tables."):
                touched = (grads[k] != 0).any(dim=1, keepdim=True)
                sd[k] = torch.where(touched, p, sd[k]); m[k] = torch.where(touched, mm, m[k]); v2[k] = torch.where(touched, vv, v2[k])
            else:
                sd[k], m[k], v2[k] = p, mm, vv
    return sd, losses


@pytest.mark.parametrize("kind,hist,graph", [("fm", 0, True), ("fm", 8, True), ("lr", 0, True), ("deepfm", 0, True),
                                             ("deep", 6, True), ("widedeep", 0, True), ("dcn", 0, True), ("deepfm", 0, False)])
def test_fused_step_matches_oracle(kind, hist, graph):
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 300, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=hist)
    B = 256
    torch.manual_seed(1)
    model = _cls(kind)(cfg)
    with torch.no_grad():
        for t in model.embedding_tables.values():
            t.weight.mul_(0.2)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batches = [synth_batch(cfg, B, seed=10 + i, label_p=0.5) for i in range(3)]
    # every row id 1 is touched in step 0 only for user_id: exercises the lazy-row rule
    ref_sd, ref_losses = _oracle_steps(kind, sd, cfg, batches, cfg["train_hparams"]["lr"])
    model = model.to(DEV)
    tr = FusedTrainer(model, B, kind=kind, use_graph=graph, table_update="sparse")  # the oracle above is the lazy-row rule
    losses = [float(tr.train_step(b).item()) for b in batches]
    bf16 = kind in ("deep", "deepfm", "widedeep", "dcn")
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= (2e-2 if bf16 else 1e-5) * max(1.0, abs(b)), (losses, ref_losses)
    new_sd = model.state_dict()
    for k, ref in ref_sd.items():
        got = new_sd[k].detach().cpu()
        assert got.shape == ref.shape
        if k.startswith("embedding_tables.") and not bf16:
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-5, msg=lambda m: f"{kind}:{k}: {m}")
        else:
            # Adam's first steps move every weight by ~lr regardless of the gradient scale: compare the update
            upd, ref_upd = got - sd[k], ref - sd[k]
            denom = ref_upd.abs().max().clamp_min(1e-12)
            frac_bad = float(((upd - ref_upd).abs() > 0.5 * denom).float().mean())
            assert frac_bad < (0.05 if bf16 else 1e-3), f"{kind}:{k}: {frac_bad:.4f} of the updates differ"


def test_graph_replay_is_deterministic():
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 300, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("deepfm", rows)
    outs = []
    for _ in range(2):
        torch.manual_seed(3)
        model = _cls("deepfm")(cfg).to(DEV)
        tr = FusedTrainer(model, 256, kind="deepfm")
        for i in range(4):
            tr.train_step(synth_batch(cfg, 256, seed=50 + i, label_p=0.5))
        outs.append({k: v.detach().clone() for k, v in model.state_dict().items()})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("impl", ["split", "flat"])
@pytest.mark.parametrize("name", ["fm", "fm_soft", "deep"])
def test_dense_mode_matches_reference_adamw_golden(name, impl):
    """table_update="dense" is the reference's optimizer (dense AdamW, wd 0.01, every row every step): three
    fused steps on the fixture batches land on the parameters the REFERENCE itself produced with its own
    torch.optim.AdamW + CosinDecayLR (tests/golden/{fm,deep}.npz `sdopt__*`, oracle/make_golden.py)."""
    from tests._golden import load, opt_batches
    from news_recsys_b200.trainer import FusedTrainer
    g = load(name)
    z = g["z"]
    batches = opt_batches(z)
    B = batches[0]["label"].shape[0]
    model = _cls(g["kind"])(g["cfg_path"])
    model.load_state_dict(g["sd"], strict=True)
    model = model.to(DEV)
    tr = FusedTrainer(model, B, kind=g["kind"], table_update="dense", dense_impl=impl)
    bf16 = g["kind"] != "fm"
    for s, b in enumerate(batches):
        loss = float(tr.train_step(b).item())
        ref = float(z[f"opt{s}_loss"])
        assert abs(loss - ref) <= (2e-2 if bf16 else 1e-5) * max(1.0, abs(ref)), (s, loss, ref)
    new_sd = model.state_dict()
    for k, v0 in g["sd"].items():
        got, ref = new_sd[k].detach().cpu(), torch.from_numpy(z["sdopt__" + k])
        if not bf16:
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-5, msg=lambda m: f"{name}:{k}: {m}")
        else:
            upd, ref_upd = got - v0, ref - v0
            denom = ref_upd.abs().max().clamp_min(1e-12)
            frac_bad = float(((upd - ref_upd).abs() > 0.5 * denom).float().mean())
            assert frac_bad < 0.05, f"{name}:{k}: {frac_bad:.4f} of the updates differ"


def test_dense_and_sparse_agree_on_touched_rows_first_step():
    """After ONE step both modes moved the touched rows identically; dense additionally decays untouched rows."""
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 3000, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("fm", rows)
    b = synth_batch(cfg, 256, seed=5, label_p=0.5)
    out = {}
    for mode in ("sparse", "dense"):
        torch.manual_seed(2)
        model = _cls("fm")(cfg).to(DEV)
        w0 = model.embedding_tables["user_id"].weight.detach().clone()
        FusedTrainer(model, 256, kind="fm", table_update=mode).train_step(b)
        out[mode] = model.embedding_tables["user_id"].weight.detach().clone()
    touched = torch.zeros(3000, dtype=torch.bool, device=DEV)
    touched[b["user_id"].view(-1).to(DEV)] = True
    touched[0] = False
    torch.testing.assert_close(out["dense"][touched], out["sparse"][touched], rtol=1e-6, atol=1e-7)
    lr = cfg["train_hparams"]["lr"]
    un = ~touched
    un[0] = False
    assert torch.equal(out["sparse"][un], w0[un])
    torch.testing.assert_close(out["dense"][un], w0[un] * (1 - lr * 0.01), rtol=1e-6, atol=1e-9)


def test_pipelined_feed_equals_step_by_step():
    """feed(pinned blob) (H2D on the copy stream, loss read one step late) == load_batch + step."""
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 300, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("deepfm", rows)
    batches = [synth_batch(cfg, 256, seed=70 + i, label_p=0.5) for i in range(5)]
    torch.manual_seed(4)
    tr = FusedTrainer(_cls("deepfm")(cfg).to(DEV), 256, kind="deepfm")
    want = [float(tr.train_step(b).item()) for b in batches]
    torch.manual_seed(4)
    tr2 = FusedTrainer(_cls("deepfm")(cfg).to(DEV), 256, kind="deepfm")
    blobs = []
    for b in batches:
        hb = torch.empty(tr2.layout.nbytes, dtype=torch.uint8).pin_memory()
        tr2.layout.pack(b, hb)
        blobs.append(hb)
    got = [tr2.feed(hb) for hb in blobs]
    assert got[0] is None
    got = got[1:] + [tr2.drain()]
    assert got == want
    assert tr2.drain() is None
    for k, v in tr.model.state_dict().items():
        assert torch.equal(v, tr2.model.state_dict()[k]), k


@pytest.mark.parametrize("kind,hist", [("fm", 0), ("deepfm", 0), ("deep", 6)])
def test_dense_split_equals_dense_flat_bitwise(kind, hist):
    """The two implementations of the dense semantics — (fused row AdamW on touched rows + g = 0 sweep of the rest) and
    (dense table gradients + one AdamW over everything) — produce the same bits, tables and tower, after 4 steps."""
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 3000, "item_id": 2000, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=hist)
    batches = [synth_batch(cfg, 256, seed=90 + i, label_p=0.5) for i in range(4)]
    out = {}
    for impl in ("split", "flat"):
        torch.manual_seed(6)
        model = _cls(kind)(cfg).to(DEV)
        tr = FusedTrainer(model, 256, kind=kind, table_update="dense", dense_impl=impl)
        losses = [float(tr.train_step(b).item()) for b in batches]
        out[impl] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()})
    assert out["split"][0] == out["flat"][0]
    for k, v in out["flat"][1].items():
        assert torch.equal(out["split"][1][k], v), k


def test_int32_ids_equal_int64_ids():
    """SURVEY §8 f1 (compact batch ingestion): the same step fed int32 ids (half the id bytes per batch) == int64 ids."""
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 3000, "item_id": 2000, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("deep", rows, history_len=6)
    batches = [synth_batch(cfg, 256, seed=30 + i, label_p=0.5) for i in range(3)]
    out = {}
    for idt in (torch.int64, torch.int32):
        torch.manual_seed(8)
        model = _cls("deep")(cfg).to(DEV)
        tr = FusedTrainer(model, 256, kind="deep", id_dtype=idt)
        losses = [float(tr.train_step(b).item()) for b in batches]
        out[idt] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()}, tr.layout.nbytes)
    assert out[torch.int32][2] < out[torch.int64][2]
    assert out[torch.int32][0] == out[torch.int64][0]
    for k, v in out[torch.int64][1].items():
        assert torch.equal(out[torch.int32][1][k], v), k


@pytest.mark.parametrize("kind", ["deep", "fm"])
def test_label_column_count_is_free(kind):
    """the reference accepts any number of label columns and its loss reads column 0 (deep/model.py:69);
    the blob layout used to hard-code (B, 2).  One-column and three-column labels train exactly like two-column ones."""
    from news_recsys_b200._lib import NrxError
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 3000, "item_id": 2000, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=6 if kind == "deep" else 0)
    batches = [synth_batch(cfg, 256, seed=40 + i, label_p=0.5) for i in range(3)]
    out = {}
    for nl in (2, 1, 3):
        torch.manual_seed(8)
        model = _cls(kind)(cfg).to(DEV)
        tr = FusedTrainer(model, 256, kind=kind, n_labels=nl)
        losses = []
        for b in batches:
            b = dict(b)
            col0 = b["label"][:, :1]
            b["label"] = col0 if nl == 1 else torch.cat([col0] + [torch.rand(256, 1)] * (nl - 1), dim=1)
            losses.append(float(tr.train_step(b).item()))
        out[nl] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()})
    for nl in (1, 3):
        assert out[nl][0] == out[2][0]
        for k, v in out[2][1].items():
            assert torch.equal(out[nl][1][k], v), (nl, k)
    tr = FusedTrainer(_cls(kind)(cfg).to(DEV), 256, kind=kind)          # default n_labels = 2
    bad = dict(batches[0]); bad["label"] = bad["label"][:, 0]
    with pytest.raises(NrxError, match="n_labels"):
        tr.train_step(bad)
