"""-m gpu: BASELINE.json's full-size configurations checked through size-independent properties (the oracle is
only run on sampled rows), plus empty-input edge cases."""
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cls(kind):
    import importlib
    name = {"fm": "FM", "deep": "Deep", "widedeep": "WideDeep", "dcn": "DCN", "deepfm": "DeepFM"}[kind]
    return getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_cfg1_deep_history_b1024_step_matches_oracle():
    """configs[0]: Deep, 50k users / 65k news, user_history L=50, batch 1024 — one fused step vs the oracle."""
    from news_recsys_b200.synthetic import CFG1_ROWS, mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    cfg = mind_config("deep", CFG1_ROWS, history_len=50)
    torch.manual_seed(0)
    model = _cls("deep")(cfg)
    assert model.user_input_dim + model.item_input_dim == 144
    with torch.no_grad():
        for t in model.embedding_tables.values():
            t.weight.mul_(0.3)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = synth_batch(cfg, 1024, seed=3, label_p=0.5, zipf=1.05)
    p_ref, l_ref, g_ref = R.loss_and_grads("deep", sd, cfg, batch)
    model = model.to(DEV)
    with torch.no_grad():
        p = model({k: v.to(DEV) for k, v in batch.items()})
    assert _rel(p, p_ref) < 1e-2
    tr = FusedTrainer(model, 1024, kind="deep", table_update="sparse")
    loss = float(tr.train_step(batch).item())
    assert abs(loss - float(l_ref)) < 1e-2
    # rows that received no gradient must be untouched by the sparse-row update; touched rows must move
    for name in ("user_id", "item_id"):
        k = f"embedding_tables.{name}.weight"
        touched = (g_ref[k] != 0).any(dim=1)
        new = model.state_dict()[k].cpu()
        assert torch.equal(new[~touched], sd[k][~touched])
        assert (new[touched] != sd[k][touched]).any(dim=1).float().mean() > 0.99


def test_cfg3_dcn_b65536_forward_and_step():
    """configs[2]: DCN d=112, bf16 tower, batch 65536: sampled rows vs the oracle, reproducible step."""
    from news_recsys_b200.synthetic import MIND_SMALL_ROWS, mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    cfg = mind_config("dcn", MIND_SMALL_ROWS)
    B = 65536
    batch = synth_batch(cfg, B, seed=4, label_p=0.5)
    losses, states = [], []
    for rep in range(2):
        torch.manual_seed(0)
        model = _cls("dcn")(cfg)
        with torch.no_grad():
            for t in model.embedding_tables.values():
                t.weight.mul_(0.3)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model = model.to(DEV)
        if rep == 0:
            with torch.no_grad():
                p = model({k: v.to(DEV) for k, v in batch.items()})
            idx = torch.randint(0, B, (256,))
            sub = {k: v[idx] for k, v in batch.items()}
            p_ref = R.model_forward("dcn", sd, cfg, sub, dcn_materialise=True)
            assert _rel(p[idx.to(DEV)], p_ref) < 1e-2
        tr = FusedTrainer(model, B, kind="dcn")
        losses.append(float(tr.train_step(batch).item()))
        states.append({k: v.detach().clone() for k, v in model.state_dict().items()})
    assert losses[0] == losses[1] and torch.isfinite(torch.tensor(losses[0]))
    for k in states[0]:
        assert torch.equal(states[0][k], states[1][k]), f"step not reproducible: {k}"


def test_cfg5_ten_million_row_table_gather_and_sparse_update():
    """configs[4] scale: a 10M x 32 hashed table (1.28 GB): gather parity on sampled rows, and the fused
    sparse-row AdamW changes exactly the touched rows (checksum of the rest unchanged)."""
    from news_recsys_b200 import ops
    rows, D, B = 10_000_000, 32, 16384
    g = torch.Generator(device=DEV).manual_seed(0)
    W = torch.randn(rows, D, device=DEV, generator=g)
    W[0] = 0
    ids = torch.randint(1, rows, (B,), device=DEV, generator=g)
    fb = ops.FeatBinding([ops.FeatSpec("h", "t", 0, D, 1, False, 0)], {"t": W}, {"h": ids})
    x = ops.embed_pool_fwd(fb, D)
    assert torch.equal(x, W[ids])                       # pure gather: bit-exact
    before_sum = W.double().sum()
    W0 = W[ids].clone()
    gout = torch.randn(B, D, device=DEV, generator=g)
    m, v = torch.zeros_like(W), torch.zeros_like(W)
    plan = ops.BwdPlan(fb)
    ops.embed_bwd_rowopt(plan, gout, [W] + [None] * 15, ops.L.BWD_ADAMW, lr=1e-3, step=1, weight_decay=0.01,
                         m_by_id=[m] + [None] * 15, v_by_id=[v] + [None] * 15)
    touched = torch.zeros(rows, dtype=torch.bool, device=DEV)
    touched[ids] = True
    assert int((m != 0).any(dim=1).sum()) == int(touched.sum())        # moments written only for touched rows
    delta = (W.double().sum() - before_sum).abs()
    moved = (W[ids] - W0).abs().max()
    assert 0 < float(moved) < 2.5e-3                    # one AdamW step of lr=1e-3 (+ decay) per touched row
    assert float(delta) < 1e-3 * B * D                  # nothing outside the touched rows changed


def test_empty_inputs():
    """B = 0 / Q = 0 / N = 0: every entry point returns cleanly with correctly shaped empty outputs."""
    from news_recsys_b200 import ops
    from news_recsys_b200.retrieval import TopkIndex
    W = torch.randn(10, 16, device=DEV)
    empty = torch.zeros(0, dtype=torch.long, device=DEV)
    fb = ops.FeatBinding([ops.FeatSpec("a", "t", 0, 16, 1, False, 0)], {"t": W}, {"a": empty})
    assert ops.embed_pool_fwd(fb, 16).shape == (0, 16)
    ws = [torch.randn(8, 16, device=DEV), torch.randn(1, 8, device=DEV)]
    bs = [torch.zeros(8, device=DEV), torch.zeros(1, device=DEV)]
    y, _ = ops.tower_fwd(torch.zeros(0, 16, device=DEV), ws, bs)
    assert y.shape == (0, 1)
    idx = TopkIndex(torch.randn(100, 16, device=DEV))
    s, i = idx.search(torch.zeros(0, 16, device=DEV), 5)
    assert s.shape == (0, 5) and i.shape == (0, 5)
    idx0 = TopkIndex(torch.zeros(0, 16, device=DEV))
    s, i = idx0.search(torch.randn(3, 16, device=DEV), 4)
    assert (i == -1).all() and s.shape == (3, 4)
