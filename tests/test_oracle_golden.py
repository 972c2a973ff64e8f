"""Pins the oracle restatement (oracle/ref_path.py) against vectors produced by the
reference's own modules (tests/golden/*.npz, written by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R
from tests._golden import MODEL_FIXTURES, load, opt_batches, GOLD


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_features_match_reference(name):
    g = load(name)
    cfg = g["cfg"]
    feats = cfg["features"]
    names = set(feats["user_feature_names"]) | set(feats["item_feature_names"])
    share = cfg["embeddings"].get("share_emb_table_features", {}) or {}
    x, dims, fnames = R.embeddings_from_batch(R._tables(g["sd"]), g["batch"], names,
                                              feats.get("array_feature_names", []) or [], share)
    assert dims == g["z"]["dims"].tolist()
    assert fnames == g["z"]["names"].tolist()
    assert torch.equal(x, torch.from_numpy(g["z"]["features"]))  # same ops, same order => bit-equal


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_forward_loss_grads_match_reference(name):
    g = load(name)
    prob, loss, grads = R.loss_and_grads(g["kind"], g["sd"], g["cfg"], g["batch"])
    ref_prob = torch.from_numpy(g["z"]["prob"])
    assert prob.shape == ref_prob.shape
    torch.testing.assert_close(prob, ref_prob, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(loss, torch.from_numpy(g["z"]["loss"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(prob, torch.from_numpy(g["z"]["inference"]), rtol=1e-6, atol=1e-7)
    for k, gr in g["grads"].items():
        torch.testing.assert_close(grads[k], gr, rtol=1e-5, atol=1e-8, msg=lambda m: f"{name}:{k}: {m}")
        if k.startswith("embedding_tables."):
            assert torch.count_nonzero(grads[k][0]) == 0  # padding row never receives gradient


def test_dcn_fused_association_within_tolerance():
    """x0*(xl.w)+b+xl (what the CUDA path computes) vs the reference's (x0 xl^T) w: 1e-5 rel."""
    g = load("dcn")
    a = R.model_forward("dcn", g["sd"], g["cfg"], g["batch"], dcn_materialise=True)
    b = R.model_forward("dcn", g["sd"], g["cfg"], g["batch"], dcn_materialise=False)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["fm", "fm_soft", "deep"])
def test_adamw_and_schedule_match_reference(name):
    """3 steps of the reference's AdamW + CosinDecayLR (deep/model.py:54-65) replayed with
    the restated update rule on oracle gradients."""
    g = load(name)
    z = g["z"]
    hp = g["cfg"]["train_hparams"]
    sd = {k: v.clone() for k, v in g["sd"].items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    for s, batch in enumerate(opt_batches(z)):
        lr = R.cosine_decay_lr(s, [hp["lr"], hp["min_lr"]], hp["lr_milestones"])
        assert lr == pytest.approx(float(z["opt_lrs"][s]), rel=1e-12)
        _, loss, grads = R.loss_and_grads(g["kind"], sd, g["cfg"], batch)
        torch.testing.assert_close(loss, torch.from_numpy(z[f"opt{s}_loss"]), rtol=1e-5, atol=1e-7)
        for k in sd:
            sd[k], m[k], v2[k] = R.adamw_step(sd[k], grads[k], m[k], v2[k], s + 1, lr)
    for k in sd:
        torch.testing.assert_close(sd[k], torch.from_numpy(z["sdopt__" + k]), rtol=2e-5, atol=2e-6,
                                   msg=lambda mm: f"{k}: {mm}")


def test_unit_vectors():
    z = np.load(f"{GOLD}/units.npz")
    t = lambda k: torch.from_numpy(z[k])
    # FM identity, F=2: second-order term == v1.v2 == 0.32
    assert float(z["fm_identity"].reshape(-1)[0]) == pytest.approx(0.32, rel=1e-5)
    p = torch.sigmoid(R.fm_logit(t("fm_w"), t("fm_v"), t("fm_bias")))
    torch.testing.assert_close(p, t("fm_prob"), rtol=1e-6, atol=1e-7)
    x = t("dcn_x")
    y = R.dcn_cross_v1(x, [t(f"dcn_w{i}") for i in range(3)], [t(f"dcn_b{i}") for i in range(3)])
    torch.testing.assert_close(y, t("dcn_y"), rtol=1e-6, atol=1e-6)
    y2 = R.dcn_cross_v2(x, [t(f"dcn2_W{i}") for i in range(3)], [t(f"dcn2_b{i}") for i in range(3)])
    torch.testing.assert_close(y2, t("dcn2_y"), rtol=1e-6, atol=1e-6)
    ym = R.mlp(x, [t(f"mlp_w{i}") for i in (0, 2, 4)], [t(f"mlp_b{i}") for i in (0, 2, 4)])
    torch.testing.assert_close(ym, t("mlp_y"), rtol=1e-6, atol=1e-6)
    sat = torch.nn.functional.binary_cross_entropy(torch.tensor([0.0, 1.0, 0.25]),
                                                   torch.tensor([1.0, 0.0, 1.0]), reduction="none")
    assert sat[:2].tolist() == [100.0, 100.0] == z["bce_sat"][:2].tolist()
    for s, lr in enumerate(z["cos_lrs"].tolist()):
        assert R.cosine_decay_lr(s, [1e-3, 5e-6], [3, 9]) == pytest.approx(lr, rel=1e-12)


def test_edge_cases_survey_8g():
    """Empty bag -> zeros, duplicates counted with multiplicity, id 0 counts in the denominator,
    no mask -> plain mean incl. pads (base_model.py:273-282)."""
    W = torch.randn(12, 4)
    W[0] = 0
    ids = torch.tensor([[5, 5, 9, 0], [0, 8, 0, 0], [0, 0, 0, 0]])
    mask = torch.tensor([[1., 1, 1, 0], [1, 1, 0, 0], [0, 0, 0, 0]])
    e = R.feature_embedding({"t": W}, {}, "t", ids)
    p = R.array_feature_pooling(e, mask)
    torch.testing.assert_close(p[0], (2 * W[5] + W[9]) / (3 + 1e-8))
    torch.testing.assert_close(p[1], W[8] / (2 + 1e-8))
    assert torch.count_nonzero(p[2]) == 0 and torch.isfinite(p).all()
    p2 = R.array_feature_pooling(e, None)
    torch.testing.assert_close(p2[0], (2 * W[5] + W[9]) / 4)


def test_topk_ties_lower_id_first():
    c = torch.tensor([[.5], [.9], [.9], [.1], [.9]])
    s, i = R.topk_ip(torch.ones(1, 1), c, 3)
    assert i.tolist() == [[1, 2, 4]]
    s, i = R.topk_ip(torch.ones(1, 1), c, 7)
    assert i.tolist()[0][5:] == [-1, -1]
    a = R.topk_ip(torch.ones(1, 1), c[:3], 2)
    b = R.topk_ip(torch.ones(1, 1), c[3:], 2)
    ms, mi = R.topk_merge([a[0], b[0]], [a[1], b[1] + 3], 3)
    assert mi.tolist() == [[1, 2, 4]]


def _dssm_fixture():
    z = np.load(f"{GOLD}/dssm.npz", allow_pickle=False)
    import os
    import yaml
    cfg = yaml.safe_load(open(os.path.join(GOLD, "configs", f"train_cf_{str(z['cfg'])}.yaml")))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    batch = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in__")}
    return z, cfg, sd, batch


def _dssm_side(sd, cfg, batch, order, leaf=None):
    """Tower input in the recorded set-iteration order of the reference process (DSSM/model.py:150,:167)."""
    feats = cfg["features"]
    share = cfg["embeddings"].get("share_emb_table_features", {}) or {}
    arrays = set(feats.get("array_feature_names", []) or [])
    tables = leaf if leaf is not None else R._tables(sd)
    cols = []
    for f in order:
        e = R.feature_embedding(tables, share, f, batch[f])
        if f in arrays:
            e = R.array_feature_pooling(e, batch.get(f + "_mask"))
        cols.append(e)
    return torch.cat(cols, dim=1)


def _dssm_params(sd, side):
    idx = (0, 2, 4, 6)
    return [sd[f"{side}_fc.{i}.weight"] for i in idx], [sd[f"{side}_fc.{i}.bias"] for i in idx]


def test_dssm_matches_reference():
    """tests/golden/dssm.npz was produced by the reference's own DSSM class (oracle/make_golden_dssm.py): tower inputs,
    towers, normalised outputs with in-batch negatives, InfoNCE with the label mask, and every gradient."""
    z, cfg, sd, batch = _dssm_fixture()
    t = lambda k: torch.from_numpy(z[k])
    uo, io = z["user_order"].tolist(), z["item_order"].tolist()
    ux, ix = _dssm_side(sd, cfg, batch, uo), _dssm_side(sd, cfg, batch, io)
    assert torch.equal(ux, t("user_vector")) and torch.equal(ix, t("item_vector"))
    up, ip = _dssm_params(sd, "user"), _dssm_params(sd, "item")
    torch.testing.assert_close(R.dssm_tower(ux, *up), t("user_tower"), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(R.dssm_tower(ix, *ip), t("item_tower"), rtol=1e-6, atol=1e-7)
    perms = [p for p in t("neg_perms")]
    u, it, neg = R.dssm_forward(ux, ix, up, ip, perms)
    torch.testing.assert_close(u, t("user_emb"), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(it, t("item_emb"), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(neg, t("neg_emb"), rtol=1e-6, atol=1e-7)
    mask = batch["label"][:, 1]
    torch.testing.assert_close(R.infonce_loss(u, it, neg, mask=mask), t("infonce"), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(R.infonce_loss(u, it, neg), t("infonce_nomask"), rtol=1e-6, atol=1e-7)
    # gradients through the restated path
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tabs = {k[len("embedding_tables."):-len(".weight")]: v for k, v in leaf.items() if k.startswith("embedding_tables.")}
    ux, ix = _dssm_side(sd, cfg, batch, uo, tabs), _dssm_side(sd, cfg, batch, io, tabs)
    u, it, neg = R.dssm_forward(ux, ix, _dssm_params(leaf, "user"), _dssm_params(leaf, "item"), perms)
    R.infonce_loss(u, it, neg, mask=mask).backward()
    for k in sd:
        ref = t("grad__" + k)
        got = leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(ref)
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-8, msg=lambda m: f"{k}: {m}")
