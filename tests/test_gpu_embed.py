"""-m gpu parity tests of K1 (gather+pool+concat), K2 (field logits / fused FM / BCE) and
K3 (sorted segment-reduce backward, dense + fused row update) against the oracle and the
reference-produced golden vectors.  Everything goes through the C ABI (libnrx.so)."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R
from tests._golden import load

pytestmark = pytest.mark.gpu

DEV = "cuda"
FP32_RTOL = 1e-5  # north_star: pooled embeddings, logits and gradients within 1e-5 relative in fp32


def _cuda_batch(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def _build(kind, cfg_path, sd):
    from news_recsys_b200.model.sort.fm.model import FM
    from news_recsys_b200.model.sort.lr.model import LR
    cls = {"fm": FM, "lr": LR}[kind]
    m = cls(cfg_path)
    m.load_state_dict(sd, strict=True)  # reference checkpoint keys load unchanged
    return m.to(DEV)


def _rel_close(a, b, rtol, what):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = b.abs().max().clamp_min(1e-30)
    err = (a - b).abs().max() / scale
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol}"


@pytest.mark.parametrize("name", ["lr", "fm", "fm_hist", "fm_soft", "fm_hist_soft", "deep", "deep_hist", "deep_hist_nomask", "widedeep_hist", "dcn_hist"])
def test_k1_features_match_reference(name):
    """get_embeddings_from_batch: sparse columns bit-exact (pure gather), pooled columns 1e-5."""
    from news_recsys_b200.model.BaseModel.base_model import BaseModel
    g = load(name)
    m = BaseModel(g["cfg_path"])
    m.load_state_dict({k: v for k, v in g["sd"].items() if k.startswith("embedding_tables.")}, strict=True)
    m = m.to(DEV)
    with torch.no_grad():
        x, dims, names = m.get_embeddings_from_batch(_cuda_batch(g["batch"]), m.user_feature_names | m.item_feature_names)
    ref = torch.from_numpy(g["z"]["features"])
    assert dims == g["z"]["dims"].tolist() and names == g["z"]["names"].tolist()
    assert x.shape == ref.shape and x.dtype == torch.float32
    col = 0
    for d, n in zip(dims, names):
        a, b = x[:, col:col + d].cpu(), ref[:, col:col + d]
        if n in m.array_feature_names:
            torch.testing.assert_close(a, b, rtol=FP32_RTOL, atol=1e-7)
        else:
            assert torch.equal(a, b), f"sparse feature {n} must be bit-exact"
        col += d


def test_k1_edge_cases():
    """duplicates counted, id 0 counts in the denominator, empty bag -> exact zeros,
    no mask -> plain mean incl. pads, int32 ids, D in {1,16,17,32}."""
    from news_recsys_b200 import ops
    torch.manual_seed(0)
    for D in (1, 16, 17, 32, 40):
        W = torch.randn(50, D)
        W[0] = 0
        ids = torch.tensor([[5, 5, 9, 0, 0], [0, 8, 0, 0, 0], [0, 0, 0, 0, 0], [1, 2, 3, 4, 49]])
        mask = torch.tensor([[1., 1, 1, 0, 0], [1, 1, 0, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]])
        for idt in (torch.int64, torch.int32):
            for use_mask in (True, False):
                batch = {"h": ids.to(idt).to(DEV)}
                if use_mask:
                    batch["h_mask"] = mask.to(DEV)
                spec = ops.FeatSpec("h", "t", 0, D, 5, True, 0)
                fb = ops.FeatBinding([spec], {"t": W.to(DEV)}, batch)
                out = ops.embed_pool_fwd(fb, D).cpu()
                ref = R.array_feature_pooling(R.feature_embedding({"t": W}, {"h": "t"}, "h", ids), mask if use_mask else None)
                torch.testing.assert_close(out, ref, rtol=FP32_RTOL, atol=1e-7)
                if use_mask:
                    assert torch.count_nonzero(out[2]) == 0 and torch.isfinite(out).all()


def test_k1_out_of_range_id_sets_status():
    from news_recsys_b200 import ops
    W = torch.randn(10, 16, device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    fb = ops.FeatBinding([ops.FeatSpec("a", "t", 0, 16, 1, False, 0)], {"t": W}, {"a": torch.tensor([1, 10, 3], device=DEV)})
    ops.embed_pool_fwd(fb, 16, status=st)
    assert int(st.item()) != 0


def _random_problem(B, rows, D, L, seed, zipf=False):
    g = torch.Generator().manual_seed(seed)
    tables = {"u": torch.randn(rows[0], D[0], generator=g), "i": torch.randn(rows[1], D[1], generator=g),
              "c": torch.randn(rows[2], D[2], generator=g)}
    for t in tables.values():
        t[0] = 0
    def draw(n, shape):
        if zipf:
            x = torch.from_numpy(np.random.default_rng(seed).zipf(1.2, size=shape).astype(np.int64))
            return x.clamp_(max=n - 1)
        return torch.randint(0, n, shape, generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).float()
    batch = {"u": draw(rows[0], (B,)), "i": draw(rows[1], (B,)), "c": draw(rows[2], (B,)),
             "h": draw(rows[1], (B, L)) * mask.long(), "h_mask": mask}
    return tables, batch


def _specs(D, L):
    from news_recsys_b200 import ops
    # sorted-name column order: c | h | i | u ; h shares table "i"
    cols = [0, D[2], D[2] + D[1], D[2] + 2 * D[1]]
    return [ops.FeatSpec("c", "c", 0, D[2], 1, False, cols[0]), ops.FeatSpec("h", "i", 1, D[1], L, True, cols[1]),
            ops.FeatSpec("i", "i", 1, D[1], 1, False, cols[2]), ops.FeatSpec("u", "u", 2, D[0], 1, False, cols[3])], cols[3] + D[0]


def _oracle_grads(tables, batch, gout):
    leaf = {k: v.clone().requires_grad_(True) for k, v in tables.items()}
    x, _, _ = R.embeddings_from_batch(leaf, batch, ["c", "h", "i", "u"], ["h"], {"h": "i"})
    (x * gout).sum().backward()
    return x.detach(), {k: v.grad for k, v in leaf.items()}


@pytest.mark.parametrize("B,rows,D,L,zipf", [
    (64, (40, 30, 7), (32, 32, 16), 12, False),
    (1000, (500, 300, 18), (32, 16, 17), 50, False),      # scalar path (D=17), long runs on an 18-row table
    (4096, (5000, 3000, 18), (32, 32, 16), 50, True),     # zipf: runs spanning many tiles
    (3, (5, 4, 3), (1, 1, 1), 1, False),
    (777, (100, 60, 20), (64, 48, 8), 7, False),          # dims > 32 (NC = 2); 7-chunk merge plan
    (1024, (90000, 60000, 18), (16, 16, 16), 15, False),  # table i: 1024 + 15360 = 16 full chunks exactly
    (2048, (5000, 3000, 2), (16, 32, 16), 12, True),      # zipf + 26 chunks with one ragged; a 2-row table (1 key bit)
    (16384, (94058, 65239, 270), (16, 16, 16), 1, False), # the bench shape: 32 chunks on the shared table
])
def test_k3_dense_backward_matches_autograd(B, rows, D, L, zipf):
    from news_recsys_b200 import ops
    tables, batch = _random_problem(B, rows, D, L, seed=B, zipf=zipf)
    specs, out_dim = _specs(D, L)
    gout = torch.randn(B, out_dim, generator=torch.Generator().manual_seed(5))
    x_ref, g_ref = _oracle_grads(tables, batch, gout)
    dt = {k: v.to(DEV) for k, v in tables.items()}
    fb = ops.FeatBinding(specs, dt, _cuda_batch(batch))
    x = ops.embed_pool_fwd(fb, out_dim)
    _rel_close(x, x_ref, FP32_RTOL, "features")
    plan = ops.BwdPlan(fb)
    by_id = [dt["c"], dt["i"], dt["u"]] + [None] * 13
    grads = ops.embed_bwd_dense(plan, gout.to(DEV), by_id)
    for tid, k in enumerate(["c", "i", "u"]):
        _rel_close(grads[tid], g_ref[k], FP32_RTOL, f"grad[{k}]")
        assert torch.count_nonzero(grads[tid][0]) == 0, "padding row must not receive gradient"
    # bitwise reproducible (fixed reduction order, no float atomics)
    grads2 = ops.embed_bwd_dense(ops.BwdPlan(fb), gout.to(DEV), by_id)
    for a, b in zip(grads[:3], grads2[:3]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["sgd", "adamw"])
def test_k3_fused_row_update(mode):
    """Fused sparse row update == the optimizer rule applied to the rows the batch touched."""
    from news_recsys_b200 import ops
    B, rows, D, L = 512, (200, 150, 18), (32, 32, 16), 20
    tables, batch = _random_problem(B, rows, D, L, seed=9)
    specs, out_dim = _specs(D, L)
    gout = torch.randn(B, out_dim, generator=torch.Generator().manual_seed(6)) * 0.1
    _, g_ref = _oracle_grads(tables, batch, gout)
    dt = {k: v.clone().to(DEV) for k, v in tables.items()}
    fb = ops.FeatBinding(specs, dt, _cuda_batch(batch))
    ops.embed_pool_fwd(fb, out_dim)  # fills inv_den
    plan = ops.BwdPlan(fb)
    by_id = [dt["c"], dt["i"], dt["u"]] + [None] * 13
    lr, wd = 0.05, 0.01
    if mode == "sgd":
        ops.embed_bwd_rowopt(plan, gout.to(DEV), by_id, ops.L.BWD_SGD, lr=lr, weight_decay=wd)
    else:
        m = [torch.zeros_like(t) if t is not None else None for t in by_id]
        v = [torch.zeros_like(t) if t is not None else None for t in by_id]
        ops.embed_bwd_rowopt(plan, gout.to(DEV), by_id, ops.L.BWD_ADAMW, lr=lr, step=1, weight_decay=wd, m_by_id=m, v_by_id=v)
    for k in ["c", "i", "u"]:
        W, G = tables[k], g_ref[k]
        touched = (G != 0).any(dim=1)
        if mode == "sgd":
            exp = W - lr * (G + wd * W)
        else:
            exp, _, _ = R.adamw_step(W, G, torch.zeros_like(W), torch.zeros_like(W), 1, lr, wd=wd)
        exp = torch.where(touched[:, None], exp, W)  # untouched rows are left alone (lazy update)
        torch.testing.assert_close(dt[k].cpu(), exp, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["lr", "fm", "fm_hist", "fm_soft", "fm_hist_soft"])
def test_model_forward_backward_matches_reference(name):
    """forward(batch), bceLoss and loss.backward() of the drop-in modules vs the reference's own outputs."""
    g = load(name)
    m = _build(g["kind"], g["cfg_path"], g["sd"])
    batch = _cuda_batch(g["batch"])
    prob = m(batch)
    ref_prob = torch.from_numpy(g["z"]["prob"])
    assert prob.shape == ref_prob.shape
    torch.testing.assert_close(prob.detach().cpu(), ref_prob, rtol=1e-5, atol=1e-6)
    loss = m.bceLoss(prob, batch["label"][:, 0])
    torch.testing.assert_close(loss.detach().cpu(), torch.from_numpy(g["z"]["loss"]), rtol=1e-5, atol=1e-6)
    loss.backward()
    for k, gr in g["grads"].items():
        p = dict(m.named_parameters())[k]
        assert p.grad is not None, k
        _rel_close(p.grad, gr, 2e-5, f"{name}:{k}")
    with torch.no_grad():
        torch.testing.assert_close(m.inference(batch).cpu(), torch.from_numpy(g["z"]["inference"]), rtol=1e-5, atol=1e-6)


def test_fm_fused_large_batch_properties():
    """BASELINE config 2 shape (MIND-small rows, D=16, B=16384): fused kernel == unfused K1+field_logit path,
    and FM identity: with all but two fields on the padding row the logit is bias + w1 + w2 + v1.v2."""
    from news_recsys_b200 import ops
    torch.manual_seed(1)
    rows = {"user_id": 94058, "item_id": 65239, "category": 18, "subcategory": 270, "user_click_category": 18}
    names = sorted(rows)
    B, D = 16384, 16
    tables = {k: torch.randn(n, D, device=DEV) * 0.3 for k, n in rows.items()}
    for t in tables.values():
        t[0] = 0
    batch = {k: torch.randint(1, n, (B,), device=DEV) for k, n in rows.items()}
    specs = [ops.FeatSpec(k, k, i, D, 1, False, i * D) for i, k in enumerate(names)]
    bias = torch.tensor([0.25], device=DEV)
    fb = ops.FeatBinding(specs, tables, batch)
    label = (torch.rand(B, device=DEV) < 0.5).float()
    prob, loss, dl, logit = ops.fm_fused_fwd(fb, bias, label, want_logit=True)
    x = ops.embed_pool_fwd(fb, 5 * D)
    logit2 = ops.field_logit_fwd(x, [i * D for i in range(5)], [D] * 5, ops.L.FIELD_FM)
    torch.testing.assert_close(logit, logit2 + 0.25, rtol=1e-5, atol=1e-5)
    xr = x.cpu()
    w, v = R.fm_split(xr, [D] * 5)
    ref = R.fm_logit(w, v, torch.tensor([0.25])).view(-1)
    torch.testing.assert_close(logit.cpu(), ref, rtol=1e-5, atol=1e-5)
    p_ref = torch.sigmoid(ref)
    torch.testing.assert_close(prob.cpu(), p_ref, rtol=1e-5, atol=1e-6)
    l_ref = torch.nn.functional.binary_cross_entropy(p_ref, label.cpu(), reduction="none")
    torch.testing.assert_close(loss.cpu(), l_ref, rtol=1e-4, atol=1e-5)
    gx = ops.fm_fused_bwd(fb, dl, 5 * D)
    gx2 = torch.zeros_like(x)
    ops.field_logit_bwd(x, [i * D for i in range(5)], [D] * 5, ops.L.FIELD_FM, dl, gx2, accumulate=False)
    torch.testing.assert_close(gx, gx2, rtol=1e-5, atol=1e-7)
    # identity
    b2 = {k: torch.zeros(4, dtype=torch.long, device=DEV) for k in rows}
    b2["user_id"] = torch.tensor([3, 4, 5, 6], device=DEV)
    b2["item_id"] = torch.tensor([7, 8, 9, 10], device=DEV)
    fb2 = ops.FeatBinding(specs, tables, b2)
    _, _, _, lg = ops.fm_fused_fwd(fb2, bias, None, want_logit=True)
    u, it = tables["user_id"][b2["user_id"]], tables["item_id"][b2["item_id"]]
    exp = 0.25 + u[:, 0] + it[:, 0] + (u[:, 1:] * it[:, 1:]).sum(1)
    torch.testing.assert_close(lg, exp, rtol=1e-5, atol=1e-6)


def test_dense_adamw_matches_reference_rule():
    from news_recsys_b200 import ops
    torch.manual_seed(2)
    n = 100003
    p, g = torch.randn(n), torch.randn(n) * 0.1
    m, v = torch.zeros(n), torch.zeros(n)
    dp, dm, dv = p.to(DEV), m.to(DEV), v.to(DEV)
    for step in (1, 2, 3):
        ops.adamw_dense_(dp, g.to(DEV), dm, dv, step, 1e-3)
        p, m, v = R.adamw_step(p, g, m, v, step, 1e-3)
    torch.testing.assert_close(dp.cpu(), p, rtol=1e-5, atol=1e-7)
    opt_p = torch.nn.Parameter(torch.randn(257))
    ref = opt_p.detach().clone()
    opt = torch.optim.AdamW([opt_p], lr=1e-3)
    gg = torch.randn(257)
    opt_p.grad = gg.clone()
    opt.step()
    d = ref.to(DEV)
    ops.adamw_dense_(d, gg.to(DEV), torch.zeros(257, device=DEV), torch.zeros(257, device=DEV), 1, 1e-3)
    torch.testing.assert_close(d.cpu(), opt_p.detach(), rtol=1e-5, atol=1e-7)


def test_model_and_trainer_raise_on_out_of_table_ids():
    """an id outside its table must not train silently on zero vectors.  The reference's nn.Embedding raises
    (base_model.py:271); here K1 raises a sticky status bit that Model.check_ids() / FusedTrainer.check_status() /
    FusedTrainer.feed() turn into an exception."""
    from news_recsys_b200._lib import NrxError
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 300, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("deep", rows)
    torch.manual_seed(0)
    m = Deep(cfg).to(DEV)
    good = {k: v.to(DEV) for k, v in synth_batch(cfg, 64, seed=1).items()}
    with torch.no_grad():
        m(good)
    m.check_ids()                                   # clean batch: no exception
    bad = dict(good)
    bad["item_id"] = good["item_id"].clone()
    bad["item_id"][5] = 200                         # == rows: one past the table
    with torch.no_grad():
        m(bad)
    with pytest.raises(IndexError):
        m.check_ids()
    m.check_ids()                                   # the flag was consumed
    tr = FusedTrainer(m, 64, kind="deep")
    tr.train_step({k: v.cpu() for k, v in good.items()})
    tr.check_status()
    tr.train_step({k: v.cpu() for k, v in bad.items()})
    with pytest.raises(NrxError, match="outside its embedding table"):
        tr.check_status()
