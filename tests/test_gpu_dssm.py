"""-m gpu: DSSM towers / InfoNCE / retrieval vs the oracle restatement (parity unpinned: the reference's
recall/DSSM/model.py is not importable and faiss is not vendored — see oracle/ref_path.py header)."""
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cfg():
    from news_recsys_b200.synthetic import mind_config
    rows = {"user_id": 400, "item_id": 300, "category": 18, "subcategory": 70, "user_click_category": 18}
    return mind_config("deep", rows, history_len=10)


def _params(seq):
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    return [m.weight.detach().cpu() for m in lin], [m.bias.detach().cpu() for m in lin]


@pytest.mark.parametrize("out_dim", [16, 128])
def test_dssm_forward_loss_and_retrieval(out_dim):
    from news_recsys_b200.model.recall.DSSM.model import DSSM
    from news_recsys_b200.synthetic import synth_batch
    cfg = _cfg()
    torch.manual_seed(0)
    m = DSSM(cfg, hparams={"out_dim": out_dim, "negative_sample_rate": 3})
    with torch.no_grad():
        for t in m.embedding_tables.values():
            t.weight.mul_(0.3)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    B = 200
    batch = synth_batch(cfg, B, seed=5, label_p=0.7)
    perms = [torch.randperm(B, generator=torch.Generator().manual_seed(i)) for i in range(3)]
    # oracle
    tables = R._tables(sd)
    share = cfg["embeddings"]["share_emb_table_features"]
    arr = cfg["features"]["array_feature_names"]
    ux, _, _ = R.embeddings_from_batch(tables, batch, cfg["features"]["user_feature_names"], arr, share)
    ix, _, _ = R.embeddings_from_batch(tables, batch, cfg["features"]["item_feature_names"], arr, share)
    u_ref, i_ref, n_ref = R.dssm_forward(ux, ix, _params(m.user_fc), _params(m.item_fc), perms)
    l_ref = R.infonce_loss(u_ref, i_ref, n_ref, mask=batch["label"][:, 1])
    m = m.to(DEV)
    db = {k: v.to(DEV) for k, v in batch.items()}
    u, it, neg = m(db, perms)
    assert u.shape == (B, out_dim) and neg.shape == (B, 3, out_dim)
    torch.testing.assert_close(u.detach().cpu(), u_ref, rtol=2e-2, atol=2e-2)   # bf16 towers, unit vectors
    torch.testing.assert_close(it.detach().cpu(), i_ref, rtol=2e-2, atol=2e-2)
    loss = m.training_step(db, neg_perms=perms)
    torch.testing.assert_close(loss.detach().cpu(), l_ref, rtol=3e-2, atol=3e-2)
    loss.backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # retrieval over the item corpus == oracle top-k on the SAME item vectors
    m.build_item_index([db])
    s, ids = m.retrieve(db, 10)
    ref_s, ref_i = R.topk_ip(torch.nn.functional.normalize(m.user_tower(db).detach().cpu(), dim=1),
                             m.all_item_embeddings.cpu(), 10)
    torch.testing.assert_close(s.cpu(), ref_s, rtol=1e-4, atol=1e-5)
    assert (ids.cpu() == ref_i).float().mean() > 0.98  # queries re-normalised on the GPU: only near-ties may move
