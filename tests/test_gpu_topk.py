"""-m gpu: exact inner-product top-k (K6) vs the oracle — bit-exact ids, ties to the lower id."""
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check(Qn, N, D, k, seed, normalize=True, dup=False):
    from news_recsys_b200.retrieval import TopkIndex
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(N, D, generator=g)
    q = torch.randn(Qn, D, generator=g)
    if normalize:
        c = torch.nn.functional.normalize(c, dim=1)
        q = torch.nn.functional.normalize(q, dim=1)
    if dup and N >= 64:  # exact duplicates -> exact score ties -> lower id first
        c[N // 2: N // 2 + 20] = c[3: 23]
        c[N - 5:] = c[0]
    ref_s, ref_i = R.topk_ip(q, c, k)
    idx = TopkIndex(c.to(DEV))
    s, i, st = idx.search(q.to(DEV), k, want_status=True)
    assert torch.equal(i.cpu(), ref_i), f"ids differ (Q={Qn}, N={N}, D={D}, k={k})"
    torch.testing.assert_close(s.cpu(), ref_s, rtol=1e-6, atol=1e-7)
    return int(st.sum().item())


@pytest.mark.parametrize("Qn,N,D,k", [
    (1, 10, 16, 3),            # tiny corpus -> exact kernel
    (5, 3883, 16, 40),         # the reference's own scale (recall/DSSM: ~3.9k movies, D=16, k=10+|hist|)
    (3, 50, 16, 100),          # k > N: -1 / -FLT_MAX padding
    (300, 50000, 128, 100),    # tensor path, several query tiles
    (64, 20000, 72, 10),       # D not a multiple of 16
    (17, 100000, 128, 100),
    (2, 70000, 200, 5),        # Dp = 208 (2-stage pipeline)
])
def test_topk_matches_oracle(Qn, N, D, k):
    _check(Qn, N, D, k, seed=N + Qn)


def test_topk_ties_lower_id_first():
    fb = _check(40, 40000, 64, 50, seed=1, dup=True)
    assert fb >= 0


def test_topk_unnormalised_and_fallback_counts():
    """Unnormalised N(0,1) vectors: wider score spread, the bound still holds; report the fallback rate."""
    fb = _check(128, 60000, 128, 100, seed=2, normalize=False)
    assert fb <= 128


def test_topk_adversarial_near_ties_use_fallback():
    """A corpus of near-identical rows defeats the bf16 filter; the completeness check must route those
    queries to the exact scan and still return the oracle's ids."""
    from news_recsys_b200.retrieval import TopkIndex
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 64, generator=g)
    c = base + 1e-4 * torch.randn(30000, 64, generator=g)
    q = base + 0.1 * torch.randn(4, 64, generator=g)
    ref_s, ref_i = R.topk_ip(q, c, 20)
    s, i, st = TopkIndex(c.to(DEV)).search(q.to(DEV), 20, want_status=True)
    assert torch.equal(i.cpu(), ref_i)
    assert int(st.sum()) > 0, "expected the exact fallback to serve the near-tie queries"


def test_topk_merge_equals_single_index():
    """Corpus sharded 4 ways (id_base per shard) + merge == one index."""
    from news_recsys_b200.retrieval import TopkIndex, topk_merge
    g = torch.Generator().manual_seed(4)
    c = torch.nn.functional.normalize(torch.randn(40000, 128, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(33, 128, generator=g), dim=1)
    c[39990:] = c[5]  # ties across shards
    ref_s, ref_i = R.topk_ip(q, c, 100)
    parts = []
    for r in range(4):
        sh = c[r * 10000:(r + 1) * 10000].to(DEV)
        parts.append(TopkIndex(sh, id_base=r * 10000).search(q.to(DEV), 100, want_scores64=True))
    for key in (2, 0):   # fp64 ordering keys (what ShardedTopk exchanges) and the fp32 scores
        s = torch.stack([p[key] for p in parts])
        i = torch.stack([p[1] for p in parts])
        ms, mi = topk_merge(s, i)
        assert torch.equal(mi.cpu(), ref_i)
        torch.testing.assert_close(ms.cpu(), ref_s, rtol=1e-6, atol=1e-7)


def test_topk_sharded_near_ties_at_1m_rows():
    """BASELINE cfg4 size (N = 1M, D = 128, k = 100), corpus in 8 shards, with pairs of rows in DIFFERENT shards whose fp64
    scores differ by less than one fp32 ulp and whose id order is the opposite of their score order.  The merged result
    must equal the single index and the oracle (the fp32 merge ordered such pairs by id)."""
    from news_recsys_b200.retrieval import TopkIndex, topk_merge
    g = torch.Generator().manual_seed(9)
    N, D, k, Qn, G = 1_000_000, 128, 100, 6, 8
    c = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(Qn, D, generator=g), dim=1)
    for j in range(Qn):   # row lo (shard 0) ~ the query; row hi (last shard) = the same row nudged UP by one ulp along q
        lo, hi = 1000 + j, N - 1000 - j
        c[lo] = torch.nn.functional.normalize(q[j] + 0.05 * torch.randn(D, generator=g), dim=0)
        c[hi] = c[lo]
        e = int(q[j].abs().argmax())
        c[hi, e] = torch.nextafter(c[lo, e], c[lo, e] + torch.sign(q[j, e]))
    ties = 0
    for j in range(Qn):
        s64 = q[j].double() @ c[[1000 + j, N - 1000 - j]].double().T
        assert 0 < float(s64[1] - s64[0]) < 6e-8
        ties += int(float(s64[1].float()) == float(s64[0].float()))
    assert ties >= 3, "fixture: the pairs must round to the same fp32 score"
    ref_s, ref_i = R.topk_ip(q, c, k)
    for j in range(Qn):
        assert ref_i[j, 0] == N - 1000 - j and ref_i[j, 1] == 1000 + j   # higher id first: score order, not id order
    qd = q.to(DEV)
    s1, i1 = TopkIndex(c.to(DEV)).search(qd, k)
    assert torch.equal(i1.cpu(), ref_i)
    parts = []
    for r in range(G):
        lo, hi = r * N // G, (r + 1) * N // G
        parts.append(TopkIndex(c[lo:hi].to(DEV), id_base=lo).search(qd, k, want_scores64=True))
    ms, mi = topk_merge(torch.stack([p[2] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi.cpu(), ref_i), "sharded merge differs from the single index"
    torch.testing.assert_close(ms.cpu(), ref_s, rtol=1e-6, atol=1e-7)


def test_topk_clustered_corpus_tiled_fallback():
    """A clustered corpus (every row within 1e-3 of one of 8 centres) makes the bf16 filter useless for most queries:
    they land on the fallback list and the corpus-sliced exact scan + merge must still return the oracle's lists."""
    from news_recsys_b200.retrieval import TopkIndex
    g = torch.Generator().manual_seed(11)
    centres = torch.nn.functional.normalize(torch.randn(8, 64, generator=g), dim=1)
    c = centres[torch.randint(0, 8, (60000,), generator=g)] + 1e-3 * torch.randn(60000, 64, generator=g)
    q = centres[:5] + 0.05 * torch.randn(5, 64, generator=g)
    ref_s, ref_i = R.topk_ip(q, c, 30)
    s, i, st = TopkIndex(c.to(DEV)).search(q.to(DEV), 30, want_status=True)
    assert torch.equal(i.cpu(), ref_i)
    torch.testing.assert_close(s.cpu(), ref_s, rtol=1e-6, atol=1e-7)
    assert int(st.sum()) >= 1


def test_topk_searcher_api():
    from news_recsys_b200.model.model_utils.TopKSearcher import TopKSearcher
    torch.manual_seed(0)
    emb = torch.nn.Embedding(5000, 32)
    ts = TopKSearcher(k=10)
    ts.update_embedding(emb, normalize=True)
    qs = [torch.randn(32) for _ in range(7)]
    ids, scores = ts.search(qs, normalize=True)
    ref_s, ref_i = R.topk_ip(torch.stack(qs), emb.weight.detach(), 10, normalize=True)
    assert len(ids) == 7 and len(ids[0]) == 10 and isinstance(ids[0][0], int)
    # normalisation is done on the GPU in fp32, so only near-ties may differ: compare score values
    torch.testing.assert_close(torch.tensor(scores), ref_s, rtol=1e-5, atol=1e-6)
    assert ts.search([], normalize=True) == ([], [])


@pytest.mark.parametrize("Q", [7, 300])
def test_split_and_fused_final_agree_bitwise(Q, monkeypatch):
    """The wide re-scoring kernel (small query batches) and the re-scoring inside the per-query final kernel use the same
    fp64 arithmetic: identical scores and ids, both equal to the oracle."""
    from news_recsys_b200.retrieval import TopkIndex
    g = torch.Generator().manual_seed(31 + Q)
    c = torch.nn.functional.normalize(torch.randn(150_000, 96, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(Q, 96, generator=g), dim=1)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NRX_TOPK_FINAL_SPLIT", mode)
        idx = TopkIndex(c.to(DEV))
        s, i, s64 = idx.search(q.to(DEV), 50, want_scores64=True)
        out[mode] = (s.cpu(), i.cpu(), s64.cpu())
    assert torch.equal(out["0"][1], out["1"][1])
    assert torch.equal(out["0"][2], out["1"][2]) and torch.equal(out["0"][0], out["1"][0])
    _, ref_i = R.topk_ip(q, c, 50)
    assert torch.equal(out["1"][1], ref_i)


def test_pipelined_host_fed_search_equals_direct_search():
    """retrieval.PipelinedSearch (H2D of the next queries and D2H of the previous results under the running search) returns,
    one call late, exactly what a direct search of the same queries returns."""
    from news_recsys_b200.retrieval import PipelinedSearch, TopkIndex
    g = torch.Generator().manual_seed(3)
    c = torch.nn.functional.normalize(torch.randn(40_000, 64, generator=g), dim=1)
    idx = TopkIndex(c.to(DEV))
    batches = [torch.nn.functional.normalize(torch.randn(200, 64, generator=g), dim=1).pin_memory() for _ in range(5)]
    pipe = PipelinedSearch(lambda q: idx.search(q, 30), 200, 64, 30, DEV)
    got = []
    for b in batches:
        r = pipe.submit(b)
        if r is not None:
            got.append((r[0].clone(), r[1].clone()))
    last = pipe.drain()
    got.append((last[0].clone(), last[1].clone()))
    assert len(got) == len(batches)
    for b, (s, i) in zip(batches, got):
        ds, di = idx.search(b.to(DEV), 30)
        assert torch.equal(i, di.cpu()) and torch.equal(s, ds.cpu())
    with pytest.raises(Exception):
        pipe.submit(torch.randn(200, 64))        # pageable memory is refused
