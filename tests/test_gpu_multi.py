"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`): data-parallel trainer on 2
ranks == the single-GPU trainer on the concatenated batch, and sharded retrieval == one index."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_RENDEZVOUS_ERRORS = ("address already in use", "eaddrinuse", "connection refused", "connection reset", "socket",
                      "store", "rendezvous", "ncclsystemerror", "ncclremoteerror", "unhandled system error")


def _spawn(fn, args_after_port, nprocs=2):
    """mp.spawn with a fresh port; ONE retry when the failure is the rendezvous itself (a port picked by _free_port() can
    be taken again before the workers bind it).  Assertion failures and kernel errors are never retried."""
    import torch.multiprocessing as mp
    for attempt in range(2):
        try:
            mp.spawn(fn, args=(nprocs, _free_port()) + tuple(args_after_port), nprocs=nprocs, join=True)
            return
        except Exception as e:   # ProcessRaisedException carries the worker's traceback as text
            msg = str(e).lower()
            if attempt == 0 and "assert" not in msg and "nrxerror" not in msg and any(k in msg for k in _RENDEZVOUS_ERRORS):
                continue
            raise


def _dp_worker(rank, world, port, kind, out_dir, mode, exchange):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import importlib
    import torch.distributed as dist
    from news_recsys_b200.parallel import DataParallelTrainer, ShardedTopk, shard_range
    from news_recsys_b200.synthetic import mind_config, synth_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rows = {"user_id": 300, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=6 if kind == "deep" else 0)
    name = {"fm": "FM", "deep": "Deep", "deepfm": "DeepFM"}[kind]
    cls = getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)
    torch.manual_seed(1)
    model = cls(cfg).to(f"cuda:{rank}")
    B = 128
    tr = DataParallelTrainer(model, B, kind=kind, table_update=mode, exchange=exchange)
    for s in range(3):
        full = synth_batch(cfg, B * world, seed=20 + s, label_p=0.5)
        local = {k: v[rank * B:(rank + 1) * B] for k, v in full.items()}
        tr.train_step(local)
    torch.cuda.synchronize()
    if mode == "dense" and exchange == "peer":
        assert not tr.peer_timed_out()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.save(sd, os.path.join(out_dir, f"dp_{kind}_{rank}.pt"))
    # sharded retrieval
    g = torch.Generator().manual_seed(0)
    c = torch.nn.functional.normalize(torch.randn(30001, 64, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(21, 64, generator=g), dim=1)
    lo, hi = shard_range(30001, rank, world)
    st = ShardedTopk(c[lo:hi].cuda(), 30001)
    s_, i_ = st.search(q.cuda(), 50)
    torch.save((s_.cpu(), i_.cpu()), os.path.join(out_dir, f"topk_{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,mode,exchange", [("fm", "dense", "peer"), ("deepfm", "dense", "peer"), ("deep", "dense", "peer"),
                                                ("fm", "dense", "nccl"), ("fm", "sparse", "nccl"), ("deepfm", "sparse", "nccl")])
def test_dp2_equals_single_gpu(kind, mode, exchange, tmp_path):
    """2 ranks x B=128 == one GPU x B=256 with the same table_update semantics (dense + peer: K7, all-reduce fused
    with AdamW over NVLink peer memory; dense + nccl: one NCCL all-reduce of the flat gradient buffer; sparse:
    global-batch row apply on every replica)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import importlib
    import torch.multiprocessing as mp
    from oracle import ref_path as R
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    _spawn(_dp_worker, (kind, str(tmp_path), mode, exchange))
    sd0 = torch.load(tmp_path / f"dp_{kind}_0.pt")
    sd1 = torch.load(tmp_path / f"dp_{kind}_1.pt")
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), f"replicas diverged on {k}"
    # single GPU on the concatenated batches
    rows = {"user_id": 300, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=6 if kind == "deep" else 0)
    name = {"fm": "FM", "deep": "Deep", "deepfm": "DeepFM"}[kind]
    cls = getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)
    torch.manual_seed(1)
    model = cls(cfg).cuda()
    tr = FusedTrainer(model, 256, kind=kind, table_update=mode)
    for s in range(3):
        tr.train_step(synth_batch(cfg, 256, seed=20 + s, label_p=0.5))
    ref = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for k in ref:
        if k.startswith("embedding_tables.") and kind == "fm":
            torch.testing.assert_close(sd0[k], ref[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
        else:
            upd, ref_upd = sd0[k] - ref[k], ref[k]
            assert float(upd.abs().max()) <= 2.5e-3, f"{k}: {float(upd.abs().max())}"  # <= ~2 Adam steps of lr=1e-3
    s0, i0 = torch.load(tmp_path / "topk_0.pt")
    s1, i1 = torch.load(tmp_path / "topk_1.pt")
    assert torch.equal(i0, i1)
    g = torch.Generator().manual_seed(0)
    c = torch.nn.functional.normalize(torch.randn(30001, 64, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(21, 64, generator=g), dim=1)
    rs, ri = R.topk_ip(q, c, 50)
    assert torch.equal(i0, ri)


def _sharded_worker(rank, world, port, kind, out_dir, exchange=None):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import importlib
    import torch.distributed as dist
    from news_recsys_b200.parallel import ShardedEmbeddingTrainer
    from news_recsys_b200.synthetic import mind_config, synth_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rows = {"user_id": 301, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=6 if kind == "deep" else 0)
    name = {"fm": "FM", "deep": "Deep", "widedeep": "WideDeep"}[kind]
    cls = getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)
    torch.manual_seed(1)
    model = cls(cfg).to(f"cuda:{rank}")
    B = 128
    tr = ShardedEmbeddingTrainer(model, B, kind=kind, shard_min_rows=100, exchange=exchange)   # user_id and item_id get sharded
    assert set(tr.shards) == {"user_id", "item_id"}
    # single-id features only (fm): owner-compute exchange over peer memory; a sharded array feature (deep + history): reduce-scatter
    assert tr.exchange == (exchange or ("peer" if kind == "fm" else "reduce_scatter"))
    assert model.embedding_tables["user_id"].weight.shape[0] < 301          # each rank holds a strict subset
    losses = []
    for s in range(3):
        full = synth_batch(cfg, B * world, seed=20 + s, label_p=0.5)
        local = {k: v[rank * B:(rank + 1) * B] for k, v in full.items()}
        losses.append(float(tr.train_step(local).item()))
    torch.cuda.synchronize()
    sd = {}
    for k, v in model.state_dict().items():
        if k.startswith("embedding_tables."):
            sd[k] = tr.gather_table(k[len("embedding_tables."):-len(".weight")]).cpu()
        else:
            sd[k] = v.detach().cpu()
    tr.check_status()
    torch.save((sd, losses), os.path.join(out_dir, f"sh_{kind}_{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,exchange", [("fm", None), ("fm", "reduce_scatter"), ("deep", None)])
def test_row_sharded_tables_equal_single_gpu(kind, exchange, tmp_path):
    """BASELINE config 5 mechanism at test scale: big tables row-sharded over 2 ranks, partial pooling +
    reduce-scatter forward, owner-side sparse-row AdamW backward == the single-GPU trainer on the same
    global batch (single-id fields and their tables: fp32-exact up to summation order 1e-5)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import importlib
    import torch.multiprocessing as mp
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    _spawn(_sharded_worker, (kind, str(tmp_path), exchange))
    sd0, l0 = torch.load(tmp_path / f"sh_{kind}_0.pt")
    sd1, l1 = torch.load(tmp_path / f"sh_{kind}_1.pt")
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), f"ranks disagree on {k}"
    rows = {"user_id": 301, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config(kind, rows, history_len=6 if kind == "deep" else 0)
    name = {"fm": "FM", "deep": "Deep"}[kind]
    cls = getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), name)
    torch.manual_seed(1)
    model = cls(cfg).cuda()
    tr = FusedTrainer(model, 256, kind=kind, table_update="sparse")   # the sharded trainer updates touched rows only
    ref_losses = [float(tr.train_step(synth_batch(cfg, 256, seed=20 + s, label_p=0.5)).item()) for s in range(3)]
    ref = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    # the global mean loss is the mean of the two ranks' local means
    for a, b, r in zip(l0, l1, ref_losses):
        assert abs(0.5 * (a + b) - r) <= (1e-5 if kind == "fm" else 2e-2) * max(1.0, abs(r))
    for k in ref:
        assert sd0[k].shape == ref[k].shape, k
        if kind == "fm":
            torch.testing.assert_close(sd0[k], ref[k], rtol=1e-5, atol=2e-6, msg=lambda m: f"{k}: {m}")
        else:
            assert float((sd0[k] - ref[k]).abs().max()) <= 2.5e-3, k   # bf16 tower: <= ~2 Adam steps of lr=1e-3


# --------------------------------------------------------------------------- sharded retrieval over peer memory
def _topk_cases():
    """(name, corpus, queries, k): deterministic, rebuilt identically in the workers and in the checking process."""
    g = torch.Generator().manual_seed(77)
    nrm = torch.nn.functional.normalize
    cases = []
    c = nrm(torch.randn(200_001, 128, generator=g), dim=1)
    q = nrm(torch.randn(301, 128, generator=g), dim=1)
    cases.append(("random", c, q, 100))
    # near ties across the shards: row `hi` (last shard) is row `lo` (first shard) nudged up by one ulp along q
    c2 = c.clone()
    q2 = q[:6].clone()
    for j in range(6):
        lo, hi = 500 + j, 200_001 - 500 - j
        c2[lo] = nrm(q2[j] + 0.05 * torch.randn(128, generator=g), dim=0)
        c2[hi] = c2[lo]
        e = int(q2[j].abs().argmax())
        c2[hi, e] = torch.nextafter(c2[lo, e], c2[lo, e] + torch.sign(q2[j, e]))
    cases.append(("near_ties", c2, q2, 100))
    # clustered: 6000 near-copies of the query's neighbourhood overflow the candidate lists -> owner-side exact scan over
    # BOTH shards through peer memory
    c3 = nrm(torch.randn(60_000, 64, generator=g), dim=1)
    q3 = nrm(torch.randn(9, 64, generator=g), dim=1)
    for j in range(3):
        rows = torch.randperm(60_000, generator=g)[:6000]
        c3[rows] = nrm(q3[j] + 0.02 * torch.randn(6000, 64, generator=g), dim=1)
    cases.append(("clustered", c3, q3, 64))
    # tiny corpus (no tensor-core filter at all), k larger than a shard, one query
    cases.append(("tiny", nrm(torch.randn(301, 32, generator=g), dim=1), nrm(torch.randn(1, 32, generator=g), dim=1), 200))
    cases.append(("k_gt_n", nrm(torch.randn(50, 16, generator=g), dim=1), nrm(torch.randn(5, 16, generator=g), dim=1), 64))
    return cases


def _topk_peer_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from news_recsys_b200.parallel import ShardedTopk, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = {}
    for name, c, q, k in _topk_cases():
        lo, hi = shard_range(c.shape[0], rank, world)
        st = ShardedTopk(c[lo:hi].cuda(), c.shape[0], exchange="peer")
        s1, i1 = st.search(q.cuda(), k)
        s2, i2 = st.search(q.cuda(), k)          # buffers and signal epochs are re-used
        assert torch.equal(i1, i2) and torch.equal(s1, s2), name
        nc = ShardedTopk(c[lo:hi].cuda(), c.shape[0], exchange="nccl")
        s3, i3 = nc.search(q.cuda(), k)
        out[name] = (s1.cpu(), i1.cpu(), s3.cpu(), i3.cpu(), st.exact_fallbacks(q.shape[0], k))
        del st, nc
    torch.cuda.synchronize()
    torch.save(out, os.path.join(out_dir, f"topk_peer_{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_topk_over_peer_memory_equals_one_index(tmp_path):
    """nrx_topk_search_peer on 2 ranks == the oracle on the whole corpus (ids bit-exact, near ties and the exact-scan
    fallback over peer-mapped shards included) == the NCCL list exchange; every rank holds the full result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import ref_path as R
    _spawn(_topk_peer_worker, (str(tmp_path),))
    r0 = torch.load(tmp_path / "topk_peer_0.pt")
    r1 = torch.load(tmp_path / "topk_peer_1.pt")
    fallbacks = {}
    for name, c, q, k in _topk_cases():
        s0, i0, sn, i_n, f0 = r0[name]
        s1, i1, _, _, f1 = r1[name]
        ref_s, ref_i = R.topk_ip(q, c, k)
        assert torch.equal(i0, i1) and torch.equal(s0, s1), f"{name}: ranks disagree"
        assert torch.equal(i0, ref_i), f"{name}: peer search differs from the oracle"
        assert torch.equal(i_n, ref_i), f"{name}: nccl exchange differs from the oracle"
        kk = min(k, c.shape[0])
        torch.testing.assert_close(s0[:, :kk], ref_s[:, :kk], rtol=1e-6, atol=1e-7)
        fallbacks[name] = f0 + f1
    assert fallbacks["random"] == 0, "the filter path must serve well-separated queries"
    assert fallbacks["clustered"] >= 3, "the clustered queries must go through the owner-side exact scan"
    assert fallbacks["tiny"] == 1 and fallbacks["k_gt_n"] == 5


# --------------------------------------------------------------------------- sharded corpus refresh of the DSSM model (f3)
def _dssm_cfg():
    from news_recsys_b200.synthetic import mind_config
    rows = {"user_id": 500, "item_id": 3002, "category": 18, "subcategory": 70, "user_click_category": 18}   # item ids 1..3001
    return mind_config("deep", rows, history_len=6)


def _dssm_sharded_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from news_recsys_b200.model.recall.DSSM.model import DSSM
    from news_recsys_b200.parallel import shard_range
    from news_recsys_b200.synthetic import synth_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = _dssm_cfg()
    torch.manual_seed(1)
    m = DSSM(cfg, hparams={"out_dim": 16}).to(dev)
    N = 3001
    items = synth_batch(cfg, N, seed=2)
    items["item_id"] = torch.randperm(N, generator=torch.Generator().manual_seed(5)) + 1    # position -> item id, not identity
    lo, hi = shard_range(N, rank, world)
    mine = {k: v[lo:hi].to(dev) for k, v in items.items()}
    m.build_item_index([mine], group=dist.group.WORLD, n_total=N)
    users = {k: v.to(dev) for k, v in synth_batch(cfg, 77, seed=11).items()}
    s, ids = m.retrieve_items(users, 25)
    # corpus refresh, epoch after epoch: the previous sharded index is closed (peer buffers unmapped and freed), results stay
    free = []
    for _ in range(4):
        m.build_item_index([mine], group=dist.group.WORLD, n_total=N)
        s2, ids2 = m.retrieve_items(users, 25)
        assert torch.equal(ids2, ids) and torch.equal(s2, s)
        torch.cuda.synchronize()
        free.append(torch.cuda.mem_get_info()[0])
    assert free[1] - free[-1] < 16 * 2 ** 20, f"peer buffers leak across refreshes: {[f >> 20 for f in free]} MiB free"
    torch.save((s.cpu(), ids.cpu(), m.index_item_ids.cpu()), os.path.join(out_dir, f"dssm_sh_{rank}.pt"))
    dist.destroy_process_group()


def test_dssm_sharded_corpus_refresh_equals_single_index(tmp_path):
    """DSSM.build_item_index(group=...): every rank embeds and indexes only its share of the corpus; retrieve_items over
    the sharded index == the same model with one index over the whole corpus (item ids bit-equal, position -> id map
    included), on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from news_recsys_b200.model.recall.DSSM.model import DSSM
    from news_recsys_b200.synthetic import synth_batch
    _spawn(_dssm_sharded_worker, (str(tmp_path),))
    s0, i0, map0 = torch.load(tmp_path / "dssm_sh_0.pt")
    s1, i1, map1 = torch.load(tmp_path / "dssm_sh_1.pt")
    assert torch.equal(i0, i1) and torch.equal(s0, s1) and torch.equal(map0, map1)
    cfg = _dssm_cfg()
    torch.manual_seed(1)
    m = DSSM(cfg, hparams={"out_dim": 16}).cuda()
    N = 3001
    items = synth_batch(cfg, N, seed=2)
    items["item_id"] = torch.randperm(N, generator=torch.Generator().manual_seed(5)) + 1
    m.build_item_index([{k: v.cuda() for k, v in items.items()}])
    users = {k: v.cuda() for k, v in synth_batch(cfg, 77, seed=11).items()}
    s, ids = m.retrieve_items(users, 25)
    assert torch.equal(map0, m.index_item_ids.cpu())
    assert torch.equal(i0, ids.cpu())
    torch.testing.assert_close(s0, s.cpu(), rtol=1e-6, atol=1e-7)
