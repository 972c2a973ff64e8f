"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU paths (shard ranges, the id
exchange of the data-parallel step, sharded top-k + merge == single index on the oracle)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    return dict(ret)


def _exchange(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from news_recsys_b200.parallel import gather_batch
    from news_recsys_b200.synthetic import mind_config, synth_batch
    cfg = mind_config("deep", {"user_id": 50, "item_id": 40, "category": 18, "subcategory": 70, "user_click_category": 18}, history_len=4)
    local = synth_batch(cfg, 8, seed=100 + rank)
    keys = [k for k in local if k != "label"]
    g = gather_batch(local, keys)
    ok = True
    for r in range(world):
        exp = synth_batch(cfg, 8, seed=100 + r)
        for k in keys:
            ok = ok and torch.equal(g[k][r * 8:(r + 1) * 8], exp[k])
    # dense-gradient exchange: AVG all-reduce == mean over ranks
    t = torch.full((5,), float(rank + 1))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t /= world
    ok = ok and torch.allclose(t, torch.full((5,), (world + 1) / 2))
    return bool(ok)


def _sharded_topk(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_path as R
    from news_recsys_b200.parallel import shard_range
    g = torch.Generator().manual_seed(0)
    c = torch.randn(1003, 16, generator=g)
    c[1000:] = c[7]  # ties across shards
    q = torch.randn(9, 16, generator=g)
    lo, hi = shard_range(1003, rank, world)
    s, i = R.topk_ip(q, c[lo:hi], 20)
    i = i + lo
    gs = [torch.empty_like(s) for _ in range(world)]
    gi = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(gs, s)
    dist.all_gather(gi, i)
    ms, mi = R.topk_merge(gs, gi, 20)
    rs, ri = R.topk_ip(q, c, 20)
    return bool(torch.equal(mi, ri) and torch.allclose(ms, rs))


def test_shard_range_partitions_exactly():
    from news_recsys_b200.parallel import shard_range
    for n in (0, 1, 7, 1000, 1000003):
        for w in (1, 2, 3, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_dp_id_exchange_world2():
    assert _run("_exchange") == {0: True, 1: True}


def test_sharded_topk_merge_equals_global_world2():
    assert _run("_sharded_topk") == {0: True, 1: True}


def _dense_dp_step(rank, world):
    """The scheme of DataParallelTrainer(table_update="dense") with the oracle standing in for the kernels: every rank
    computes dense gradients (tables included) on ITS half of the batch with a local-mean loss, the flat gradient is
    averaged across ranks, every rank owns one slice of the flat parameter buffer (same split as K7 / shard_range),
    runs AdamW on its slice only and the slices are exchanged."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_path as R
    from news_recsys_b200.parallel import shard_range
    from news_recsys_b200.synthetic import mind_config, synth_batch
    import importlib
    cfg = mind_config("fm", {"user_id": 50, "item_id": 40, "category": 18, "subcategory": 70, "user_click_category": 18})
    torch.manual_seed(3)
    model = importlib.import_module("news_recsys_b200.model.sort.fm.model").FM(cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    full = synth_batch(cfg, 32, seed=9, label_p=0.5)
    local = {k: v[rank * 16:(rank + 1) * 16] for k, v in full.items()}
    _, _, grads = R.loss_and_grads("fm", sd, cfg, local)
    names = sorted(sd)
    flat_g = torch.cat([grads[k].reshape(-1) for k in names])
    flat_p = torch.cat([sd[k].reshape(-1) for k in names])
    pad = (-flat_g.numel()) % 4
    flat_g = torch.cat([flat_g, torch.zeros(pad)]); flat_p = torch.cat([flat_p, torch.zeros(pad)])
    dist.all_reduce(flat_g, op=dist.ReduceOp.SUM)
    flat_g /= world
    lo, hi = shard_range(flat_g.numel() // 4, rank, world)
    lo, hi = 4 * lo, 4 * hi
    new_slice, _, _ = R.adamw_step(flat_p[lo:hi], flat_g[lo:hi], torch.zeros(hi - lo), torch.zeros(hi - lo), 1, 1e-3)
    parts = [torch.zeros(4 * (shard_range(flat_g.numel() // 4, r, world)[1] - shard_range(flat_g.numel() // 4, r, world)[0]))
             for r in range(world)]
    if parts[0].numel() == parts[1].numel():
        dist.all_gather(parts, new_slice.contiguous())
    else:   # gloo all_gather needs equal sizes: pad the shorter slice
        m = max(p.numel() for p in parts)
        buf = [torch.zeros(m) for _ in range(world)]
        dist.all_gather(buf, torch.cat([new_slice, torch.zeros(m - new_slice.numel())]))
        parts = [b[:p.numel()] for b, p in zip(buf, parts)]
    new_p = torch.cat(parts)
    # single process on the whole batch
    _, _, g_full = R.loss_and_grads("fm", sd, cfg, full)
    ref_g = torch.cat([g_full[k].reshape(-1) for k in names] + [torch.zeros(pad)])
    ref_p, _, _ = R.adamw_step(flat_p, ref_g, torch.zeros_like(flat_p), torch.zeros_like(flat_p), 1, 1e-3)
    return bool(torch.allclose(flat_g, ref_g, rtol=1e-5, atol=1e-8) and torch.allclose(new_p, ref_p, rtol=1e-6, atol=1e-8))


def test_dense_data_parallel_scheme_equals_single_process():
    out = _run("_dense_dp_step")
    assert out == {0: True, 1: True}
