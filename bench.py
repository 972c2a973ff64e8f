#!/usr/bin/env python
"""bench.py — the measurement contract of this repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload deepfm] [--impl ours|reference]

A "step" is one full training step (forward + BCE + backward + optimizer update) of the hot path over one
batch of synthetic MIND-small-shaped input.  Default workload = BASELINE.json configs[1]: DeepFM ranking,
MIND-small table sizes, D = 16, batch 16384.

  value : train samples/s with the batches already resident in HBM (device pool larger than L2, one
          D2D copy of the batch blob + one CUDA-graph replay per step), CUDA-event timed, max over ranks
  e2e   : the same through the public FusedTrainer API from pinned HOST batches: per step one H2D copy of
          the batch blob, the step, and a D2H read of the loss
  roofline     : dominant API call of the step, timed live with CUDA events behind a queued blocker
  cpu_baseline : the oracle port of the reference path (torch CPU, oracle/ref_path.py) on the host cores
  --impl reference : times that same CPU arm as the main line (the reference is pure Python; /root/reference
          does not exist on the GPU box, so the arm is the oracle port — kind "port")
  legs  : the other BASELINE.json configurations in the same JSON line — cfg3 (DCN, B = 65536, data parallel, weak and
          strong), cfg1 (Deep + history, B = 1024), cfg5 (WideDeep, 1M / 160k / 10M-row tables, row-sharded at N >= 2),
          cfg4 (`retrieval`) — each with its own roofline; `eager_gpu`: the eager-PyTorch restatement of the reference
          step on the same GPU; `parity`: at N > 1, replicas bitwise equal + N-rank vs 1-rank steps
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/s (DeepFM fwd+bwd+update, MIND-small shape, batch 16384)"
UNIT = "samples/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j["bf16_tflops"]), "measured (MEASURED_PEAKS.json, burst)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def workload_cfg(name):
    from news_recsys_b200.synthetic import CFG1_ROWS, MIND_SMALL_ROWS, mind_config
    if name == "deepfm":
        return "deepfm", mind_config("deepfm", MIND_SMALL_ROWS), 16384, "cfg2: DeepFM, MIND-small rows, D=16, B=16384"
    if name == "fm":
        return "fm", mind_config("fm", MIND_SMALL_ROWS), 16384, "cfg2: FM, MIND-small rows, D=16, B=16384"
    if name == "dcn":
        return "dcn", mind_config("dcn", MIND_SMALL_ROWS), 65536, "cfg3: DCN d=112 bf16 tower, B=65536"
    if name == "deep":
        return "deep", mind_config("deep", CFG1_ROWS, history_len=50), 1024, "cfg1: Deep + user_history L=50, B=1024"
    if name == "widedeep":
        return "widedeep", mind_config("widedeep", MIND_SMALL_ROWS), 16384, "WideDeep, MIND-small rows, B=16384"
    if name == "widedeep_large":
        from news_recsys_b200.synthetic import MIND_LARGE_ROWS
        return ("widedeep", mind_config("widedeep", MIND_LARGE_ROWS), 16384,
                "cfg5: WideDeep, 1,000,001 users / 160,001 news / 10,000,000-row hashed table (D 32/32/17), B=16384 per GPU")
    raise SystemExit(f"unknown workload {name}")


def model_class(kind):
    import importlib
    cls = {"fm": "FM", "deep": "Deep", "widedeep": "WideDeep", "dcn": "DCN", "deepfm": "DeepFM", "lr": "LR"}[kind]
    return getattr(importlib.import_module(f"news_recsys_b200.model.sort.{kind}.model"), cls)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path
# ------------------------------------------------------------------------------------------------------------
def cpu_arm(kind, cfg, B, steps, warmup, seed=42):
    from oracle import ref_path as R
    from news_recsys_b200.synthetic import synth_batch
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    model = model_class(kind)(cfg)  # only used for shapes / the reference's own init; never leaves the CPU
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(leaf.values()), lr=cfg["train_hparams"]["lr"], betas=(0.9, 0.999))
    batches = [synth_batch(cfg, B, seed=1000 + i) for i in range(4)]
    times = []
    for s in range(warmup + steps):
        b = batches[s % len(batches)]
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = R.bce(R.model_forward(kind, leaf, cfg, b, dcn_materialise=False), b["label"][:, 0])
        loss.backward()
        opt.step()
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return B * len(times) / total, total / len(times) * 1e3, torch.get_num_threads()


# ------------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t_start = time.time()
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, windows=None):
        """windows: list of (t0, t1) wall-clock intervals that count as 'under load'."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                if windows:
                    ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if not any(a <= ts <= b for a, b in windows):
                        continue
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------------------
# per-API timing (CUDA events on the launching stream, queued behind a blocker so launch gaps do not count)
# ------------------------------------------------------------------------------------------------------------
class TimedLib:
    def __init__(self, lib, rec):
        self._lib, self._rec = lib, rec

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("nrx_") or name.endswith("_bytes") or name in ("nrx_last_error", "nrx_version", "nrx_tower_image_layout",
                                                                                   "nrx_embed_bwd_plan_is_staged"):
            return fn

        def timed(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            self._rec.append((name, s, e))
            return rc
        return timed


def profile_apis(trainer, pool, n=10):
    from news_recsys_b200 import _lib as L
    rec = []
    real = L._lib
    proxy = TimedLib(real, rec)
    L._lib = proxy
    trainer.lib = proxy
    graph, trainer.graph = trainer.graph, None
    snap = trainer._snapshot()
    per_step_launches = 0
    try:
        for i in range(n):
            trainer.load_blob(pool[i % len(pool)])
            torch.cuda.synchronize()
            torch.cuda._sleep(int(4e7))  # ~20 ms blocker: everything below is queued before the GPU gets to it
            c0 = L.launch_count
            trainer.step()
            per_step_launches = L.launch_count - c0
        torch.cuda.synchronize()
    finally:
        L._lib = real
        trainer.lib = real
        trainer.graph = graph
        trainer._restore(snap)
    agg = {}
    alias = {"nrx_embed_bwd_plan_stage": "nrx_embed_bwd_plan"}   # the plan enqueued in two halves is still one call of the path
    for name, s, e in rec:
        agg.setdefault(alias.get(name, name), []).append(s.elapsed_time(e) * 1e3)  # us
    return {k: sum(v) / len(v) * (len(v) / n) for k, v in agg.items()}, per_step_launches  # us per step per API


def _binding(flops, nbytes):
    """(bound, algorithmic quantity) of a tower call: the roofline that BINDS is the one with the larger time floor at
    the measured peaks.  At these layer widths (<= 224 -> 128 -> 128 -> 128 -> 64 -> 1) a training pass has to move the saved
    bf16 activation / dz images through HBM: ~50-100 flop per byte, well under the ridge (1725 TF/s / 6.45 TB/s = 267
    flop/B), so forward-with-save, dX and dW are HBM-bound; only the inference forward is tensor-bound."""
    hbm_peak, tf_peak, _ = peaks()
    t_tensor = flops / (tf_peak * 1e12)
    t_hbm = nbytes / (hbm_peak * 1e9)
    return ("tensor", flops) if t_tensor >= t_hbm else ("hbm", nbytes)


def algorithmic(kind, cfg, B, table_update="sparse"):
    """Algorithmic bytes / flops per launch of each API (formulas in DESIGN.md §Kernels)."""
    emb = cfg["embeddings"]
    feats = cfg["features"]
    share = emb.get("share_emb_table_features", {}) or {}
    names = sorted(set(feats["user_feature_names"]) | set(feats["item_feature_names"]))
    arr = set(feats.get("array_feature_names", []) or [])
    dims = {n: emb["embedding_size"][share.get(n, n)] for n in names}
    sd = sum(dims.values())
    k1 = 0.0
    n_occ = 0
    for n in names:
        if n in arr:
            Lh = feats["array_max_length"][n]
            k1 += Lh * (8 + 4) + (Lh / 2) * 4 * dims[n]
            n_occ += Lh
        else:
            k1 += 8 + 4 * dims[n]
            n_occ += 1
    k1 += 4 * sd
    mlp_in = {"deep": sd, "deepfm": sd, "dcn": 2 * sd, "widedeep": sd}.get(kind, 0)
    layers = [mlp_in, 128, 128, 128, 64, 1]
    tower_flops = 2 * sum(layers[i] * layers[i + 1] for i in range(5)) if mlp_in else 0
    # bf16 tile images of the hidden activations a_1..a_4 a TRAINING forward saves (2 bytes per unit); dz images are the same size
    tower_img_bytes = 2 * sum(layers[1:5]) if mlp_in else 0
    uniq_row_bytes = sum(dims[n] * 4 * 7 for n in names if n not in arr) + sum((feats["array_max_length"][n] / 2) * dims[n] * 4 * 7 for n in arr)
    n_table = sum(emb["embedding_table_size"][t] * emb["embedding_size"][t] for t in emb["embedding_size"])
    n_tower = sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(5)) if mlp_in else 0
    apply_bytes = B * (4 * sd + 8 * n_occ) + B * uniq_row_bytes
    adamw_bytes = 28 * n_tower          # p, g, m, v read + p, m, v written
    if table_update == "dense":         # K3 zero-fills and scatters dense table gradients; AdamW sweeps the tables too
        apply_bytes = B * (4 * sd + 8 * n_occ) + 4 * n_table + B * uniq_row_bytes / 7
        adamw_bytes = 28 * (n_tower + n_table)
    return {
        "nrx_embed_pool_fwd": ("hbm", B * k1),
        "nrx_fm_fused_fwd": ("hbm", B * (sum(8 + 4 * d for d in dims.values()) + 4 + 12)),
        "nrx_fm_fused_bwd": ("hbm", B * (sum(8 + 4 * d for d in dims.values()) + 4 + 4 * sd)),
        "nrx_field_logit_fwd": ("hbm", B * (4 * sd + 4)),
        "nrx_field_logit_bwd": ("hbm", B * (4 * sd + 4 + 8 * sd)),
        "nrx_logit_loss_fwd": ("hbm", B * 24),
        "nrx_tower_fwd": _binding(B * tower_flops, B * (2 * mlp_in + tower_img_bytes)),
        "nrx_tower_fwd_head": _binding(B * tower_flops, B * (2 * mlp_in + tower_img_bytes + 16)),
        "nrx_embed_pool_fwd_img": ("hbm", B * (k1 - 4 * sd + 2 * sd + (4 * sd if kind != "deep" else 0))),
        "nrx_dcn_cross_fwd_img": ("hbm", B * (4 * sd + 2 * 2 * sd)),
        "nrx_tower_bwd": _binding(2 * B * tower_flops, B * (4 * tower_img_bytes + 2 * mlp_in + 4 * mlp_in)),
        # dX: reads every saved activation image (the act' gates), writes every dz image and grad_x (fp32)
        "nrx_tower_bwd_dx": _binding(B * tower_flops, B * (2 * tower_img_bytes + 4 * mlp_in)),
        # dW: reads every a_l image (incl. the input image) and every dz_l image once
        "nrx_tower_bwd_dw": _binding(B * tower_flops, B * (2 * tower_img_bytes + 2 * mlp_in)),
        "nrx_dcn_cross_fwd": ("hbm", B * 4 * (sd + 2 * sd)),
        "nrx_dcn_cross_bwd": ("hbm", B * 4 * (sd + 2 * sd + sd)),
        "nrx_embed_bwd_plan": ("hbm", B * n_occ * (8 + 8 + 3 * 16)),
        "nrx_embed_bwd_apply": ("hbm", apply_bytes),
        "nrx_adamw_dense_dev": ("hbm", adamw_bytes),
    }


def _retrieval_traffic():
    """DRAM bytes of one cfg4 search from the committed ncu launch list (profiles/roofline_traffic.json), or None."""
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(tpath):
        return None
    return json.load(open(tpath)).get("retrieval", {}).get("nrx_topk_search")


def retrieval_leg(dev, world, rank, dist, quick):
    """BASELINE config 4: DSSM top-100 over a synthetic 1M x 128 corpus (row-sharded across ranks), Q = 1024
    L2-normalised queries per search.  queries/s = Q / (device time of one search incl. the shard merge)."""
    from news_recsys_b200.parallel import ShardedTopk, shard_range
    from news_recsys_b200.retrieval import TopkIndex
    N, D, K, Q = 1_000_000, 128, 100, 1024
    g = torch.Generator(device=dev).manual_seed(1234)  # identical stream on every rank: same corpus, same queries
    lo, hi = shard_range(N, rank, world)
    corpus = torch.empty((hi - lo, D), dtype=torch.float32, device=dev)
    chunk = 125_000
    for s in range(0, N, chunk):  # generate the global corpus in order, keep this rank's rows
        blk = torch.nn.functional.normalize(torch.randn((min(chunk, N - s), D), generator=g, device=dev), dim=1)
        a, b = max(s, lo), min(s + blk.shape[0], hi)
        if a < b:
            corpus[a - lo:b - lo] = blk[a - s:b - s]
    n_q = 4
    queries = [torch.nn.functional.normalize(torch.randn((Q, D), generator=g, device=dev), dim=1) for _ in range(n_q)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    exchange = os.environ.get("NRX_TOPK_EXCHANGE", "peer")
    if world > 1:
        index = ShardedTopk(corpus, N, exchange=exchange)
        search = (lambda q: index.search_peer_(q, K)) if exchange == "peer" else (lambda q: index.search(q, K))
        search_api = lambda q: index.search(q, K)
    else:
        index = TopkIndex(corpus)
        search = search_api = lambda q: index.search(q, K)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    for i in range(3):
        search(queries[i % n_q])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    iters = 3 if quick else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        s_, i_ = search(queries[i % n_q])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # status: how many queries needed the exact fallback scan (single-GPU index only)
    fb, sharded = None, None
    if world == 1:
        _, _, st = index.search(queries[0], K, want_status=True)
        fb = int(st.sum().item())
    else:
        # sharded result == the per-shard complete lists merged with their fp64 keys (the round-1 exchange), on every rank
        s_p, i_p = index.search(queries[0], K)
        t_fb = torch.tensor([index.exact_fallbacks(Q, K) if exchange == "peer" else 0], dtype=torch.int64, device=dev)
        other = ShardedTopk(corpus, N, exchange="nccl" if exchange == "peer" else "peer")
        s_n, i_n = other.search(queries[0], K)
        same = torch.tensor([int(torch.equal(i_p, i_n) and torch.equal(s_p, s_n))], dtype=torch.int64, device=dev)
        for i in range(3):
            other.search(queries[i % n_q], K)
        torch.cuda.synchronize()
        dist.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(iters):
            other.search_peer_(queries[i % n_q], K) if other.exchange == "peer" else other.search(queries[i % n_q], K)
        e3.record()
        torch.cuda.synchronize()
        t_o = torch.tensor([e2.elapsed_time(e3) / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t_fb)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        dist.all_reduce(t_o, op=dist.ReduceOp.MAX)
        fb = int(t_fb.item())
        sharded = {"exchange": exchange, "parity_ok": bool(same.item()),
                   "parity_check": "ids and scores of the peer search == per-shard complete lists all-gathered with fp64 keys and merged, on every rank",
                   f"ms_per_search_{other.exchange}": float(t_o.item())}
        del other
    # single-query latency ("also report Q=1 latency"): one search of one query, device-timed
    q1_ms = None
    if world == 1:
        q1 = queries[0][:1].contiguous()
        for _ in range(3):
            search(q1)
        torch.cuda.synchronize()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for _ in range(iters):
            search(q1)
        e5.record()
        torch.cuda.synchronize()
        q1_ms = e4.elapsed_time(e5) / iters
    # end to end: pinned host queries -> H2D -> search -> D2H of (scores, ids)
    from news_recsys_b200.retrieval import PipelinedSearch
    hq = [q.cpu().pin_memory() for q in queries]
    # public host-fed API, software-pipelined one deep like FusedTrainer.feed(): the H2D copy of search i + 1 and the D2H
    # copy of search i - 1 run on a copy stream under search i; every search's queries come from pinned host memory and its
    # (scores, ids) land in pinned host memory inside the timed region
    pipe = PipelinedSearch(search, Q, D, K, dev, copy_out=(world > 1))
    for i in range(3):
        pipe.submit(hq[i % n_q])
    pipe.drain()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        pipe.submit(hq[i % n_q])
    hs, hi_ = pipe.drain()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / iters
    s_chk, i_chk = search_api(queries[(iters - 1) % n_q])
    e2e_ok = bool(torch.equal(hi_, i_chk.cpu()) and torch.equal(hs, s_chk.cpu()))
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return None
    hbm_peak, tf_peak, peak_src = peaks()
    flops = 2.0 * Q * N * D  # algorithmic: one inner product per (query, corpus row); the implementation scans 1 + 1/8 times
    achieved = flops / (ms * 1e-3) / 1e12
    out = {"metric": "DSSM top-100 retrieval queries/s (1M x 128 corpus)", "value": Q / (ms * 1e-3), "unit": "queries/s",
           "ms_per_search": ms, "scaling": "strong (corpus sharded N/G per GPU)",
           "config": {"workload": "cfg4: N=1,000,000 x D=128 L2-normalised, k=100, Q=1024 per search", "shards": world,
                      "ordering": "(fp64 inner product desc, id asc), bit-exact vs oracle"},
           "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * K * 12,
                   "api": "retrieval.PipelinedSearch.submit(pinned queries): H2D + search + D2H of (scores, ids) every search, "
                          "pipelined one deep", "last_result_equals_direct_search": e2e_ok},
           "index_build_s": build_s, "fallback_queries": fb, "sharded": sharded, "q1_latency_ms": q1_ms,
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                        "traffic": _retrieval_traffic() if world == 1 else None,
                        "kernel": ("nrx_topk_search (query pack + sample scan of 1/8 of the tiles + theta + one full filter scan + final)"
                                   if world == 1 else
                                   "nrx_topk_search_peer, one CUDA graph per rank (shard scan + shard-side final + flag barrier + "
                                   "owner-side merge / proof / exact re-scan + flag barrier)"),
                        "peak_source": peak_src,
                        "algorithmic_per_launch": flops}}
    if world == 1 and not quick:
        # what faiss-cpu's IndexFlatIP does: blocked fp32 sgemm + top-k selection (no sort of the whole row); the oracle's
        # own definition (fp64 scores + stable sort, oracle/ref_path.py topk_ip) is timed next to it for reference
        from oracle import ref_path as R
        torch.set_num_threads(os.cpu_count() or 1)
        cc = corpus.cpu()
        nq_cpu = 256
        cq = queries[0][:nq_cpu].cpu()
        t0 = time.perf_counter()
        for a in range(0, nq_cpu, 64):
            torch.topk(cq[a:a + 64] @ cc.T, K, dim=1)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        R.topk_ip(cq[:16], cc, K, chunk=16)
        dt64 = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nq_cpu / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{nq_cpu} queries against the full 1M x 128 corpus: fp32 sgemm (blocks of 64 queries) + "
                                         "torch.topk, the faiss-cpu IndexFlatIP algorithm (reference TopKSearcher.py:77); near-ties "
                                         "are NOT pinned by it",
                               "oracle_fp64_stable_sort_queries_per_s": 16 / dt64}
    return out


# ------------------------------------------------------------------------------------------------------------
# legs: the other BASELINE configurations, measured with the same rules as the headline (device pool > L2 or a working
# set > L2, CUDA events around exactly K steps, max over ranks), each with the roofline of its dominant call
# ------------------------------------------------------------------------------------------------------------
def _pools(trainer, cfg, B, rank, dev, seed0=4242):
    from news_recsys_b200.synthetic import synth_batch
    blob_bytes = trainer.layout.nbytes
    host = []
    for i in range(4):
        hb = torch.empty(blob_bytes, dtype=torch.uint8).pin_memory()
        trainer.layout.pack(synth_batch(cfg, B, seed=seed0 + rank * 100003 + i), hb)
        host.append(hb)
    n_dev = max(4, min(96, int(160e6 // blob_bytes) + 1))
    pool = torch.empty((n_dev, blob_bytes), dtype=torch.uint8, device=dev)
    for i in range(n_dev):
        pool[i].copy_(host[i % len(host)])
    torch.cuda.synchronize()
    return host, pool


def _roofline(prof_trainer, pool, kind, cfg, B, table_update, workload, quick):
    per_api, launches_per_step = profile_apis(prof_trainer, pool, n=2 if quick else 6)
    alg = algorithmic(kind, cfg, B, table_update)
    hbm_peak, tf_peak, peak_src = peaks()
    forked = {"nrx_embed_bwd_plan", "nrx_hparams_step", "nrx_tower_pack"}
    if kind in ("deep", "deepfm", "widedeep", "dcn"):
        forked |= {"nrx_field_logit_bwd", "nrx_reduce2_f32", "nrx_reduce_f32"}
        if table_update == "sparse":
            forked.add("nrx_adamw_dense_dev")
    dom = max((k for k in per_api if k not in forked), key=lambda k: per_api[k])
    bound, qty = alg.get(dom, ("hbm", 0))
    dur_s = per_api[dom] * 1e-6
    if bound == "hbm":
        achieved, peak, runit = qty / dur_s / 1e9, hbm_peak, "GB/s"
    else:
        achieved, peak, runit = qty / dur_s / 1e12, tf_peak, "TFLOP/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(workload, {}).get(dom)
    roof = {"bound": bound, "achieved": achieved, "peak": peak, "unit": runit, "frac": achieved / peak, "traffic": traffic,
            "kernel": dom, "kernel_us": per_api[dom], "peak_source": peak_src, "algorithmic_per_launch": qty}
    breakdown = {}
    for k, us in sorted(per_api.items(), key=lambda kv: -kv[1]):
        b_, q_ = alg.get(k, ("hbm", 0))
        a_ = (q_ / (us * 1e-6) / 1e9) if b_ == "hbm" else (q_ / (us * 1e-6) / 1e12)
        breakdown[k] = {"us_per_step": round(us, 2), "stream": "forked" if k in forked else "main", "bound": b_,
                        "achieved": round(a_, 1), "frac": round(a_ / (hbm_peak if b_ == "hbm" else tf_peak), 4)}
    return roof, breakdown, launches_per_step


def train_leg(workload, dev, world, rank, dist, steps, warmup, mode, B=None, table_update="dense", quick=False, scaling="weak"):
    """One training workload on all ranks.  mode: "dp" (FusedTrainer at N = 1, DataParallelTrainer above) or "sharded"
    (FusedTrainer at N = 1, ShardedEmbeddingTrainer above).  Returns the leg's dict on rank 0, None elsewhere."""
    from news_recsys_b200.trainer import FusedTrainer
    kind, cfg, B0, wl_desc = workload_cfg(workload)
    B = B or B0
    torch.manual_seed(42)
    model = model_class(kind)(cfg).to(dev)
    if world == 1:
        trainer = FusedTrainer(model, B, kind=kind, table_update=table_update)
    elif mode == "dp":
        from news_recsys_b200.parallel import DataParallelTrainer
        trainer = DataParallelTrainer(model, B, kind=kind, table_update=table_update)
    else:
        from news_recsys_b200.parallel import ShardedEmbeddingTrainer
        trainer = ShardedEmbeddingTrainer(model, B, kind=kind)
    host, pool = _pools(trainer, cfg, B, rank, dev)
    n_pool = pool.shape[0]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        trainer.load_blob(pool[i % n_pool])
        trainer.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        trainer.load_blob(pool[(i + 3) % n_pool])
        trainer.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    final_loss = float(trainer.loss.item())
    for i in range(3):
        trainer.feed(host[i % len(host)])
    trainer.drain()
    barrier()
    e_steps = min(steps, 100)
    t0 = time.perf_counter()
    for i in range(e_steps):
        trainer.feed(host[i % len(host)])
    trainer.drain()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    exch = None
    if world > 1 and mode == "sharded":
        exch = trainer.exchange_bandwidth(iters=10)
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    out = None
    if rank == 0:
        blob_bytes = trainer.layout.nbytes
        out = {"metric": f"train samples/s ({wl_desc})", "value": world * B * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
               "batch_per_gpu": B, "global_batch": world * B, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
               "scaling": scaling, "table_update": trainer.table_update,
               "trainer": type(trainer).__name__,
               "e2e": {"value": world * B * e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": blob_bytes,
                       "d2h_bytes_per_step": 8, "steps": e_steps},
               "l2": f"inputs rotate over a {n_pool}-slot device pool ({n_pool * blob_bytes / 1e6:.0f} MB)", "final_loss": final_loss}
        if exch is not None:
            out["exchange"] = exch
        prof = trainer
        if world > 1:
            if workload == "widedeep_large":
                prof = None   # a private replica of the 1.3 GB tables only for per-call timing is not worth its memory traffic
            else:
                torch.manual_seed(42)
                prof = FusedTrainer(model_class(kind)(cfg).to(dev), B, kind=kind, table_update=trainer.table_update)
        if prof is not None:
            roof, breakdown, lps = _roofline(prof, pool, kind, cfg, B, trainer.table_update, workload, quick)
            out["roofline"], out["launches_per_step"] = roof, lps
            out["kernels"] = {k: v for k, v in list(breakdown.items())[:6]}
    del trainer, pool
    torch.cuda.empty_cache()
    return out


def eager_gpu_leg(workload, dev, steps=30, warmup=5):
    """the reference's step as eager PyTorch ops ON THE SAME GPU (the CPU restatement oracle/ref_path.py
    with its tensors on the device; the reference itself is pure PyTorch and cannot travel to the GPU box).  This is the
    competitor that matters — a measurement leg like cpu_baseline, never a product path."""
    from oracle import ref_path as R
    from news_recsys_b200.synthetic import synth_batch
    kind, cfg, B, desc = workload_cfg(workload)
    torch.manual_seed(42)
    model = model_class(kind)(cfg)
    leaf = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(leaf.values()), lr=cfg["train_hparams"]["lr"], betas=(0.9, 0.999))
    batches = [{k: v.to(dev) for k, v in synth_batch(cfg, B, seed=1000 + i).items()} for i in range(4)]

    def step(b):
        opt.zero_grad(set_to_none=True)
        loss = R.bce(R.model_forward(kind, leaf, cfg, b, dcn_materialise=False), b["label"][:, 0])
        loss.backward()
        opt.step()
        return loss

    for i in range(warmup):
        step(batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "batch": B, "workload": desc,
            "what": "eager PyTorch restatement of the reference step (fwd + BCE + bwd + torch.optim.AdamW) on this GPU, fp32, "
                    "DCN cross in its algebraically fused form (the reference's [B,d,d] form is slower still)"}


def parity_leg(dev, world, rank, dist):
    """N > 1: (1) replicas bitwise equal after data-parallel steps; (2) 3 steps at N ranks x B == 3 steps at 1 rank x N*B
    (tolerances of tests/test_gpu_multi.py: fp32 FM 1e-5 relative, bf16-tower DeepFM 2.5e-3 absolute)."""
    from news_recsys_b200.parallel import DataParallelTrainer
    from news_recsys_b200.synthetic import mind_config, synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    rows = {"user_id": 3001, "item_id": 2001, "category": 18, "subcategory": 270, "user_click_category": 18}
    res = {}
    ok = True
    for kind in ("fm", "deepfm"):
        cfg = mind_config(kind, rows)
        Bs = 256
        torch.manual_seed(7)
        model = model_class(kind)(cfg).to(dev)
        tr = DataParallelTrainer(model, Bs, kind=kind, table_update="dense")
        for s in range(3):
            full = synth_batch(cfg, Bs * world, seed=900 + s, label_p=0.5)
            tr.train_step({k: v[rank * Bs:(rank + 1) * Bs] for k, v in full.items()})
        torch.cuda.synchronize()
        tr.check_status()
        chk = tr.flat_p.view(torch.int32).to(torch.int64).sum().view(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        equal = all(int(c) == int(allc[0]) for c in allc)
        maxdiff = None
        if rank == 0:
            torch.manual_seed(7)
            ref_model = model_class(kind)(cfg).to(dev)
            ref = FusedTrainer(ref_model, Bs * world, kind=kind, table_update="dense")
            for s in range(3):
                ref.train_step(synth_batch(cfg, Bs * world, seed=900 + s, label_p=0.5))
            torch.cuda.synchronize()
            sd, rsd = model.state_dict(), ref_model.state_dict()
            if kind == "fm":
                maxdiff = max(float(((sd[k] - rsd[k]).abs() / rsd[k].abs().clamp_min(1e-3)).max()) for k in sd)
                good = maxdiff <= 1e-4
            else:
                maxdiff = max(float((sd[k] - rsd[k]).abs().max()) for k in sd)
                good = maxdiff <= 2.5e-3
            ok = ok and good
        ok = ok and equal
        res[kind] = {"replicas_bitwise_equal": equal, "n_rank_vs_1_rank_max_diff": maxdiff,
                     "tolerance": "1e-4 relative (fp32)" if kind == "fm" else "2.5e-3 absolute (bf16 tower, ~2 Adam steps of lr 1e-3)"}
        del tr
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["parity_ok"] = bool(int(flag))
    res["what"] = "3 data-parallel steps (K7 peer exchange) at N ranks x 256 vs 1 rank x N*256, same batches"
    return res if rank == 0 else None


_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="deepfm")
    ap.add_argument("--table-update", default="dense", choices=["dense", "sparse"],
                    help="dense: the reference's optimizer (dense AdamW over every table row, every step); "
                         "sparse: fused lazy sparse-row AdamW inside K3 (touched rows only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1, dense mode: peer = gradient all-reduce fused with AdamW over NVLink peer memory (one graph); "
                         "nccl = NCCL all-reduce between two graphs")
    ap.add_argument("--cpu-steps", type=int, default=0, help="override the CPU arm's step count")
    ap.add_argument("--no-retrieval", action="store_true", help="skip the DSSM top-100 retrieval leg (BASELINE config 4)")
    ap.add_argument("--no-legs", action="store_true", help="skip the cfg1 / cfg3 / cfg5 / eager-GPU / parity legs")
    ap.add_argument("--quick", action="store_true",
                    help="profiler mode: skip the clock-sampling load loop, shorten the per-API pass and the CPU arm")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, cfg, B, wl_desc = workload_cfg(args.workload)
    metric = METRIC if args.workload == "deepfm" else f"train samples/s ({wl_desc})"

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(3, min(args.steps, 200))     # a CPU step of this workload is ~16 ms: the whole arm stays under a minute
        warm = max(1, min(args.warmup, 50))
        v, ms, cores = cpu_arm(kind, cfg, B, steps, warm)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": wl_desc, "batch": B, "step": "fwd+bce+bwd+AdamW (oracle port, torch CPU)"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{steps} full steps of B={B} after {warm} warm-up"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    from news_recsys_b200 import _lib as L
    from news_recsys_b200.synthetic import synth_batch
    from news_recsys_b200.trainer import FusedTrainer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(42)  # same init on every rank (replicated parameters)
    model = model_class(kind)(cfg).to(dev)
    if world > 1:
        from news_recsys_b200.parallel import DataParallelTrainer
        trainer = DataParallelTrainer(model, B, kind=kind, table_update=args.table_update, exchange=args.exchange)
    else:
        trainer = FusedTrainer(model, B, kind=kind, table_update=args.table_update, dense_impl=os.environ.get('NRX_DENSE_IMPL', 'flat'))
    # batch pools: device pool > L2 (126 MB) so consecutive steps never find their inputs in L2
    blob_bytes = trainer.layout.nbytes
    n_pool = 8 if args.quick else max(8, int(160e6 // blob_bytes) + 1)  # --quick (profiler runs): few setup kernels
    host_pool = []
    for i in range(8):
        hb = torch.empty(blob_bytes, dtype=torch.uint8).pin_memory()
        trainer.layout.pack(synth_batch(cfg, B, seed=42 + rank * 100003 + i), hb)
        host_pool.append(hb)
    pool = torch.empty((n_pool, blob_bytes), dtype=torch.uint8, device=dev)
    for i in range(n_pool):
        pool[i].copy_(host_pool[i % len(host_pool)])
        if i >= len(host_pool):  # decorrelate the ids of the replicated slots
            v = trainer.layout.views(pool[i])
            for key, dt, shape, off in trainer.layout.fields:
                if dt in (torch.int64, torch.int32) and key in cfg["features"]["sparse_feature_names"]:
                    rows = cfg["embeddings"]["embedding_table_size"][key]
                    v[key].copy_((v[key] * 7919 + i) % (rows - 1) + 1)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region -------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        trainer.load_blob(pool[i % n_pool])
        trainer.step()
    clocks = Clocks(local_rank) if rank == 0 else None
    if clocks:
        time.sleep(0.3)  # let nvidia-smi start sampling
    barrier()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.load_blob(pool[(i + 7) % n_pool])
        trainer.step()
    e1.record()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    final_loss = float(trainer.loss.item())
    clk = {}
    windows, src = [(w0, w1)], "timed region"
    # If the timed region is shorter than the sampler period, keep the identical load running for ~1.5 s right
    # after it and sample the clocks there (NOT timed, does not enter `value`).  Every rank runs the same number
    # of extra steps (the steps contain collectives), decided from rank 0's measurement.
    n_extra = torch.tensor([0 if (args.quick or ms >= 1000.0) else int(1500.0 / max(ms / args.steps, 1e-3))],
                           dtype=torch.int64, device=dev)
    if dist is not None:
        dist.broadcast(n_extra, src=0)
    n_extra = int(n_extra.item())
    if n_extra > 0:
        snap = trainer._snapshot()
        for i in range(n_extra):
            trainer.load_blob(pool[i % n_pool])
            trainer.step()
            if i % 64 == 63:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        windows, src = [(w0, time.time())], "timed region + the same loop repeated for ~1.5 s right after it"
        trainer._restore(snap)
    barrier()
    if clocks:
        clk = clocks.stop(windows)
        clk["source"] = src
    # ---- end to end from pinned host memory ----------------------------------------------------------------
    # public API `feed(pinned_blob)`: per step one H2D copy of the step's inputs and one D2H read of its loss, software-
    # pipelined one deep (the copy of step i+1 overlaps the kernels of step i; the loss is read one step late)
    for i in range(3):
        trainer.feed(host_pool[i % len(host_pool)])
    trainer.drain()
    barrier()
    e_steps = min(args.steps, 200)
    t0 = time.perf_counter()
    for i in range(e_steps):
        trainer.feed(host_pool[i % len(host_pool)])
    e2e_loss = trainer.drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    tms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    retrieval = None
    if args.workload == "deepfm" and not args.no_retrieval:
        retrieval = retrieval_leg(dev, world, rank, dist, args.quick)
    legs, parity, eager = {}, None, None
    only = set(filter(None, os.environ.get("NRX_BENCH_LEGS", "").split(",")))   # developer filter, e.g. NRX_BENCH_LEGS=cfg5
    want = lambda name: not only or any(name.startswith(o) for o in only)
    if args.workload == "deepfm" and not args.no_legs and not args.quick:
        lsteps = max(10, min(args.steps, 50))
        # cfg3: DCN, B = 65536 — data parallel, weak (65536 per GPU) and strong (65536 in total)
        if want("cfg3"):
            legs["cfg3_dcn_weak"] = train_leg("dcn", dev, world, rank, dist, lsteps, args.warmup, "dp", quick=args.quick, scaling="weak")
        if world > 1 and want("cfg3"):
            legs["cfg3_dcn_strong"] = train_leg("dcn", dev, world, rank, dist, lsteps, args.warmup, "dp", B=65536 // world,
                                                quick=args.quick, scaling="strong")
        # cfg1: the reference's own CPU-runnable case (Deep + user_history, B = 1024)
        if want("cfg1"):
            legs["cfg1_deep_hist"] = train_leg("deep", dev, world, rank, dist, lsteps, args.warmup, "dp", quick=args.quick, scaling="weak")
        # cfg5: WideDeep with 1M / 160k / 10M-row tables; lazy sparse-row update (a dense sweep of 1.3 GB of tables per
        # step is what the reference's optimizer would do); row-sharded tables at N >= 2
        if want("cfg5"):
            legs["cfg5_widedeep_sharded"] = train_leg("widedeep_large", dev, world, rank, dist, lsteps, args.warmup, "sharded",
                                                      table_update="sparse", quick=args.quick, scaling="weak")
        if world > 1 and want("parity"):
            parity = parity_leg(dev, world, rank, dist)
        elif world == 1 and rank == 0 and want("eager"):
            eager = {"cfg2_deepfm": eager_gpu_leg("deepfm", dev), "cfg3_dcn": eager_gpu_leg("dcn", dev, steps=10, warmup=3)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (ms * 1e-3)
    e2e_v = world * B * e_steps / (e2e_ms * 1e-3)
    prof_trainer = trainer
    if world > 1:
        # per-kernel timing must not issue collectives from rank 0 alone: time the same per-GPU kernels on a
        # private single-GPU trainer (same model class / config / batch size)
        torch.manual_seed(42)
        prof_trainer = FusedTrainer(model_class(kind)(cfg).to(dev), B, kind=kind, table_update=args.table_update)
    variants = None
    if world == 1 and not args.quick:
        # the other optimizer semantics on the same pool, same timing rules (reported beside the headline, not as it)
        other = "sparse" if args.table_update == "dense" else "dense"
        torch.manual_seed(42)
        tr2 = FusedTrainer(model_class(kind)(cfg).to(dev), B, kind=kind, table_update=other)
        for i in range(max(args.warmup, 3)):
            tr2.load_blob(pool[i % n_pool]); tr2.step()
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for i in range(args.steps):
            tr2.load_blob(pool[(i + 7) % n_pool]); tr2.step()
        v1.record()
        torch.cuda.synchronize()
        vms = v0.elapsed_time(v1)
        variants = {f"table_update={other}": {"value": B * args.steps / (vms * 1e-3), "unit": UNIT, "ms_per_step": vms / args.steps,
                                              "note": ("lazy rows: fused sparse-row AdamW inside K3, untouched rows do not decay"
                                                       if other == "sparse" else "the reference's dense AdamW over every row")}}
        del tr2
        # ids both uniform and Zipf(1.05) (heavy-tailed click popularity: many duplicate rows per batch in the
        # sorted backward).  Same trainer, same timing rules, a second device pool of Zipf batches.
        from news_recsys_b200.synthetic import synth_batch
        zpool = torch.empty((n_pool, trainer.layout.nbytes), dtype=torch.uint8, device=dev)
        zhost = torch.empty(trainer.layout.nbytes, dtype=torch.uint8)
        for i in range(4):
            trainer.layout.pack(synth_batch(cfg, B, seed=777 + i, zipf=1.05), zhost)
            zpool[i::4].copy_(zhost.to(dev).unsqueeze(0).expand(zpool[i::4].shape[0], -1))
        snap = trainer._snapshot()
        for i in range(max(args.warmup, 3)):
            trainer.load_blob(zpool[i % zpool.shape[0]]); trainer.step()
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for i in range(args.steps):
            trainer.load_blob(zpool[(i + 3) % zpool.shape[0]]); trainer.step()
        v1.record()
        torch.cuda.synchronize()
        zms = v0.elapsed_time(v1)
        trainer._restore(snap)
        variants["ids=zipf(1.05)"] = {"value": B * args.steps / (zms * 1e-3), "unit": UNIT, "ms_per_step": zms / args.steps,
                                      "note": "the headline step on Zipf-distributed ids (the headline draws ids uniformly); "
                                              f"device pool of {zpool.shape[0]} slots ({zpool.numel() / 1e6:.0f} MB)"}
        del zpool
    per_api, launches_per_step = profile_apis(prof_trainer, pool, n=2 if args.quick else 10)
    alg = algorithmic(kind, cfg, B, args.table_update)
    hbm_peak, tf_peak, peak_src = peaks()
    # dominant call ON THE CRITICAL PATH: the sort plan and the optimizer clock run on the forked stream,
    # concurrently with forward + backward, and are listed in `kernels` with "stream": "forked"
    forked = {"nrx_embed_bwd_plan", "nrx_hparams_step", "nrx_tower_pack"}
    if kind in ("deep", "deepfm", "widedeep", "dcn"):  # these run beside the tower kernels on the third stream
        forked |= {"nrx_field_logit_bwd", "nrx_reduce2_f32", "nrx_reduce_f32"}
        if args.table_update == "sparse":
            forked.add("nrx_adamw_dense_dev")
    dom = max((k for k in per_api if k not in forked), key=lambda k: per_api[k])
    bound, qty = alg.get(dom, ("hbm", 0))
    dur_s = per_api[dom] * 1e-6
    if bound == "hbm":
        achieved, peak, runit = qty / dur_s / 1e9, hbm_peak, "GB/s"
    else:
        achieved, peak, runit = qty / dur_s / 1e12, tf_peak, "TFLOP/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
    roof = {"bound": bound, "achieved": achieved, "peak": peak, "unit": runit, "frac": achieved / peak, "traffic": traffic,
            "kernel": dom, "kernel_us": per_api[dom], "peak_source": peak_src,
            "algorithmic_per_launch": qty,
            "note": "dominant C-ABI call on the main stream (a call may be 2 kernels); at this batch every kernel moves a few MB "
                    "/ a few GFLOP and is launch-latency bound - the same kernels at B up to 1M are in profiles/r2_kernel_sweep_final.log"}
    breakdown = {}
    for k, us in sorted(per_api.items(), key=lambda kv: -kv[1]):
        b_, q_ = alg.get(k, ("hbm", 0))
        a_ = (q_ / (us * 1e-6) / 1e9) if b_ == "hbm" else (q_ / (us * 1e-6) / 1e12)
        breakdown[k] = {"us_per_step": round(us, 2), "stream": "forked" if k in forked else "main", "bound": b_,
                        "achieved": round(a_, 1),
                        "frac": round(a_ / (hbm_peak if b_ == "hbm" else tf_peak), 4)}
    cpu_steps = args.cpu_steps or (1 if args.quick else 20)
    cpu_base = None
    if world == 1:  # the CPU arm is timed on rank 0 at N = 1 only
        cv, cms, cores = cpu_arm(kind, cfg, B, cpu_steps, 2)
        cpu_base = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{cpu_steps} full steps of B={B} (oracle/ref_path.py, torch CPU, fwd+bce+bwd+AdamW)",
                    "ms_per_step": cms}
    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": wl_desc, "batch_per_gpu": B,
                   "step": ("fwd + BCE + bwd + dense AdamW over tables and tower (the reference's optimizer semantics), one CUDA graph"
                            if args.table_update == "dense" else
                            "fwd + BCE + bwd + optimizer (fused lazy sparse-row AdamW on tables, dense AdamW on the tower), one CUDA graph"),
                   "table_update": args.table_update,
                   "exchange": (None if world == 1 else (args.exchange if args.table_update == "dense" else "nccl")),
                   "tables": "fp32", "tower": "bf16 tcgen05, fp32 accumulate",
                   "l2": f"inputs rotate over a {n_pool}-slot device pool ({n_pool * blob_bytes / 1e6:.0f} MB > 126 MB L2); "
                         "the MIND-small tables (10 MB) are L2-resident by nature of the workload",
                   "parallelism": f"dp{world}"},
        "clocks": clk,
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": blob_bytes, "d2h_bytes_per_step": 4, "steps": e_steps,
                "api": "FusedTrainer.feed(pinned blob): H2D + step + D2H loss every step, pipelined one deep", "last_loss": e2e_loss},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "roofline": roof,
        "kernels": breakdown,
        "cpu_baseline": cpu_base,
        "final_loss": final_loss,
    }
    if variants is not None:
        line["variants"] = variants
    if retrieval is not None:
        line["retrieval"] = retrieval
    if legs:
        line["legs"] = legs
    if eager is not None:
        line["eager_gpu"] = eager
    if parity is not None:
        line["parity"] = parity
        line["parity_ok"] = parity["parity_ok"]
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
