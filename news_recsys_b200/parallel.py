"""Multi-GPU paths (one process per GPU, torch.distributed over NCCL / NVLink).

The reference is single-GPU only (`devices=1`, sort/deep/train.py:41-42); everything here is new.

* DataParallelTrainer — synchronous data parallelism for the sort models with REPLICATED parameters:
    ids (and masks) of every rank are all-gathered at step start (they are known before the forward),
    so the sort plan of the GLOBAL batch runs on the forked stream while each rank does forward/backward
    on its own B samples;  then one all-reduce (AVG) of the flat dense-gradient buffer and one all-gather
    of the per-sample embedding gradients [B, ΣD];  every rank then applies the identical fused sparse-row
    AdamW over the global batch in the identical order, so replicas stay bitwise equal without ever
    broadcasting parameters.  Two captured CUDA graphs (fwd+bwd | update) with the NCCL calls in between.
* ShardedTopk — retrieval with the corpus row-sharded across ranks: per-shard exact top-k with global ids
    (`id_base`), all-gather of the [Q, k] lists, local merge (score desc, id asc) => identical to one index.
* host-side helpers (shard ranges, batch exchange) are backend-agnostic and covered by gloo tests on CPU.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from .trainer import FusedTrainer


# --------------------------------------------------------------------------- #
# host-side helpers (no device code; tested with gloo)                         #
# --------------------------------------------------------------------------- #

def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of shard `rank`: the first n % world shards get one extra row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_batch(local: Dict[str, torch.Tensor], keys: List[str], group=None) -> Dict[str, torch.Tensor]:
    """All-gather the listed batch tensors along dim 0 (rank order) — the id exchange of the DP step."""
    world = dist.get_world_size(group)
    out = {}
    for k in keys:
        t = local[k].contiguous()
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t, group=group)
        out[k] = buf
    return out


# --------------------------------------------------------------------------- #
# data-parallel training                                                       #
# --------------------------------------------------------------------------- #

class DataParallelTrainer(FusedTrainer):
    def __init__(self, model, B: int, kind=None, group=None, use_graph: bool = True, **kw):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._dp_ready = False
        super().__init__(model, B, kind=kind, use_graph=False, **kw)
        dev = self.dev
        G = self.world
        # global batch (ids / masks of every rank, rank-major) and the global feature binding for the plan
        self.id_keys = [key for key, dt, shape, off in self.layout.fields if key != "label"]
        self.gbatch = {}
        for key, dt, shape, off in self.layout.fields:
            if key == "label":
                continue
            self.gbatch[key] = torch.zeros((G * shape[0],) + tuple(shape[1:]), dtype=dt, device=dev)
        self.gfb = ops.FeatBinding(self.fb.specs, model._weights(), self.gbatch, want_inv_den=True)
        self.gx_global = torch.zeros((G * B, self.out_dim), dtype=torch.float32, device=dev)
        self._dp_ready = True
        self.graph_a = self.graph_b = None
        # eager warm-up (also initialises NCCL communicators), then capture the two halves
        snap = self._snapshot()
        self._dp_step_eager()
        torch.cuda.synchronize(dev)
        self._restore(snap)
        if use_graph:
            self.graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_a):
                self._fwd_bwd()
                self._gx.mul_(1.0 / G)   # loss = mean over the GLOBAL batch
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                self._update(self.gfb, self._plan, self.gx_global)
            torch.cuda.synchronize(dev)
            self._restore(snap)

    def _plan_fb(self):
        return self.gfb if self._dp_ready else self.fb

    # ---- exchange steps ------------------------------------------------------------------------
    def _gather_ids(self):
        with dist._coalescing_manager(group=self.group, device=self.dev, async_ops=False):
            for k in self.id_keys:
                dist.all_gather_into_tensor(self.gbatch[k], self.batch[k], group=self.group)
        for name, inv in self.gfb.inv_den.items():  # masked-mean denominators of the global batch
            torch.reciprocal(self.gbatch[name + "_mask"].sum(dim=1) + 1e-8, out=inv)

    def _exchange_grads(self):
        if self.n_dense > 0:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.AVG, group=self.group)
        dist.all_gather_into_tensor(self.gx_global, self._gx, group=self.group)

    def _dp_step_eager(self):
        self._gather_ids()
        self._fwd_bwd()
        self._gx.mul_(1.0 / self.world)
        self._exchange_grads()
        self._update(self.gfb, self._plan, self.gx_global)

    def step(self) -> torch.Tensor:
        if self.graph_a is None:
            self._dp_step_eager()
            return self.loss
        self._gather_ids()
        self.graph_a.replay()
        self._exchange_grads()
        self.graph_b.replay()
        return self.loss

    def global_loss(self) -> torch.Tensor:
        t = self.loss.clone()
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        return t


# --------------------------------------------------------------------------- #
# sharded retrieval                                                            #
# --------------------------------------------------------------------------- #

class ShardedTopk:
    """Corpus rows [lo, hi) of this rank behind one TopkIndex; `search` returns the GLOBAL top-k on every rank."""

    def __init__(self, local_corpus: torch.Tensor, n_total: int, group=None):
        from .retrieval import TopkIndex
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        if local_corpus.shape[0] != self.hi - self.lo:
            raise L.NrxError(f"rank {self.rank}: shard has {local_corpus.shape[0]} rows, expected {self.hi - self.lo}")
        self.index = TopkIndex(local_corpus, id_base=self.lo)

    def search_local(self, queries: torch.Tensor, k: int):
        return self.index.search(queries, k)

    def search(self, queries: torch.Tensor, k: int):
        """`queries` must be identical on every rank (replicated)."""
        from .retrieval import topk_merge
        s, i = self.index.search(queries, k)
        Q = s.shape[0]
        gs = torch.empty((self.world, Q, k), dtype=torch.float32, device=s.device)
        gi = torch.empty((self.world, Q, k), dtype=torch.int64, device=s.device)
        dist.all_gather_into_tensor(gs, s, group=self.group)
        dist.all_gather_into_tensor(gi, i, group=self.group)
        return topk_merge(gs, gi)
