"""Multi-GPU paths (one process per GPU, torch.distributed over NCCL / NVLink).

The reference is single-GPU only (`devices=1`, sort/deep/train.py:41-42); everything here is new.

* DataParallelTrainer — synchronous data parallelism for the sort models with REPLICATED parameters.
    table_update="dense" (default; the reference's optimizer semantics, dense AdamW over every row):
    each rank runs forward/backward and K3 on its OWN B samples, K3 writing dense table gradients into
    the same flat buffer as the tower gradients; ONE all-reduce (AVG) of that buffer; one dense AdamW.
    Per-rank work does not grow with the world size; the all-reduce moves the table bytes, so this is the
    scheme for tables that fit the NVLink budget of a step (MIND-small: ~8 MB) — larger tables shard.
    table_update="sparse" (lazy rows, touched rows only):
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

import ctypes as C
import os
from typing import Optional, Dict, List, Tuple

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
# NVLink peer memory (K7 plumbing)                                             #
# --------------------------------------------------------------------------- #

class PeerBuffer:
    """Device memory from `nrx_peer_alloc` (cudaMalloc'd, zero-filled, IPC-exportable), viewed as a torch tensor
    through `__cuda_array_interface__` (zero copy; the tensor keeps this object alive)."""

    def __init__(self, nbytes: int, device, typestr: str = "<f4"):
        self.lib = L.load()
        ptr = C.c_void_p()
        with torch.cuda.device(device):
            L.check(self.lib.nrx_peer_alloc(nbytes, C.byref(ptr)), "nrx_peer_alloc")
        self.ptr, self.nbytes, self.device = int(ptr.value), int(nbytes), torch.device(device)
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": typestr, "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device=self.device)

    def export(self) -> bytes:
        h = C.create_string_buffer(64)
        L.check(self.lib.nrx_peer_export(self.ptr, h), "nrx_peer_export")
        return h.raw

    def free(self):
        """cudaFree.  Only after every peer has closed its mapping (close_peers + a barrier) and every tensor view of this
        buffer has been dropped."""
        if self.ptr:
            with torch.cuda.device(self.device):
                L.check(self.lib.nrx_peer_free(self.ptr), "nrx_peer_free")
            self.ptr = 0


def open_peers(bufs: List["PeerBuffer"], group=None) -> List[List[int]]:
    """Exchange the IPC handles of `bufs` (same list on every rank) and map every peer's copy.
    Returns ptrs[k][j] = device address, valid on THIS rank, of rank j's k-th buffer (own buffers as they are)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = [b.export() for b in bufs]
    allh = [None] * world
    dist.all_gather_object(allh, mine, group=group)
    lib = L.load()
    out = []
    for k, b in enumerate(bufs):
        row = []
        for j in range(world):
            if j == rank:
                row.append(b.ptr)
                continue
            p = C.c_void_p()
            with torch.cuda.device(b.device):
                L.check(lib.nrx_peer_open(allh[j][k], C.byref(p)), "nrx_peer_open")
            row.append(int(p.value))
        out.append(row)
    return out


def close_peers(ptr_rows: List[List[int]], rank: int, device) -> None:
    """Unmap what open_peers mapped (every pointer of another rank's buffer)."""
    lib = L.load()
    with torch.cuda.device(device):
        for row in ptr_rows:
            for j, p in enumerate(row):
                if j != rank and p:
                    L.check(lib.nrx_peer_close(p), "nrx_peer_close")


# --------------------------------------------------------------------------- #
# data-parallel training                                                       #
# --------------------------------------------------------------------------- #

class DataParallelTrainer(FusedTrainer):
    _inline_update = False  # the row update needs the all-gathered gradients of every rank

    def __init__(self, model, B: int, kind=None, group=None, use_graph: bool = True, table_update: str = "dense",
                 exchange: str = "peer", peer_timeout_ms: int = 20000, **kw):
        """exchange (dense mode only): "peer" — K7, gradient all-reduce fused with AdamW over NVLink peer memory,
        the whole step one CUDA graph, no NCCL on the data path (single node); "nccl" — all-reduce between two graphs.
        peer_timeout_ms: how long K7 waits for a peer before it declares the exchange dead.  That is FATAL and sticky
        (no rank updates anything any more); step() / feed() raise NrxError at their next status read-back."""
        self.peer_timeout_ms = peer_timeout_ms
        if exchange not in ("peer", "nccl"):
            raise L.NrxError(f"exchange must be 'peer' or 'nccl', got {exchange!r}")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._dp_ready = False
        self._use_peer = table_update == "dense" and exchange == "peer"
        if self._use_peer and self.world > L.NRX_MAX_PEERS:
            raise L.NrxError(f"exchange='peer' supports up to {L.NRX_MAX_PEERS} ranks of one node")
        self._peer_bufs: List[PeerBuffer] = []
        super().__init__(model, B, kind=kind, use_graph=False, table_update=table_update, dense_impl="flat", **kw)
        dev = self.dev
        G = self.world
        self.graph_a = self.graph_b = None
        if self._use_peer:
            self._init_peer(use_graph)
            return
        if self.table_update == "dense":
            self._init_dense(use_graph)
            return
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

    # ---- dense mode over peer memory (K7): the step is the single-GPU graph with a different optimizer kernel ----
    def _alloc_flat(self, n: int) -> torch.Tensor:
        if not self._use_peer:
            return super()._alloc_flat(n)
        buf = PeerBuffer(4 * n, self.dev)
        self._peer_bufs.append(buf)   # order: flat_p, flat_g
        return buf.tensor()

    def _init_peer(self, use_graph):
        sig = PeerBuffer(4 * L.NRX_PEER_SIG_WORDS, self.dev, typestr="<i4")
        self._peer_bufs.append(sig)
        self.sig = sig.tensor()
        ptrs = open_peers(self._peer_bufs, self.group)   # [p, g, sig][rank]
        st = L.NrxPeerStep()
        st.rank, st.world = self.rank, self.world
        for j in range(self.world):
            st.p[j], st.g[j], st.sig[j] = ptrs[0][j], ptrs[1][j], ptrs[2][j]
        st.m, st.v, st.n = self.flat_m.data_ptr(), self.flat_v.data_ptr(), self.n_dense
        st.d_hparams = self.d_hp.data_ptr()
        st.beta1, st.beta2, st.eps, st.weight_decay = self.betas[0], self.betas[1], self.eps, self.wd
        st.status = self.id_status.data_ptr()    # bit 1: the exchange is dead (read back with the loss, raised by the host)
        st.timeout_ms = int(self.peer_timeout_ms)
        self._peer_step = st
        dist.barrier(group=self.group)    # every rank has mapped every buffer before the first kernel touches them
        if use_graph:
            self._capture()               # warm-up step (all ranks in lockstep) + capture of the whole step

    def _adamw_flat(self):
        if not self._use_peer:
            return super()._adamw_flat()
        L.check(self.lib.nrx_adamw_allreduce_peer(C.byref(self._peer_step), self._sp()), "nrx_adamw_allreduce_peer")

    def peer_timed_out(self) -> bool:
        """True if a K7 launch gave up waiting for a peer: the exchange is dead on every rank (sticky), parameters and
        moments were left untouched from that step on.  Synchronises the stream."""
        flag = C.c_int32(0)
        L.check(self.lib.nrx_peer_status(self.sig.data_ptr(), C.byref(flag), self._sp()), "nrx_peer_status")
        return bool(flag.value)

    # ---- dense mode: local K3 into the flat gradient buffer, ONE all-reduce, one AdamW ----------
    def _init_dense(self, use_graph):
        snap = self._snapshot()
        self._dense_step_eager()       # warm-up; also initialises the NCCL communicator
        torch.cuda.synchronize(self.dev)
        self._restore(snap)
        if use_graph:
            self.graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_a):
                self._fwd_bwd()
                self._dense_table_grads(self.fb, self._plan, self._gx)
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                self._adamw_flat()
            torch.cuda.synchronize(self.dev)
            self._restore(snap)

    def _dense_step_eager(self):
        self._fwd_bwd()
        self._dense_table_grads(self.fb, self._plan, self._gx)
        dist.all_reduce(self.flat_g[: self.n_dense], op=dist.ReduceOp.AVG, group=self.group)
        self._adamw_flat()

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
        if self._use_peer:
            return FusedTrainer.step(self)   # one graph; the exchange happens inside K7
        if self.table_update == "dense":
            if self.graph_a is None:
                self._dense_step_eager()
                return self.loss
            self.graph_a.replay()
            dist.all_reduce(self.flat_g[: self.n_dense], op=dist.ReduceOp.AVG, group=self.group)
            self.graph_b.replay()
            return self.loss
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
# row-sharded embedding tables (BASELINE config 5)                             #
# --------------------------------------------------------------------------- #

class ShardedEmbeddingTrainer(FusedTrainer):
    """Data-parallel step with the BIG embedding tables row-sharded across ranks.

    Rank r keeps rows [lo_r, hi_r) of every table with >= `shard_min_rows` rows, stored behind one extra
    all-zero row (local row 0); small tables stay replicated.  Per step, with fixed shapes throughout:
      1. ids / masks of all ranks are all-gathered (as in DataParallelTrainer);
      2. ids are remapped to LOCAL rows: owned -> id - lo + 1, not owned -> 0 (the zero row).  For replicated
         tables a rank keeps only the ids of its own sample block, so every (sample, feature) is produced once;
      3. K1 runs over the GLOBAL batch on the local shard: each rank emits the partial pooled features of the
         rows it owns (masked-mean denominators come from the full mask, so partials simply add up);
      4. one reduce-scatter(sum) hands every rank the complete features of its own B samples — single-id
         fields get exactly one non-zero contribution (bit-exact vs one GPU), pooled fields differ only in
         summation order;
      5. heads forward/backward locally; dense grads all-reduced (AVG), per-sample embedding grads all-gathered;
      6. K3 over the global batch with the local ids updates exactly the rows this rank owns (not-owned
         occurrences map to the padding sentinel and are skipped); replicated tables see all ids on every
         rank and stay bitwise identical.
    That exchange (`exchange="reduce_scatter"`) moves (G-1)*B*ΣD*4 bytes per rank and direction whatever the ids are.

    `exchange="peer"` (default when every sharded feature is a single-id feature) is the owner-compute exchange of SURVEY
    §8e over NVLink peer memory, with no collective on the vectors:
      1. ids all-gathered (tiny: 8 bytes per id);  K1 gathers the REPLICATED features of the local batch into x;
      2. `nrx_shard_push`: every rank gathers, for the samples of EVERY rank whose ids it owns, the rows of its shard and
         stores them straight into the requester's feature matrix x_r[b, cols] (peer stores) — each rank receives exactly
         B*ΣD_sharded*4 bytes; `nrx_peer_barrier` (flag barrier over the K7 signal pads) publishes x;
      3. heads forward/backward locally, dense grads all-reduced (AVG) — which also orders every rank's gradient matrix
         before step 4;
      4. `nrx_shard_pull`: every owner copies, from every rank's gradient matrix, the columns of the samples whose ids it
         owns (peer loads; replicated features: all samples) into its local [G*B, ΣD] buffer, and ONE K3 over the global
         batch with the masked local ids updates exactly the owned rows.
    Single-id features have exactly one contributor, so x is bit-equal to the single-GPU gather."""

    _inline_update = False

    def __init__(self, model, B: int, kind=None, group=None, shard_min_rows: int = 100_000, exchange: Optional[str] = None,
                 peer_timeout_ms: int = 20000, **kw):
        if exchange not in (None, "peer", "reduce_scatter"):
            raise L.NrxError(f"exchange must be 'peer' or 'reduce_scatter', got {exchange!r}")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        G = self.world
        dev = next(model.parameters()).device
        # ---- shard the big tables in place: [zero row | rows lo..hi) ----
        self.shards: Dict[str, Tuple[int, int, int]] = {}
        for name, emb in model.embedding_tables.items():
            rows = emb.weight.shape[0]
            if rows >= shard_min_rows and G > 1:
                lo, hi = shard_range(rows, self.rank, G)
                w = torch.zeros((hi - lo + 1, emb.weight.shape[1]), dtype=torch.float32, device=dev)
                w[1:].copy_(emb.weight.data[lo:hi])
                if lo == 0:
                    w[1].zero_()  # global padding row
                emb.weight = torch.nn.Parameter(w)
                self.shards[name] = (lo, hi, rows)
        # Replicated tables get the LOWEST table ids: the sort key is (table id, row), so their occurrences then
        # occupy the same sorted positions on every rank regardless of how many sharded rows a rank owns, which
        # keeps K3's summation tree — and therefore the replicas — bitwise identical across ranks.
        rep = [n for n in model.embedding_tables.keys() if n not in self.shards]
        sh = [n for n in model.embedding_tables.keys() if n in self.shards]
        model._table_ids = {n: i for i, n in enumerate(rep + sh)}
        kw.pop("use_graph", None)
        kw.setdefault("table_update", "sparse")   # every rank updates only the rows it owns (touched rows)
        super().__init__(model, B, kind=kind, use_graph=False, **kw)
        self.fm_fused = False  # the gather is distributed: K1 + field logits instead of the gather-fused FM
        self.id_keys = [key for key, dt, shape, off in self.layout.fields if key != "label"]
        mk = lambda: {key: torch.zeros((G * shape[0],) + tuple(shape[1:]), dtype=dt, device=dev)
                      for key, dt, shape, off in self.layout.fields if key != "label"}
        self.gbatch = mk()       # true global ids
        self.fbatch = mk()       # ids K1 sees: sharded -> local rows, replicated -> own sample block only
        self.bbatch = mk()       # ids the backward plan sees: sharded -> local rows, replicated -> global ids
        for k in self.id_keys:   # masks are shared
            if k.endswith("_mask"):
                self.fbatch[k] = self.gbatch[k]
                self.bbatch[k] = self.gbatch[k]
        self.gfb_fwd = ops.FeatBinding(self.fb.specs, model._weights(), self.fbatch, want_inv_den=True)
        self.gfb_bwd = ops.FeatBinding(self.fb.specs, model._weights(), self.bbatch, want_inv_den=True)
        for name, inv in self.gfb_fwd.inv_den.items():  # one set of denominators (written by K1, read by K3)
            self.gfb_bwd.inv_den[name] = inv
        for i, s in enumerate(self.gfb_bwd.specs):
            if s.name in self.gfb_fwd.inv_den:
                self.gfb_bwd.arr[i].inv_den = self.gfb_fwd.inv_den[s.name].data_ptr()
        self._bar_word = torch.zeros(1, dtype=torch.float32, device=dev)
        self.partial = torch.zeros((G * B, self.out_dim), dtype=torch.float32, device=dev)
        self.x_local = torch.zeros((B, self.out_dim), dtype=torch.float32, device=dev)
        self.gx_global = torch.zeros((G * B, self.out_dim), dtype=torch.float32, device=dev)
        own = torch.zeros(G * B, dtype=torch.bool, device=dev)
        own[self.rank * B:(self.rank + 1) * B] = True
        self._own_rows = own
        sharded_arrays = [sp.name for sp in self.fb.specs if sp.table in self.shards and sp.is_array]
        if exchange == "peer" and sharded_arrays:
            raise L.NrxError(f"exchange='peer' handles single-id features on sharded tables; array features {sharded_arrays} "
                             "need exchange='reduce_scatter' (owner-side partial pooling over peer memory is not built)")
        self.exchange = exchange or ("reduce_scatter" if (sharded_arrays or G == 1) else "peer")
        if self.exchange == "peer":
            self._init_peer_exchange(peer_timeout_ms)
        self._sharded_ready = True
        self.graph_a = self.graph_b = None
        if self.exchange == "peer" and os.environ.get("NRX_SHARDED_GRAPH", "1") == "1":
            self._capture_sharded()

    def _part_a(self):
        """Everything between the id all-gather and the dense-gradient all-reduce (graph A)."""
        self._remap_ids()
        self._fwd_bwd()
        torch.mul(self._gx, 1.0 / self.world, out=self.gx_peer)

    def _part_b(self):
        """Everything after the all-reduce: pull the gradient columns this rank owns, K3 + optimizer (graph B)."""
        L.check(self.lib.nrx_shard_pull(self._sh_feats, self._n_sh, self._rep_feats, self._n_rep, self.rank, self.world, self.B,
                                        self._g_ptrs, self.out_dim, self.gx_global.data_ptr(), self._sp()), "nrx_shard_pull")
        self._update(self.gfb_bwd, self._plan, self.gx_global)

    def _capture_sharded(self):
        """The step minus its two NCCL calls as two CUDA graphs (the eager step is ~50 launches + their Python: 0.9 ms at
        cfg5 on 2 GPUs against 0.18 ms for the single-GPU graph).  The peer barrier inside graph A is an ordinary kernel."""
        snap = self._snapshot()
        self._gather_ids()
        self._part_a()                       # warm-up (lazy module loading must not happen inside capture), lockstep on all ranks
        dist.all_reduce(self.flat_g if self.n_dense > 0 else self._bar_word, op=dist.ReduceOp.AVG, group=self.group)
        self._part_b()
        torch.cuda.synchronize(self.dev)
        self._restore(snap)
        dist.barrier(group=self.group)
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            self._part_a()
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, pool=ga.pool()):
            self._part_b()
        torch.cuda.synchronize(self.dev)
        self._restore(snap)
        self.graph_a, self.graph_b = ga, gb
        dist.barrier(group=self.group)

    # ---- owner-compute exchange over peer memory ---------------------------------------------------------------
    def _init_peer_exchange(self, timeout_ms: int):
        G, B, dev = self.world, self.B, self.dev
        n = B * self.out_dim
        bx, bg = PeerBuffer(4 * n, dev), PeerBuffer(4 * n, dev)
        bsig = PeerBuffer(4 * L.NRX_PEER_SIG_WORDS, dev, typestr="<i4")
        self._peer_keep = [bx, bg, bsig]
        self.x_peer = bx.tensor().view(B, self.out_dim)
        self.gx_peer = bg.tensor().view(B, self.out_dim)
        ptrs = open_peers(self._peer_keep, self.group)     # [x, gx, sig][rank]
        self._x_ptrs = L.ptr_array(ptrs[0], L.NRX_MAX_PEERS)
        self._g_ptrs = L.ptr_array(ptrs[1], L.NRX_MAX_PEERS)
        st = L.NrxPeerStep()
        st.rank, st.world = self.rank, G
        for j in range(G):
            st.sig[j] = ptrs[2][j]
        st.status = self.id_status.data_ptr()
        st.timeout_ms = int(timeout_ms)
        self._barrier_step = st
        sh = [sp for sp in self.fb.specs if sp.table in self.shards]
        rep = [sp for sp in self.fb.specs if sp.table not in self.shards]
        weights = self.model._weights()

        def mk(specs, sharded):
            arr = (L.NrxShardFeat * max(len(specs), 1))()
            for i, sp in enumerate(specs):
                w = weights[sp.table]
                ids = self.gbatch[sp.name]
                arr[i].table = w.data_ptr()
                arr[i].ids = ids.data_ptr()
                if sharded:
                    arr[i].lo, arr[i].hi = self.shards[sp.table][0], self.shards[sp.table][1]
                else:
                    arr[i].lo, arr[i].hi = 0, w.shape[0]
                arr[i].dim, arr[i].row_stride, arr[i].out_col = sp.dim, w.stride(0), sp.out_col
                arr[i].idx_dtype = L.IDX_I32 if ids.dtype == torch.int32 else L.IDX_I64
            return arr, len(specs)

        self._sh_feats, self._n_sh = mk(sh, True)
        self._rep_feats, self._n_rep = mk(rep, False)
        # K1 over the LOCAL batch for the replicated features only, writing their columns of x
        self._rep_fb = ops.FeatBinding(rep, weights, self.batch, want_inv_den=True) if rep else None
        if self._rep_fb is not None:    # K3 reads the denominators through the global binding: keep one set (local block)
            for name, inv in self._rep_fb.inv_den.items():
                pass
        dist.barrier(group=self.group)

    def _peer_barrier(self):
        L.check(self.lib.nrx_peer_barrier(C.byref(self._barrier_step), self._sp()), "nrx_peer_barrier")

    def _plan_fb(self):
        return self.gfb_bwd if getattr(self, "_sharded_ready", False) else self.fb

    def _exchange_ids(self):
        self._gather_ids()
        self._remap_ids()

    def _gather_ids(self):
        with dist._coalescing_manager(group=self.group, device=self.dev, async_ops=False):
            for k in self.id_keys:
                dist.all_gather_into_tensor(self.gbatch[k], self.batch[k], group=self.group)

    def _remap_ids(self):
        """Global ids -> the id views K1 / K3 use (static buffers, pure device work: part of the captured graph)."""
        peer = getattr(self, "exchange", "reduce_scatter") == "peer"
        for s in self.fb.specs:
            g = self.gbatch[s.name]
            if s.table in self.shards:
                lo, hi, _ = self.shards[s.table]
                local = torch.where((g >= lo) & (g < hi), g - (lo - 1), torch.zeros_like(g))
                if lo == 0:
                    local = torch.where(g == 0, torch.zeros_like(g), local)  # global pad id stays the pad row
                if not peer:
                    self.fbatch[s.name].copy_(local)
                self.bbatch[s.name].copy_(local)
            else:
                if not peer:
                    own = self._own_rows if g.dim() == 1 else self._own_rows[:, None]
                    self.fbatch[s.name].copy_(torch.where(own, g, torch.zeros_like(g)))
                self.bbatch[s.name].copy_(g)
        if peer:   # K1 only saw the local block: the masked-mean denominators of the global batch come from the gathered masks
            for name, inv in self.gfb_bwd.inv_den.items():
                torch.reciprocal(self.gbatch[name + "_mask"].sum(dim=1) + 1e-8, out=inv)

    def _embed_fwd(self):
        if getattr(self, "exchange", "reduce_scatter") == "peer":
            if self._rep_fb is not None:   # replicated features of my own samples -> their columns of x
                L.check(self.lib.nrx_embed_pool_fwd(self._rep_fb.arr, self._rep_fb.n, self.B, self.x_peer.data_ptr(), self.out_dim,
                                                    self.id_status.data_ptr(), self._sp()), "nrx_embed_pool_fwd")
            L.check(self.lib.nrx_shard_push(self._sh_feats, self._n_sh, self.rank, self.world, self.B, self._x_ptrs, self.out_dim,
                                            self._sp()), "nrx_shard_push")
            self._peer_barrier()           # every owner has written its rows into everybody's x
            return self.x_peer
        ops_out = ops.embed_pool_fwd(self.gfb_fwd, self.out_dim)  # partial features of the rows this rank owns
        dist.reduce_scatter_tensor(self.x_local, ops_out, op=dist.ReduceOp.SUM, group=self.group)
        return self.x_local

    def step(self) -> torch.Tensor:
        self._poll_status()
        if self.exchange == "peer":
            self._gather_ids()                 # all-gather: also orders step N's reads of x / gx before step N + 1's writes
            if self.graph_a is not None:
                self.graph_a.replay()
            else:
                self._part_a()
            # the all-reduce of the dense gradients doubles as the barrier "every rank's gx is written"
            dist.all_reduce(self.flat_g if self.n_dense > 0 else self._bar_word, op=dist.ReduceOp.AVG, group=self.group)
            if self.graph_b is not None:
                self.graph_b.replay()
            else:
                self._part_b()
            self._post_status()
            return self.loss
        self._exchange_ids()
        self._fwd_bwd()
        self._gx.mul_(1.0 / self.world)
        if self.n_dense > 0:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.AVG, group=self.group)
        dist.all_gather_into_tensor(self.gx_global, self._gx, group=self.group)
        self._update(self.gfb_bwd, self._plan, self.gx_global)
        self._post_status()
        return self.loss

    def exchange_bytes(self) -> int:
        """Bytes one rank sends in the forward exchange of one step (reduce-scatter of the [G*B, ΣD] partials: G - 1 blocks)."""
        return (self.world - 1) * self.B * self.out_dim * 4

    def exchange_bandwidth(self, iters: int = 10) -> dict:
        """Device-timed bandwidth of the forward exchange alone (same buffers as the step), max over ranks."""
        if self.exchange == "peer":
            return self._peer_exchange_bandwidth(iters)
        for _ in range(2):
            dist.reduce_scatter_tensor(self.x_local, self.partial, op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            dist.reduce_scatter_tensor(self.x_local, self.partial, op=dist.ReduceOp.SUM, group=self.group)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        ms = float(t[0])
        return {"collective": "reduce_scatter(sum) of the [G*B, ΣD] fp32 partial features", "bytes_sent_per_rank": self.exchange_bytes(),
                "ms": ms, "gb_per_s_per_rank": self.exchange_bytes() / (ms * 1e-3) / 1e9,
                "nvlink_peak_gb_per_s": 770.0, "note": "770 GB/s = measured peer copy per direction (B200_PROFILING.md)"}

    def _peer_exchange_bandwidth(self, iters: int) -> dict:
        sharded_cols = sum(sp.dim for sp in self.fb.specs if sp.table in self.shards)
        recv = self.B * sharded_cols * 4 * (self.world - 1) // self.world   # expected bytes arriving over NVLink (uniform ids)
        dist.barrier(group=self.group)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            L.check(self.lib.nrx_shard_push(self._sh_feats, self._n_sh, self.rank, self.world, self.B, self._x_ptrs, self.out_dim,
                                            self._sp()), "nrx_shard_push")
            self._peer_barrier()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        ms = float(t[0])
        return {"collective": "nrx_shard_push (owner-side gather stored straight into the requesters' feature rows over NVLink peer "
                              "memory) + nrx_peer_barrier; no NCCL on the vectors",
                "bytes_received_per_rank": recv, "bytes_reduce_scatter_would_send": self.exchange_bytes(),
                "ms": ms, "gb_per_s_per_rank": recv / (ms * 1e-3) / 1e9, "nvlink_peak_gb_per_s": 770.0,
                "note": "770 GB/s = measured peer copy per direction (B200_PROFILING.md); at this size the exchange is latency-bound"}

    def gather_table(self, name: str) -> torch.Tensor:
        """Full [rows, D] table on every rank (tests / checkpointing)."""
        w = self.model.embedding_tables[name].weight.data
        if name not in self.shards:
            return w.clone()
        lo, hi, rows = self.shards[name]
        base, rem = divmod(rows, self.world)
        mx = base + (1 if rem else 0)
        pad = torch.zeros((mx, w.shape[1]), dtype=w.dtype, device=w.device)
        pad[: hi - lo].copy_(w[1:])
        allp = torch.empty((self.world * mx, w.shape[1]), dtype=w.dtype, device=w.device)
        dist.all_gather_into_tensor(allp, pad, group=self.group)
        parts = []
        for r in range(self.world):
            a, b = shard_range(rows, r, self.world)
            parts.append(allp[r * mx: r * mx + (b - a)])
        return torch.cat(parts, dim=0)


# --------------------------------------------------------------------------- #
# sharded retrieval                                                            #
# --------------------------------------------------------------------------- #

class ShardedTopk:
    """Corpus rows [lo, hi) of this rank behind one TopkIndex; `search` returns the GLOBAL top-k on every rank.

    exchange="peer" (default, single node): `nrx_topk_search_peer` — every rank scans its shard for all queries and ships
    its exactly re-scored candidates (fp64 score + global id) into the inbox of the query's OWNER over NVLink peer memory;
    the owner of Q / world queries merges, proves completeness against every shard's bound, re-scans exactly through peer
    memory what it cannot prove, and stores the result into every rank's output.  No NCCL on the data path, and the
    per-query work (final sort, merge) divides by the world size.  exchange="nccl": per-shard complete top-k lists,
    all-gathered with their fp64 keys and merged on every rank (works across nodes; per-query work does not shrink)."""

    def __init__(self, local_corpus: torch.Tensor, n_total: int, group=None, exchange: str = "peer",
                 peer_timeout_ms: int = 20000, kprime: int = 0, use_graph: Optional[bool] = None):
        from .retrieval import TopkIndex
        self.use_graph = (os.environ.get("NRX_TOPK_GRAPH", "1") == "1") if use_graph is None else bool(use_graph)
        self.kprime = int(kprime)     # per-shard threshold rank of the peer search; 0 = library default
        if exchange not in ("peer", "nccl"):
            raise L.NrxError(f"exchange must be 'peer' or 'nccl', got {exchange!r}")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.exchange = exchange
        self.peer_timeout_ms = int(peer_timeout_ms)
        self.n_total = int(n_total)
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        if local_corpus.shape[0] != self.hi - self.lo:
            raise L.NrxError(f"rank {self.rank}: shard has {local_corpus.shape[0]} rows, expected {self.hi - self.lo}")
        if not local_corpus.is_cuda:
            raise L.NrxError("ShardedTopk: the corpus shard must be a CUDA tensor (no CPU fallback)")
        self.lib = L.load()
        self._peer = {}
        if exchange == "peer":
            if self.world > L.NRX_MAX_PEERS:
                raise L.NrxError(f"peer exchange supports at most {L.NRX_MAX_PEERS} ranks")
            n, d = local_corpus.shape
            self._cbuf = PeerBuffer(max(n * d * 4, 256), local_corpus.device)
            shard = self._cbuf.tensor()[: n * d].view(n, d)
            shard.copy_(local_corpus.detach().float())
            self._sig = PeerBuffer(4 * L.NRX_PEER_SIG_WORDS, local_corpus.device, typestr="<i4")
            ptrs = open_peers([self._cbuf, self._sig], group)
            self._corpus_ptrs, self._sig_ptrs = ptrs[0], ptrs[1]
            self._dev = local_corpus.device
            self._status = torch.zeros(1, dtype=torch.int32, device=local_corpus.device)
            self.index = TopkIndex(shard, id_base=self.lo)
        else:
            self.index = TopkIndex(local_corpus, id_base=self.lo)

    def search_local(self, queries: torch.Tensor, k: int):
        return self.index.search(queries, k)

    def close(self):
        """Collective: unmap the peers' buffers and free this rank's (corpus shard copy, signal pad, and the inbox / result
        buffers of every (Q, k) searched so far).  The object is unusable afterwards.  Peer buffers are cudaMalloc'd outside
        torch's allocator and mapped into every peer, so they are NOT released by garbage collection: a service that
        rebuilds its index (DSSM corpus refresh every epoch) must close the old one — DSSM.build_item_index does."""
        if self.exchange != "peer" or getattr(self, "_closed", False):
            return
        self._closed = True
        torch.cuda.synchronize(self._dev)
        dist.barrier(group=self.group)                      # nobody is inside a search any more
        rows = [self._corpus_ptrs, self._sig_ptrs]
        own = [self._cbuf, self._sig]
        for st in self._peer.values():
            rows += list(st["ptrs"])
            own += list(st["keep"])
        close_peers(rows, self.rank, self._dev)
        torch.cuda.synchronize(self._dev)
        dist.barrier(group=self.group)                      # every rank has dropped its mappings of my buffers
        self._peer.clear()
        self.index = None                                   # the tensor views die before the memory does
        for b in own:
            b.free()

    # -- peer path ----------------------------------------------------------------------------------------------
    def _peer_state(self, Q: int, k: int, dev, warm_queries: Optional[torch.Tensor] = None):
        """Inbox + result buffers for (Q, k), mapped on every rank (collective: every rank must search the same shapes)."""
        key = (Q, k)
        st = self._peer.get(key)
        if st is not None:
            return st
        inbox = PeerBuffer(max(int(self.lib.nrx_topk_peer_inbox_bytes(Q, self.world, k)), 256), dev)
        out_s = PeerBuffer(max(Q * k * 4, 256), dev)
        out_i = PeerBuffer(max(Q * k * 8, 256), dev)
        ptrs = open_peers([inbox, out_s, out_i], self.group)
        d = L.NrxTopkPeer()
        d.rank, d.world = self.rank, self.world
        for j in range(self.world):
            lo, hi = shard_range(self.n_total, j, self.world)
            d.corpus[j], d.n_rows[j] = self._corpus_ptrs[j], hi - lo
            d.inbox[j], d.out_scores[j], d.out_ids[j], d.sig[j] = ptrs[0][j], ptrs[1][j], ptrs[2][j], self._sig_ptrs[j]
        d.status = self._status.data_ptr()
        d.timeout_ms = self.peer_timeout_ms
        d.kprime = self.kprime
        ws = torch.empty(max(int(self.lib.nrx_topk_search_workspace_bytes(Q, self.index.N, self.index.D, k)), 16),
                         dtype=torch.uint8, device=dev)
        st = dict(desc=d, keep=(inbox, out_s, out_i), ptrs=ptrs, ws=ws,
                  out_s=out_s.tensor()[: Q * k].view(Q, k), out_i=out_i.tensor()[: Q * k * 2].view(torch.int64).view(Q, k),
                  status=torch.zeros(max(Q, 1), dtype=torch.int32, device=dev),
                  q=torch.zeros((max(Q, 1), self.index.D), dtype=torch.float32, device=dev), graph=None)
        self._peer[key] = st
        if Q and self.use_graph:
            # One search = 10 kernels + 2 memsets; enqueued one by one through ctypes the HOST is the bottleneck once the
            # shards are small (8 GPUs: ~120 us of device work per search).  Capture the whole search — flag barriers
            # included, as K7 does — and replay it: every rank captures the same sequence, so replays pair up like launches.
            if warm_queries is not None:
                st["q"].copy_(warm_queries.detach())
            self._launch_peer(st, Q, k)                 # eager once: lazy module loading must not happen inside capture
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.group)
            g = torch.cuda.CUDAGraph()
            cs = torch.cuda.Stream(device=dev)
            cs.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cs), torch.cuda.graph(g, stream=cs):
                self._launch_peer(st, Q, k)
            torch.cuda.current_stream(dev).wait_stream(cs)
            st["graph"] = g
        return st

    def _launch_peer(self, st, Q: int, k: int):
        q = st["q"]
        L.check(self.lib.nrx_topk_search_peer(self.index.index.data_ptr(), self.index.N, self.index.D, q.data_ptr(), q.stride(0), Q, k,
                                              C.byref(st["desc"]), st["status"].data_ptr(), st["ws"].data_ptr(), st["ws"].numel(),
                                              L.stream_ptr(q.device)), "nrx_topk_search_peer")

    def search_peer_(self, queries: torch.Tensor, k: int):
        """The peer search without the result copy: returns views of this rank's (re-used) output buffers, valid until the
        next search of the same (Q, k).  One enqueue, no host synchronisation."""
        Q = queries.shape[0]
        if not queries.is_cuda:
            raise L.NrxError("ShardedTopk: queries must be CUDA tensors")
        st = self._peer_state(Q, k, queries.device, warm_queries=queries)
        if Q:
            st["q"].copy_(queries.detach(), non_blocking=True)     # the search reads its static query buffer
            if st["graph"] is not None:
                st["graph"].replay()
                L.launch_count += L.KERNELS_PER_CALL["nrx_topk_search_peer"]
            else:
                self._launch_peer(st, Q, k)
        return st["out_s"], st["out_i"]

    def exact_fallbacks(self, Q: int, k: int) -> int:
        """How many of THIS rank's owned queries of the last (Q, k) peer search went through the exact scan."""
        return int(self._peer[(Q, k)]["status"].sum().item())

    def search(self, queries: torch.Tensor, k: int):
        """`queries` must be identical on every rank (replicated).  Scores travel with their fp64 ordering
        keys on both paths, so the result equals one index over the whole corpus bit for bit (rows closer than one fp32
        ulp included)."""
        if self.exchange == "peer":
            s, i = self.search_peer_(queries, k)
            if int(self._status.item()) & 2:
                raise L.NrxError("ShardedTopk: a peer did not arrive within peer_timeout_ms; the exchange is dead")
            return s.clone(), i.clone()
        from .retrieval import topk_merge
        s, i, s64 = self.index.search(queries, k, want_scores64=True)
        Q = s.shape[0]
        gs = torch.empty((self.world, Q, k), dtype=torch.float64, device=s.device)
        gi = torch.empty((self.world, Q, k), dtype=torch.int64, device=s.device)
        dist.all_gather_into_tensor(gs, s64, group=self.group)
        dist.all_gather_into_tensor(gi, i, group=self.group)
        return topk_merge(gs, gi)
