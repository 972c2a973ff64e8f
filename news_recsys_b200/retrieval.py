"""Host-side wrappers of K6 (exact inner-product top-k) over the C ABI."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L


class TopkIndex:
    """faiss.IndexFlatIP equivalent: `add` packs the corpus for the tensor-core scan, `search` returns the
    exact top-k ordered by (inner product desc, id asc)."""

    def __init__(self, corpus: torch.Tensor, id_base: int = 0):
        if not corpus.is_cuda:
            raise L.NrxError("TopkIndex: the corpus must be a CUDA tensor (no CPU fallback)")
        self.corpus = corpus.detach().float().contiguous()
        self.N, self.D = self.corpus.shape
        self.id_base = int(id_base)
        lib = L.load()
        nbytes = int(lib.nrx_topk_index_bytes(self.N, self.D))
        if nbytes == 0:
            L.check(-2, "nrx_topk_index_bytes")
        self.index = torch.empty(nbytes, dtype=torch.uint8, device=corpus.device)
        L.check(lib.nrx_topk_index_build(self.corpus.data_ptr(), self.corpus.stride(0) if self.N else self.D, self.N, self.D,
                                         self.index.data_ptr(), nbytes, L.stream_ptr(corpus.device)), "nrx_topk_index_build")
        self._ws = None

    def search(self, queries: torch.Tensor, k: int, want_status: bool = False, want_scores64: bool = False):
        """-> (scores fp32 [Q,k], ids int64 [Q,k][, status][, scores64]): scores64 are the fp64 ordering keys."""
        q = queries.detach().float().contiguous()
        if not q.is_cuda:
            raise L.NrxError("TopkIndex.search: queries must be CUDA tensors")
        Q = q.shape[0]
        lib = L.load()
        nbytes = int(lib.nrx_topk_search_workspace_bytes(Q, self.N, self.D, k))
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=q.device)
        scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
        ids = torch.empty((Q, k), dtype=torch.int64, device=q.device)
        status = torch.zeros(max(Q, 1), dtype=torch.int32, device=q.device) if want_status else None
        s64 = torch.empty((Q, k), dtype=torch.float64, device=q.device) if want_scores64 else None
        L.check(lib.nrx_topk_search64(self.index.data_ptr(), self.corpus.data_ptr(), self.corpus.stride(0) if self.N else self.D,
                                      self.N, self.D, q.data_ptr(), q.stride(0) if Q else self.D, Q, k, self.id_base,
                                      scores.data_ptr(), L.ptr(s64), ids.data_ptr(), L.ptr(status), self._ws.data_ptr(),
                                      self._ws.numel(), L.stream_ptr(q.device)), "nrx_topk_search64")
        out = (scores, ids)
        if want_status:
            out += (status,)
        if want_scores64:
            out += (s64,)
        return out


def topk_merge(scores: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_lists, Q, k] per-shard lists -> global [Q, k] under (score desc, id asc).  fp64 scores (the ordering keys of
    search(want_scores64=True)) make the merge equal to one index bit for bit; fp32 scores order near-ties by id."""
    n, Q, k = scores.shape
    scores, ids = scores.contiguous(), ids.contiguous()
    out_s = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    lib = L.load()
    fn, name = (lib.nrx_topk_merge64, "nrx_topk_merge64") if scores.dtype == torch.float64 else (lib.nrx_topk_merge, "nrx_topk_merge")
    L.check(fn(scores.data_ptr(), ids.data_ptr(), n, Q, k, out_s.data_ptr(), out_i.data_ptr(), L.stream_ptr(scores.device)), name)
    return out_s, out_i


class PipelinedSearch:
    """Host-fed searches, software-pipelined one deep — the retrieval counterpart of `FusedTrainer.feed()`.

    `submit(host_queries)` (pinned float32 [Q, D]) enqueues the H2D copy of these queries on a copy stream (it overlaps the
    search submitted by the previous call), the search, and the D2H copy of its (scores, ids) into pinned buffers; it then
    returns the PREVIOUS submission's results (host tensors, valid until the call after next) or None on the first call.
    `drain()` returns the last one.  `search_fn(device_queries) -> (scores, ids)` is e.g. `TopkIndex(...).search` with k bound,
    or `ShardedTopk.search_peer_` (results in re-used buffers: pass copy_out=True so they are cloned before the next search
    may overwrite them)."""

    def __init__(self, search_fn, Q: int, D: int, k: int, device, copy_out: bool = False):
        self.fn, self.dev, self.copy_out = search_fn, torch.device(device), copy_out
        # two copy streams: on one, the H2D of search i + 1 would queue behind the D2H of search i - 1, which itself waits
        # for search i - 1 to finish — the next search could then never start before the previous one had ended AND been
        # copied out (measured: 398 instead of 340 us per search)
        self.copy = torch.cuda.Stream(device=self.dev)
        self.copy_out_stream = torch.cuda.Stream(device=self.dev)
        self.dq = [torch.empty((Q, D), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.hs = [torch.empty((Q, k), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.hi = [torch.empty((Q, k), dtype=torch.int64).pin_memory() for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]     # search finished: its query slot may be refilled
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.keep = [None, None]
        self.n = 0

    def submit(self, host_queries: torch.Tensor):
        if not host_queries.is_pinned():
            raise L.NrxError("PipelinedSearch.submit: queries must be pinned host memory")
        slot = self.n & 1
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy):
            if self.n >= 2:
                self.copy.wait_event(self.ev_done[slot])           # the search that read this slot two calls ago
            self.dq[slot].copy_(host_queries, non_blocking=True)
            self.ev_in[slot].record(self.copy)
        main.wait_event(self.ev_in[slot])
        s, i = self.fn(self.dq[slot])[:2]
        if self.copy_out:
            s, i = s.clone(), i.clone()
        self.ev_done[slot].record(main)
        self.keep[slot] = (s, i)                                   # alive until their D2H has been consumed
        with torch.cuda.stream(self.copy_out_stream):
            self.copy_out_stream.wait_event(self.ev_done[slot])
            self.hs[slot].copy_(s, non_blocking=True)
            self.hi[slot].copy_(i, non_blocking=True)
            self.ev_out[slot].record(self.copy_out_stream)
        prev = None
        if self.n > 0:
            p = slot ^ 1
            self.ev_out[p].synchronize()
            prev = (self.hs[p], self.hi[p])
        self.n += 1
        return prev

    def drain(self):
        if self.n == 0:
            return None
        p = (self.n - 1) & 1
        self.ev_out[p].synchronize()
        return self.hs[p], self.hi[p]
