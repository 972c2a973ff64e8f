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
