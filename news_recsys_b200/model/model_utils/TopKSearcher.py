"""TopKSearcher — mirror of reference src/model/model_utils/TopKSearcher.py (faiss.IndexFlatIP wrapper):
`update_embedding(nn.Embedding, normalize)` (:19-49) and `search(List[Tensor[D]], normalize)` (:51-83)
-> (List[List[int]], List[List[float]]).  The index and the search run on the GPU (K6); results are the
exact top-k ordered by (inner product desc, id asc)."""
from typing import List, Tuple

import torch
import torch.nn as nn

from ... import ops
from ...retrieval import TopkIndex


class TopKSearcher:
    def __init__(self, k: int, use_gpu: bool = True, device="cuda"):
        self.k = k
        self.index = None
        self.use_gpu = use_gpu  # kept for signature parity; this implementation is GPU-only
        self.dimension = None
        self.device = torch.device(device)

    def update_embedding(self, emb_layer: nn.Embedding, normalize: bool = False):
        w = emb_layer.weight.detach().to(self.device, torch.float32)
        self.dimension = w.shape[1]
        if normalize:
            w = ops.l2_normalize(w)
        self.index = TopkIndex(w)
        print(f"[TopKSearcher] Index updated. Size: {w.shape[0]}, Dim: {self.dimension}")

    def search(self, query_embeddings: List[torch.Tensor], normalize: bool = False) -> Tuple[List[List[int]], List[List[float]]]:
        if self.index is None:
            raise ValueError("Index not initialized. Please call update_embedding first.")
        if len(query_embeddings) == 0:
            return [], []
        q = torch.stack(query_embeddings).detach().to(self.device, torch.float32)
        if normalize:
            q = ops.l2_normalize(q)
        scores, ids = self.index.search(q, self.k)
        return ids.tolist(), scores.tolist()
