"""Validation scoring / metrics on the device.

What it replaces: `BaseModel.validation_step` (src/model/BaseModel/base_model.py:320-330: per batch a D2H copy and a
Python loop appending one (score, label) tuple per sample to a dict of lists) and `on_validation_epoch_end` (:333-478:
per user a Python sort, `roc_auc_score`, hand-rolled HR / NDCG / MRR @10, warm / cold split, overall AUC + log-loss).

Here `update()` keeps the batch on the device, `compute()` sorts once (two stable `torch.sort`s: score descending, then
user), runs `nrx_grouped_rank_metrics` (one thread per user, fp64) and returns the SAME dict the reference builds
(`Overall` / `Warm_Start` / `Cold_Start` with AUC, LogLoss, GAUC, NDCG@k, HR@k, MRR@k, User_Count).

Label pairing: the reference zips `user_id.view(-1)`, `scores.view(-1)` and `label.view(-1)` (:323-327), i.e. sample i is
paired with element i of the FLATTENED [B, n_labels] label tensor.  `label_pairing="reference"` (default) reproduces
that; `label_pairing="column0"` pairs sample i with label[i, 0] (what the training loss uses, deep/model.py:47)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib as L


class ValidationMetrics:
    def __init__(self, k: int = 10, user_in_train_set: Optional[Iterable] = None, label_pairing: str = "reference"):
        if label_pairing not in ("reference", "column0"):
            raise L.NrxError("label_pairing must be 'reference' or 'column0'")
        self.k, self.label_pairing = int(k), label_pairing
        self._u, self._s, self._l = [], [], []
        self.warm = None
        if user_in_train_set:
            ids = set()
            for x in user_in_train_set:   # the reference accepts the id as int or as its decimal string (:366)
                if isinstance(x, (int, np.integer)):
                    ids.add(int(x))
                elif isinstance(x, str) and x.lstrip("-").isdigit() and str(int(x)) == x:
                    ids.add(int(x))
            self.warm = torch.tensor(sorted(ids), dtype=torch.int64)

    def update(self, user_ids: torch.Tensor, scores: torch.Tensor, label: torch.Tensor):
        if scores.device.type != "cuda":
            raise L.NrxError("ValidationMetrics.update needs CUDA tensors (no CPU fallback)")
        u, s = user_ids.reshape(-1).to(torch.int64), scores.reshape(-1).to(torch.float32)
        l = label.reshape(-1) if self.label_pairing == "reference" else label[:, 0]
        n = min(u.numel(), s.numel(), l.numel())   # zip() stops at the shortest
        self._u.append(u[:n]); self._s.append(s[:n].detach()); self._l.append(l[:n].to(torch.float32))

    @staticmethod
    def _auc_sorted(s_desc: torch.Tensor, pos: torch.Tensor) -> float:
        """Tie-aware AUC of samples already sorted by score descending (fp64 on the device)."""
        n_pos = float(pos.sum()); n_neg = float(pos.numel()) - n_pos
        if n_pos == 0 or n_neg == 0:
            return 0.0
        change = torch.ones_like(s_desc, dtype=torch.bool)
        change[1:] = s_desc[1:] != s_desc[:-1]
        gid = torch.cumsum(change.to(torch.int64), 0) - 1
        G = int(gid[-1]) + 1
        p_g = torch.zeros(G, dtype=torch.float64, device=s_desc.device).index_add_(0, gid, pos.to(torch.float64))
        q_g = torch.zeros(G, dtype=torch.float64, device=s_desc.device).index_add_(0, gid, (~pos).to(torch.float64))
        neg_below = n_neg - (torch.cumsum(q_g, 0) - q_g) - q_g
        return float((p_g * (neg_below + 0.5 * q_g)).sum() / (n_pos * n_neg))

    @staticmethod
    def _logloss(p: torch.Tensor, l: torch.Tensor) -> float:
        """:449-455 — numpy float32 arithmetic on the clipped predictions; the mean is taken in fp64 here."""
        if p.numel() == 0:
            return 0.0
        pa = p.clamp(min=1e-15, max=float(np.float32(1 - 1e-15)))
        t = l * torch.log(pa) + (1 - l) * torch.log(1 - pa)
        return float(-t.to(torch.float64).mean())

    def compute(self) -> Dict[str, Dict[str, float]]:
        k = self.k
        names = ("Overall", "Warm_Start", "Cold_Start")
        if not self._u:
            return {n: ({"AUC": 0.0, "LogLoss": 0.0, "GAUC": 0.0, f"NDCG@{k}": 0.0, f"HR@{k}": 0.0, f"MRR@{k}": 0.0} |
                        ({} if n == "Overall" else {"User_Count": 0})) for n in names}
        u, s, l = torch.cat(self._u), torch.cat(self._s), torch.cat(self._l)
        dev = s.device
        o1 = torch.sort(s, descending=True, stable=True).indices
        o2 = torch.sort(u[o1], stable=True).indices
        perm = o1[o2]
        us, ss, ls = u[perm].contiguous(), s[perm].contiguous(), l[perm].contiguous()
        uniq, counts = torch.unique_consecutive(us, return_counts=True)
        U = uniq.numel()
        seg = torch.zeros(U + 1, dtype=torch.int64, device=dev)
        seg[1:] = torch.cumsum(counts, 0)
        per_user = torch.empty((U, 4), dtype=torch.float64, device=dev)
        flags = torch.empty(U, dtype=torch.int32, device=dev)
        L.check(L.load().nrx_grouped_rank_metrics(ss.data_ptr(), ls.data_ptr(), seg.data_ptr(), U, k, per_user.data_ptr(),
                                                  flags.data_ptr(), L.stream_ptr(dev)), "nrx_grouped_rank_metrics")
        cold_user = torch.zeros(U, dtype=torch.bool, device=dev)
        if self.warm is not None:   # a NON-EMPTY train set was given (base_model.py:364); ids it does not list are cold, also
            # when none of its entries parses as a user id (the reference then marks every user cold, :366)
            cold_user = ~torch.isin(uniq, self.warm.to(dev)) if self.warm.numel() > 0 else torch.ones(U, dtype=torch.bool, device=dev)
        # sample-level views in global score-descending order for the overall AUC / log-loss of each group
        s1, pos1 = s[o1], l[o1] == 1
        inv = torch.searchsorted(uniq, u[o1])
        cold1 = cold_user[inv]
        pu, fl, cu = per_user.cpu().numpy(), flags.cpu().numpy(), cold_user.cpu().numpy()
        out = {}
        for name, umask, smask in (("Overall", np.ones(U, bool), None), ("Warm_Start", ~cu, ~cold1), ("Cold_Start", cu, cold1)):
            sel_s, sel_pos, sel_l = (s1, pos1, l[o1]) if smask is None else (s1[smask], pos1[smask], l[o1][smask])
            has_auc = umask & ((fl & 1) == 1)
            mean = lambda a: float(np.mean(a)) if a.size else 0.0
            d = {"AUC": self._auc_sorted(sel_s, sel_pos) if sel_s.numel() else 0.0,
                 "LogLoss": self._logloss(sel_s, sel_l),
                 "GAUC": mean(pu[has_auc, 0]), f"NDCG@{k}": mean(pu[umask, 1]), f"HR@{k}": mean(pu[umask, 2]),
                 f"MRR@{k}": mean(pu[umask, 3])}
            if name != "Overall":
                d["User_Count"] = int(umask.sum())
            out[name] = d
        return out

    def reset(self):
        self._u, self._s, self._l = [], [], []
