"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (PyTorch, fp32 by default, fp64 on request) of the
embedding-and-interaction hot path of ZhangHaoyang493/News_Recsys.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module, and only as the checker / the CPU
arm that is timed beside the GPU number.

Parity status
-------------
* Sort models (LR / FM / Deep / WideDeep / DCN): PINNED.  `oracle/make_golden.py`
  imports the reference's own, unmodified modules (behind `oracle/refshim`, which
  holds no arithmetic) in the build container and writes inputs, parameters,
  outputs and gradients to `tests/golden/*.npz`; `tests/test_oracle_golden.py`
  checks every function below against those vectors.
* DSSM (towers, per-feature gather / masked pooling, forward with in-batch negatives and
  L2 normalisation, InfoNCE): PINNED.  `recall/DSSM/model.py` is not importable as
  shipped, but `oracle/make_golden_dssm.py` runs the UNMODIFIED class with import aliases
  only (top-level `BaseModel` / `model_utils` / `DataReader` names bound to the same
  reference modules, an empty `faiss` stub, the stale `get_features_embedding` name
  aliased to `get_feature_embedding`) and writes `tests/golden/dssm.npz`;
  `tests/test_oracle_golden.py::test_dssm_matches_reference` checks tower inputs
  (bit-equal), towers, outputs, loss and every gradient.
* Batch ingestion (`DataReader.__getitem__` + default collate): PINNED.
  `oracle/make_golden_ingest.py` writes a text file in the reference's wire format and
  the batches the reference's own DataReader + default collate produce from it
  (`tests/golden/ingest_features.txt`, `ingest.npz`).
* Validation scoring / metrics (`validation_step` + `on_validation_epoch_end`): PINNED.
  `oracle/make_golden_valmetrics.py` runs the reference's own methods on a reference Deep
  model and captures the `results` dict they build (`tests/golden/valmetrics.npz`).
* DeepFM and inner-product top-k: "parity unpinned".  The reference has no DeepFM
  class (composed from the pinned FM logit + pinned MLP) and `faiss` is not vendored,
  so these are restatements of the cited lines with no reference-produced vector
  behind them.

Each function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# Embedding gather + pooling + concat                                          #
# --------------------------------------------------------------------------- #

def table_name(feature: str, share: Dict[str, str]) -> str:
    """src/model/BaseModel/base_model.py:119-122 (`_get_emb_feature_name`)."""
    return share.get(feature, feature)


def input_dim(features: Iterable[str], emb_size: Dict[str, int], share: Dict[str, str]) -> int:
    """base_model.py:124-139 (`_calculate_input_dim`), sparse/array features only
    (dense features crash in the reference, base_model.py:94,129)."""
    total = 0
    for f in features:
        d = emb_size.get(table_name(f, share))
        total += 8 if d is None else d  # default 8: base_model.py:134-137
    return total


def feature_embedding(tables: Dict[str, Tensor], share: Dict[str, str], feature: str,
                      value: Tensor) -> Tensor:
    """base_model.py:262-271 (`get_feature_embedding`), embedding branch:
    `nn.Embedding(size, dim, padding_idx=0)(value.long())` — forward is a plain
    row gather (padding_idx only affects init / gradient)."""
    t = table_name(feature, share)
    if t not in tables:
        raise ValueError(f"Embedding table not found for {feature} (mapped to {t})")
    return F.embedding(value.long(), tables[t], padding_idx=0)


def array_feature_pooling(emb: Tensor, mask: Optional[Tensor]) -> Tensor:
    """base_model.py:273-282: no mask -> mean over L incl. pad rows;
    else sum(emb*mask) / (sum(mask) + 1e-8)."""
    if mask is None:
        return emb.mean(dim=1)
    m = mask.unsqueeze(-1).to(emb.dtype)
    return (emb * m).sum(dim=1) / (m.sum(dim=1) + 1e-8)


def embeddings_from_batch(tables: Dict[str, Tensor], batch: Dict[str, Tensor],
                          feature_names: Iterable[str], array_names: Iterable[str],
                          share: Dict[str, str]) -> Tuple[Tensor, List[int], List[str]]:
    """base_model.py:284-308 (`get_embeddings_from_batch`): sorted feature names,
    per-feature gather (+ pooling for array features, mask key `<name>_mask`),
    `torch.cat(dim=1)`; names missing from the batch are skipped but still
    listed in the returned names (base_model.py:291-293,308)."""
    names = sorted(list(feature_names))
    array_names = set(array_names)
    embs, dims = [], []
    for f in names:
        if f not in batch:
            continue
        e = feature_embedding(tables, share, f, batch[f])
        if f in array_names:
            e = array_feature_pooling(e, batch.get(f"{f}_mask", None))
        embs.append(e)
        dims.append(e.shape[1])
    if not embs:
        return torch.tensor([]), [], []
    return torch.cat(embs, dim=1), dims, names


# --------------------------------------------------------------------------- #
# Heads                                                                        #
# --------------------------------------------------------------------------- #

def fm_split(features: Tensor, dims: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """src/model/sort/fm/model.py:48-59: per field col 0 -> w, cols 1.. -> v
    (torch.stack => all fields must share D)."""
    w, v, s = [], [], 0
    for d in dims:
        w.append(features[:, s:s + 1])
        v.append(features[:, s + 1:s + d])
        s += d
    return torch.cat(w, dim=1), torch.stack(v, dim=1)


def fm_logit(w: Tensor, v: Tensor, bias: Tensor) -> Tensor:
    """fm/model.py:18-25 without the sigmoid: bias + sum_f w + 0.5*sum_d((sum_f v)^2 - sum_f v^2)."""
    first = torch.sum(w, dim=1, keepdim=True)
    second = 0.5 * torch.sum(torch.pow(torch.sum(v, dim=1), 2) - torch.sum(torch.pow(v, 2), dim=1),
                             dim=1, keepdim=True)
    return bias + first + second


def widedeep_split(features: Tensor, dims: Sequence[int], names: Sequence[str],
                   wide_names: Iterable[str]) -> Tuple[Tensor, Tensor]:
    """src/model/sort/widedeep/model.py:53-69."""
    wide_names = set(wide_names)
    wide, deep, s = [], [], 0
    for d, n in zip(dims, names):
        if n in wide_names:
            wide.append(features[:, s:s + 1])
            deep.append(features[:, s + 1:s + d])
        else:
            deep.append(features[:, s:s + d])
        s += d
    return torch.cat(wide, dim=1), torch.cat(deep, dim=1)


def mlp(x: Tensor, weights: Sequence[Tensor], biases: Sequence[Tensor],
        negative_slope: Optional[float] = None) -> Tensor:
    """src/model/model_utils/utils.py:6-17 (`MLP`): Linear, ReLU between layers, none
    after the last.  `negative_slope` selects LeakyReLU for the DSSM towers
    (src/model/recall/DSSM/model.py:26-44)."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w, b)
        if i < n - 1:
            x = F.relu(x) if negative_slope is None else F.leaky_relu(x, negative_slope)
    return x


def dcn_cross_v1(x: Tensor, ws: Sequence[Tensor], bs: Sequence[Tensor],
                 materialise: bool = True) -> Tensor:
    """src/model/sort/dcn/dcn_arch.py:14-30 (layer) and :63-70 (net).
    `materialise=True` keeps the reference's association `(x0 xl^T) w`
    (B x d x d intermediate); False uses the algebraically equal
    `x0 * (xl . w) + b + xl` for sizes where the intermediate cannot be held."""
    x0 = x
    for w, b in zip(ws, bs):
        if materialise:
            xl_ = x.unsqueeze(-1)
            x0_ = x0.unsqueeze(-1)
            cross = torch.matmul(torch.matmul(x0_, xl_.transpose(1, 2)), w)
            x = (cross + b + xl_).squeeze(-1)
        else:
            x = x0 * (x @ w) + b.view(1, -1) + x
    return x


def dcn_cross_v2(x: Tensor, Ws: Sequence[Tensor], bs: Sequence[Tensor]) -> Tensor:
    """dcn_arch.py:33-50 (`DCNv2Layer`) + :73-91 (`DCNv2Net`): x0 * Linear(xl) + xl, ReLU after each layer."""
    x0 = x
    for W, b in zip(Ws, bs):
        x = F.relu(x0 * F.linear(x, W, b) + x)
    return x


def bce(prob: Tensor, label: Tensor) -> Tensor:
    """`bceLoss` (sort/deep/model.py:32-33 and siblings): mean BCE on probabilities
    (torch clamps log at -100)."""
    return F.binary_cross_entropy(prob.view(-1), label.view(-1), reduction="mean")


# --------------------------------------------------------------------------- #
# Whole-model forwards keyed by the reference's state_dict names               #
# --------------------------------------------------------------------------- #

def _tables(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    pre, suf = "embedding_tables.", ".weight"
    return {k[len(pre):-len(suf)]: v for k, v in sd.items() if k.startswith(pre) and k.endswith(suf)}


def _mlp_params(sd: Dict[str, Tensor], prefix: str):
    ws, bs, i = [], [], 0
    while f"{prefix}.{i}.weight" in sd:
        ws.append(sd[f"{prefix}.{i}.weight"])
        bs.append(sd[f"{prefix}.{i}.bias"])
        i += 2  # nn.Sequential: Linear at even slots, activation at odd (utils.py:10-14)
    return ws, bs


def model_forward(kind: str, sd: Dict[str, Tensor], cfg: dict, batch: Dict[str, Tensor],
                  dcn_materialise: bool = True) -> Tensor:
    """forward(batch) of the reference model `kind` restated on a state_dict.

    lr       sort/lr/model.py:24-31
    fm       sort/fm/model.py:43-59, 18-26
    deep     sort/deep/model.py:36-43, 20-21
    widedeep sort/widedeep/model.py:48-69, 24-27
    dcn      sort/dcn/model.py:43-48, 25-29
    deepfm   (absent from the reference; documents/config_file_introduction.md:157-176)
             = sigmoid(FM logit over `deepfm_cfg.fm_feature_names` + MLP(all embeddings)),
             composed from fm/model.py:18-26 and utils.py:6-17.  Parity unpinned.
    """
    feats = cfg["features"]
    share = (cfg.get("embeddings", {}) or {}).get("share_emb_table_features", {}) or {}
    names = set(feats["user_feature_names"]) | set(feats["item_feature_names"])
    x, dims, fnames = embeddings_from_batch(_tables(sd), batch, names,
                                            feats.get("array_feature_names", []) or [], share)
    if kind == "lr":
        return torch.sigmoid(torch.sum(x, dim=1))
    if kind == "fm":
        w, v = fm_split(x, dims)
        return torch.sigmoid(fm_logit(w, v, sd["score_fc.bias"]))
    if kind == "deep":
        ws, bs = _mlp_params(sd, "score_fc.network.network")
        return torch.sigmoid(mlp(x, ws, bs))
    if kind == "widedeep":
        wide, deep = widedeep_split(x, dims, fnames, cfg["wide_and_deep_cfg"]["wide_feature_names"])
        ws, bs = _mlp_params(sd, "score_fc.deep_network.network")
        wide_out = torch.sum(wide, dim=1, keepdim=True) + sd["score_fc.bias"]
        return torch.sigmoid(wide_out + mlp(deep, ws, bs))
    if kind == "dcn":
        cw, cb, i = [], [], 0
        while f"score_fc.cross_net.cross_net.{i}.w" in sd:
            cw.append(sd[f"score_fc.cross_net.cross_net.{i}.w"])
            cb.append(sd[f"score_fc.cross_net.cross_net.{i}.b"])
            i += 1
        ws, bs = _mlp_params(sd, "score_fc.score_fc.network")
        xc = dcn_cross_v1(x, cw, cb, materialise=dcn_materialise)
        return torch.sigmoid(mlp(torch.cat([x, xc], dim=1), ws, bs))
    if kind == "deepfm":
        fm_names = set(cfg["deepfm_cfg"]["fm_feature_names"])
        fm_dim = int(cfg["deepfm_cfg"]["fm_dim"])
        w, v, s = [], [], 0
        for d, n in zip(dims, fnames):
            if n in fm_names:
                assert d == fm_dim + 1 or d == fm_dim, "deepfm: fm field width"
                w.append(x[:, s:s + 1])
                v.append(x[:, s + 1:s + d])
            s += d
        fm_out = fm_logit(torch.cat(w, dim=1), torch.stack(v, dim=1), sd["score_fc.bias"])
        ws, bs = _mlp_params(sd, "score_fc.deep_network.network")
        return torch.sigmoid(fm_out + mlp(x, ws, bs))
    raise ValueError(kind)


def train_loss(kind: str, sd: Dict[str, Tensor], cfg: dict, batch: Dict[str, Tensor], **kw) -> Tensor:
    """training_step minus logging/AUC (e.g. sort/deep/model.py:45-52): BCE(forward, label[:,0])."""
    return bce(model_forward(kind, sd, cfg, batch, **kw), batch["label"][:, 0])


def loss_and_grads(kind: str, sd: Dict[str, Tensor], cfg: dict, batch: Dict[str, Tensor], **kw):
    """fwd + loss + autograd backward; dense per-table gradients exactly as
    `aten::embedding_dense_backward` produces them for padding_idx=0."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    prob = model_forward(kind, leaf, cfg, batch, **kw)
    loss = bce(prob, batch["label"][:, 0])
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return prob.detach(), loss.detach(), grads


# --------------------------------------------------------------------------- #
# Optimizer (AdamW as configured by the reference)                             #
# --------------------------------------------------------------------------- #

def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, wd: float = 0.01):
    """torch.optim.AdamW single-tensor update as configured at sort/deep/model.py:55
    (betas (0.9,0.999), default eps 1e-8, default weight_decay 0.01); `step` is 1-based."""
    p = p * (1.0 - lr * wd)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def cosine_decay_lr(step: int, lrs: Sequence[float], milestones: Sequence[int]) -> float:
    """src/model/model_utils/lr_schedule.py:16-28 (`CosinDecayLR.get_lr`)."""
    import math
    if step < milestones[0]:
        return lrs[0]
    if step >= milestones[-1]:
        return lrs[-1]
    prog = (step - milestones[0]) / max(1, milestones[1] - milestones[0])
    return lrs[1] + (lrs[0] - lrs[1]) * 0.5 * (1.0 + math.cos(math.pi * prog))


# --------------------------------------------------------------------------- #
# Batch ingestion (pinned: tests/golden/ingest.npz)                            #
# --------------------------------------------------------------------------- #

def datareader_getitem(raw_line: str, cfg: dict, idx: int = 0) -> Dict[str, object]:
    """src/dataset/DataReader/data_reader.py:54-114 (`DataReader.__getitem__`) for one stripped line:
    "name:value ... \t labels"; sparse -> int, dense -> float, array -> int64[L] right-padded with 0 / truncated to the
    first L + float32 mask `<name>_mask`; names not in the config are ignored; label -> float32[n_labels]."""
    feats = cfg["features"]
    sparse = set(feats.get("sparse_feature_names") or [])
    dense = set(feats.get("dense_feature_names") or [])
    arrays = set(feats.get("array_feature_names") or [])
    amax = feats.get("array_max_length") or {}
    try:
        feature_part, label_part = raw_line.split("\t")
    except ValueError:
        raise ValueError(f"Line {idx} format error: missing tab separator between features and labels.")
    out: Dict[str, object] = {}
    for item in feature_part.split(" "):
        if ":" not in item:
            raise ValueError(f"Feature item format error: '{item}' does not contain ':' separator.")
        name, val = item.split(":", 1)
        if name in sparse:
            out[name] = int(val)
        elif name in dense:
            out[name] = float(val)
        elif name in arrays:
            max_len = amax.get(name)
            if max_len is None:
                raise ValueError(f"Max length for array feature '{name}' missing in config.")
            ids = [int(x) for x in val.split(",")] if val else []
            n = len(ids)
            if n < max_len:
                ids = ids + [0] * (max_len - n)
                mask = [1.0] * n + [0.0] * (max_len - n)
            else:
                ids = ids[:max_len]
                mask = [1.0] * max_len
            out[name] = torch.tensor(ids, dtype=torch.long)
            out[f"{name}_mask"] = torch.tensor(mask, dtype=torch.float32)
    out["label"] = torch.tensor([float(l) for l in label_part.strip().split(" ")], dtype=torch.float32)
    return out


def datareader_batch(lines: Sequence[str], cfg: dict, rows: Sequence[int]) -> Dict[str, Tensor]:
    """DataReader rows -> torch default collate (pl_dataloader.py:77-95): ints -> int64[B], floats -> float64[B],
    tensors stacked."""
    samples = [datareader_getitem(lines[i], cfg, i) for i in rows]
    out: Dict[str, Tensor] = {}
    for k in samples[0]:
        v0 = samples[0][k]
        if isinstance(v0, torch.Tensor):
            out[k] = torch.stack([s[k] for s in samples])
        elif isinstance(v0, int):
            out[k] = torch.tensor([s[k] for s in samples], dtype=torch.int64)
        else:
            out[k] = torch.tensor([s[k] for s in samples], dtype=torch.float64)
    return out


# --------------------------------------------------------------------------- #
# Validation scoring / metrics (pinned: tests/golden/valmetrics.npz)           #
# --------------------------------------------------------------------------- #

def validation_pairs(user_ids: Tensor, scores: Tensor, label: Tensor):
    """base_model.py:320-330 (`validation_step`): `zip(user_id.view(-1), scores.view(-1), label.view(-1))`.
    NOTE the reference flattens the WHOLE [B, n_labels] label tensor, so zip pairs sample i with
    `label.view(-1)[i]` (= label[i // n, i % n]), not with label[i, 0]; restated as is."""
    u = user_ids.view(-1).cpu().numpy()
    s = scores.view(-1).cpu().numpy()
    l = label.view(-1).cpu().numpy()
    n = min(len(u), len(s), len(l))
    return u[:n], s[:n], l[:n]


def roc_auc(labels, preds) -> float:
    """sklearn.metrics.roc_auc_score for binary labels (the call at base_model.py:379,445): area under the ROC
    curve == Mann-Whitney statistic with ties counted one half."""
    import numpy as np
    y = np.asarray(labels, dtype=np.float64)
    p = np.asarray(preds, dtype=np.float64)
    n_pos, n_neg = float((y == 1).sum()), float((y != 1).sum())
    order = np.argsort(p, kind="mergesort")
    ps, ys = p[order], y[order]
    ranks = np.empty(len(p), dtype=np.float64)
    i = 0
    while i < len(ps):          # average ranks inside tie groups
        j = i
        while j + 1 < len(ps) and ps[j + 1] == ps[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    return float((ranks[ys == 1].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def validation_metrics(user_ids, preds, labels, k: int = 10, user_in_train_set=None) -> Dict[str, Dict[str, float]]:
    """base_model.py:333-478 (`on_validation_epoch_end`): users in first-appearance order (dict insertion, :327-330);
    per user AUC when both classes occur (:376-383), top-k by score descending with a STABLE sort (:387), HR / NDCG /
    MRR @k (0 for users without positives, :393-401); warm / cold split by membership of uid or str(uid) in
    `user_in_train_set` (:362-366); overall AUC + log-loss per group (:443-461, float32 arithmetic as numpy does it)."""
    import numpy as np
    per_user: Dict[object, list] = {}
    for uid, sc, lb in zip(user_ids, preds, labels):
        per_user.setdefault(uid, []).append((sc, lb))
    groups = {g: {"preds": [], "labels": [], "auc": [], "ndcg": [], "hr": [], "mrr": []} for g in ("all", "warm", "cold")}
    for uid, items in per_user.items():
        ps = [x[0] for x in items]
        ls = [x[1] for x in items]
        cold = bool(user_in_train_set) and uid not in user_in_train_set and str(uid) not in user_in_train_set
        tgt = groups["cold" if cold else "warm"]
        for g in (groups["all"], tgt):
            g["preds"].extend(ps)
            g["labels"].extend(ls)
        if len(set(ls)) > 1:
            a = roc_auc(ls, ps)
            groups["all"]["auc"].append(a)
            tgt["auc"].append(a)
        top = sorted(items, key=lambda x: x[0], reverse=True)[:k]
        n_pos = sum(1 for x in items if x[1] == 1)
        if n_pos == 0:
            for g in (groups["all"], tgt):
                g["hr"].append(0.0); g["ndcg"].append(0.0); g["mrr"].append(0.0)
            continue
        hr = 1.0 if any(x[1] == 1 for x in top) else 0.0
        dcg = sum(1.0 / np.log2(r + 1) for r, (_, lb) in enumerate(top, start=1) if lb == 1)
        idcg = sum(1.0 / np.log2(r + 1) for r in range(1, min(n_pos, k) + 1))
        ndcg = dcg / idcg if idcg > 0 else 0.0
        mrr = 0.0
        for r, (_, lb) in enumerate(top, start=1):
            if lb == 1:
                mrr = 1.0 / r
                break
        for g in (groups["all"], tgt):
            g["hr"].append(hr); g["ndcg"].append(ndcg); g["mrr"].append(mrr)

    def mean(l):
        return float(np.mean(l)) if l else 0.0

    def auc_logloss(ps, ls):
        auc, ll = 0.0, 0.0
        if len(ps) > 0:
            if len(set(ls)) > 1:
                auc = roc_auc(ls, ps)
            eps = 1e-15
            pa = np.clip(ps, eps, 1 - eps)
            la = np.array(ls)
            ll = float(-np.mean(la * np.log(pa) + (1 - la) * np.log(1 - pa)))
        return auc, ll

    out = {}
    for name, g in (("Overall", "all"), ("Warm_Start", "warm"), ("Cold_Start", "cold")):
        G = groups[g]
        auc, ll = auc_logloss(G["preds"], G["labels"])
        out[name] = {"AUC": auc, "LogLoss": ll, "GAUC": mean(G["auc"]), f"NDCG@{k}": mean(G["ndcg"]), f"HR@{k}": mean(G["hr"]),
                     f"MRR@{k}": mean(G["mrr"])}
        if name != "Overall":
            out[name]["User_Count"] = len(G["hr"])
    return out


# --------------------------------------------------------------------------- #
# DSSM (pinned: tests/golden/dssm.npz)                                         #
# --------------------------------------------------------------------------- #

def dssm_tower(x: Tensor, weights: Sequence[Tensor], biases: Sequence[Tensor]) -> Tensor:
    """recall/DSSM/model.py:26-44: Linear-LeakyReLU(0.2) x3, Linear."""
    return mlp(x, weights, biases, negative_slope=0.2)


def dssm_forward(user_x: Tensor, item_x: Tensor, user_params, item_params,
                 neg_perms: Sequence[Tensor]):
    """recall/DSSM/model.py:51-73.  The `torch.randperm` draws (:63) are passed in."""
    u = dssm_tower(user_x, *user_params)
    it = dssm_tower(item_x, *item_params)
    neg = torch.stack([it[p] for p in neg_perms], dim=1)
    return F.normalize(u, p=2, dim=1), F.normalize(it, p=2, dim=1), F.normalize(neg, p=2, dim=-1)


def infonce_loss(user_emb: Tensor, pos: Tensor, neg: Tensor, temperature: float = 0.1,
                 mask: Optional[Tensor] = None) -> Tensor:
    """recall/DSSM/model.py:92-110."""
    pos_s = torch.sum(user_emb * pos, dim=1) / temperature
    neg_s = torch.bmm(user_emb.unsqueeze(1), neg.permute(0, 2, 1)).squeeze(1) / temperature
    logits = torch.cat([pos_s.unsqueeze(1), neg_s], dim=1)
    labels = torch.zeros(user_emb.size(0), dtype=torch.long)
    losses = F.cross_entropy(logits, labels, reduction="none")
    if mask is not None:
        losses = losses * mask
    return losses.mean()


def hit_rate_filtered(ranked_ids: Sequence[Sequence[int]], histories: Sequence[Iterable[int]], targets: Sequence[int],
                      k: int = 10) -> float:
    """recall/DSSM/model.py:183-229 (`hit_rate`), one query per user: the searcher returned `k + len(history)`
    candidates (:207-209); candidates the user already interacted with are dropped, the first k survivors kept
    (:213-220), a hit is the target among them (:221-223); hit rate = hits / users (:226).
    (The id re-mapping dictionaries of the MovieLens-era code, :204-205,:212,:216, are identity here.)"""
    hits, n = 0, 0
    for I, hist, tgt in zip(ranked_ids, histories, targets):
        hist = set(hist)
        kept = []
        for item in list(I)[: k + len(hist)]:
            if item not in hist:
                kept.append(item)
            if len(kept) >= k:
                break
        hits += int(tgt in kept)
        n += 1
    return hits / n if n > 0 else 0


# --------------------------------------------------------------------------- #
# Exact inner-product top-k (faiss.IndexFlatIP restated — parity unpinned)     #
# --------------------------------------------------------------------------- #

def topk_ip(queries: Tensor, corpus: Tensor, k: int, normalize: bool = False,
            chunk: int = 256) -> Tuple[Tensor, Tensor]:
    """faiss==1.7.4 `IndexFlatIP.add/search` as called at recall/DSSM/model.py:209,249-251
    and model_utils/TopKSearcher.py:34-47,73-77: exact fp32 inner products, the k
    largest per query in descending order, ids = insertion positions (int64),
    -1 / -inf padding when k > N (faiss convention: -3.4e38 scores).  Tie order is
    pinned by north_star to (score desc, id asc) => stable descending sort."""
    q = queries.float()
    c = corpus.float()
    if normalize:  # faiss.normalize_L2 (TopKSearcher.py:30-31,69-70)
        q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)
        c = c / c.norm(dim=1, keepdim=True).clamp_min(1e-30)
    n = c.shape[0]
    kk = min(k, n)
    out_s = torch.full((q.shape[0], k), -3.4028234663852886e38, dtype=torch.float32)
    out_i = torch.full((q.shape[0], k), -1, dtype=torch.int64)
    if kk == 0:
        return out_s, out_i
    # faiss scores with an fp32 BLAS whose summation order is unspecified, so rows whose scores differ by
    # less than ~1e-6 come out in an implementation-defined order.  The contract here pins the order with
    # the inner product of the fp32 inputs accumulated in fp64 (products of fp32 values are exact in fp64),
    # ties -> lower id; the returned score is that value rounded to fp32 (within 1e-6 of faiss').
    qd, cd = q.double(), c.double()
    for s in range(0, q.shape[0], chunk):
        sc = qd[s:s + chunk] @ cd.T
        vals, idx = torch.sort(sc, dim=1, descending=True, stable=True)
        out_s[s:s + chunk, :kk] = vals[:, :kk].float()
        out_i[s:s + chunk, :kk] = idx[:, :kk]
    return out_s, out_i


def topk_merge(scores: Sequence[Tensor], ids: Sequence[Tensor], k: int) -> Tuple[Tensor, Tensor]:
    """Merge per-shard top-k lists into the global top-k under (score desc, id asc)."""
    s = torch.cat(list(scores), dim=1)
    i = torch.cat(list(ids), dim=1)
    # order by id asc first, then stable sort by score desc => (score desc, id asc)
    o = torch.argsort(i, dim=1, stable=True)
    s, i = torch.gather(s, 1, o), torch.gather(i, 1, o)
    o = torch.argsort(s, dim=1, descending=True, stable=True)
    return torch.gather(s, 1, o)[:, :k], torch.gather(i, 1, o)[:, :k]


# --------------------------------------------------------------------------- #
# bf16 restatement of the MLP (same reference lines + the rounding points of    #
# the tensor-core path) — used to check each forward/backward step exactly      #
# --------------------------------------------------------------------------- #

def bf16(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(torch.float64)


def mlp_bf16_layer(a: Tensor, w: Tensor, b: Tensor, last: bool, negative_slope: Optional[float]) -> Tensor:
    """One utils.py:11 Linear (+ activation) with bf16 operands, wide accumulation, bf16 output."""
    z = bf16(a) @ bf16(w).T + b.double()
    if last:
        return z
    s = 0.0 if negative_slope is None else negative_slope
    return bf16(torch.where(z > 0, z, z * s).float())


def mlp_bf16_dx_step(dz: Tensor, w: Tensor, a_in: Tensor, negative_slope: Optional[float]) -> Tensor:
    """dz_{l-1} = bf16((dz_l W_l) * act'(a_l)) — autograd of Linear+ReLU/LeakyReLU with bf16 operands."""
    s = 0.0 if negative_slope is None else negative_slope
    return bf16(((bf16(dz) @ bf16(w)) * torch.where(a_in > 0, 1.0, s)).float())
