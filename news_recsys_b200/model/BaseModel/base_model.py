"""BaseModel — host-side mirror of the reference's `src/model/BaseModel/base_model.py`
for the embedding part of the hot path (config loading :69-106, table construction
:141-166, `get_feature_embedding` :262-271, `array_feature_pooling` :273-282,
`get_embeddings_from_batch` :284-308), with the arithmetic done by libnrx (K1/K3).

Same constructor (`Model(config_path)`), attributes and `state_dict` keys
(`embedding_tables.<table>.weight`), so reference checkpoints load with
`strict=True`.  The Lightning validation/metrics harness (:181-256, :320-528) is
out of scope; when `lightning` is importable the class
derives from `LightningModule` so a Trainer can still drive it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Set, Tuple

import torch
import torch.nn as nn

from ... import ops
from ..._lib import NrxError
from ...config import load_config, to_container

try:  # optional: drop into a Lightning Trainer when it exists
    import lightning as _L  # type: ignore
    _Base = _L.LightningModule
except Exception:  # pragma: no cover - lightning is absent in the build image
    _Base = nn.Module


class BaseModel(_Base):
    def __init__(self, config_path):
        super().__init__()
        self._load_config(config_path)
        self.item_input_dim = self._calculate_input_dim(self.item_feature_names)
        self.user_input_dim = self._calculate_input_dim(self.user_feature_names)
        self.embedding_tables = self._build_embedding_tables()
        self._table_ids = {name: i for i, name in enumerate(self.embedding_tables.keys())}
        if len(self._table_ids) > ops.L.NRX_MAX_TABLES:
            raise NrxError(f"{len(self._table_ids)} embedding tables > {ops.L.NRX_MAX_TABLES}")
        self.feature_id_mapper = None

    # ---- config (base_model.py:69-106) -------------------------------------------------
    def _load_config(self, config_path):
        self.config = load_config(config_path)
        paths = self.config.get("paths", {}) or {}
        self.out_basedir = paths.get("out_basedir", "")
        self.user_history_path = paths.get("user_history_path", "")
        f = self.config.get("features", {}) or {}
        self.sparse_feature_names: Set[str] = set(f.get("sparse_feature_names", []) or [])
        self.dense_feature_names: Set[str] = set(f.get("dense_feature_names", []) or [])
        self.array_feature_names: Set[str] = set(f.get("array_feature_names", []) or [])
        self.item_feature_names: Set[str] = set(f.get("item_feature_names", []) or [])
        self.user_feature_names: Set[str] = set(f.get("user_feature_names", []) or [])
        self.array_max_length: Dict[str, int] = to_container(f.get("array_max_length", {}) or {})
        e = self.config.get("embeddings", {}) or {}
        self.embedding_size: Dict[str, int] = to_container(e.get("embedding_size", {}) or {})
        self.embedding_table_size: Dict[str, int] = to_container(e.get("embedding_table_size", {}) or {})
        self.share_emb_table_features: Dict[str, str] = to_container(e.get("share_emb_table_features", {}) or {})
        self.dataset_cfg = self.config.get("dataset", {})
        self.train_hparams = self.config.get("train_hparams", {})

    def _get_emb_feature_name(self, feature_name: str) -> str:
        """base_model.py:119-122."""
        return self.share_emb_table_features.get(feature_name, feature_name)

    def _calculate_input_dim(self, feature_names: Set[str]) -> int:
        """base_model.py:124-139 (dense features fail here exactly like the reference, :94,:129)."""
        total = 0
        for fname in feature_names:
            if fname in self.dense_feature_names:
                total += self.dense_feature_dim  # AttributeError, as in the reference
            else:
                dim = self.embedding_size.get(self._get_emb_feature_name(fname))
                total += 8 if dim is None else dim
        return total

    def _build_embedding_tables(self) -> nn.ModuleDict:
        """base_model.py:141-166: one nn.Embedding(size, dim, padding_idx=0) per distinct table
        (fp32, N(0,1), row 0 zero).  Only `.weight` is used; the gather runs in K1."""
        tables = nn.ModuleDict()
        # sorted: the reference iterates a Python set here, whose order (hence the RNG draw order of the
        # N(0,1) init and the ModuleDict order) changes from process to process; ranks must agree
        for fname in sorted(self.sparse_feature_names.union(self.array_feature_names)):
            t = self._get_emb_feature_name(fname)
            if t in tables:
                continue
            size, dim = self.embedding_table_size.get(t), self.embedding_size.get(t)
            if size is None or dim is None:
                continue  # the reference logs an error and skips (:158-161)
            tables[t] = nn.Embedding(size, dim, padding_idx=0)
        return tables

    # ---- embedding lookups ---------------------------------------------------------------
    def _specs_for(self, names: List[str], batch) -> Tuple[List[ops.FeatSpec], List[int]]:
        specs, dims, col = [], [], 0
        for fname in names:
            if fname not in batch:
                continue  # the reference silently skips (:291-293)
            if fname in self.dense_feature_names:
                raise NrxError("dense features are not supported on this path (the reference crashes on them too)")
            t = self._get_emb_feature_name(fname)
            if t not in self.embedding_tables:
                raise ValueError(f"Embedding table not found for {fname} (mapped to {t})")  # :268-269
            dim = self.embedding_tables[t].weight.shape[1]
            is_arr = fname in self.array_feature_names
            specs.append(ops.FeatSpec(fname, t, self._table_ids[t], dim,
                                      batch[fname].shape[1] if is_arr else 1, is_arr, col))
            dims.append(dim)
            col += dim
        return specs, dims

    def _weights(self) -> Dict[str, torch.Tensor]:
        return {k: m.weight for k, m in self.embedding_tables.items()}

    # ---- out-of-table ids: the reference's nn.Embedding raises (base_model.py:271); K1 raises a status bit -----------
    def _id_status(self, device) -> torch.Tensor:
        """Persistent int32 status word the gather kernels OR into when an id is outside [0, rows)."""
        st = getattr(self, "_id_status_buf", None)
        if st is None or st.device != device:
            st = torch.zeros(1, dtype=torch.int32, device=device)
            self._id_status_buf = st
        return st

    def check_ids(self):
        """Synchronises and raises if any forward since the last check saw an id outside its embedding table — what the
        reference's nn.Embedding reports as an IndexError / device assert.  Cheap to call every N steps or with the loss
        read-back; the flag is sticky until checked."""
        st = getattr(self, "_id_status_buf", None)
        if st is not None and int(st.item()) != 0:
            st.zero_()
            raise IndexError("a feature id outside its embedding table reached the gather kernel "
                             "(vocabulary / config mismatch; the reference's nn.Embedding raises here, base_model.py:271)")

    def bind_features(self, batch, feature_names, want_inv_den=True):
        """NrxFeat[] for `feature_names` (sorted) of this batch -> (binding, dims, names, out_dim)."""
        names = sorted(list(feature_names))
        specs, dims = self._specs_for(names, batch)
        if not specs:
            return None, [], names, 0
        fb = ops.FeatBinding(specs, self._weights(), batch, want_inv_den=want_inv_den)
        return fb, dims, names, sum(dims)

    def get_embeddings_from_batch(self, batch, feature_names) -> Tuple[torch.Tensor, List[int], List[str]]:
        """base_model.py:284-308 -> (float32[B, ΣD] in sorted-name column order, dims, sorted names)."""
        fb, dims, names, out_dim = self.bind_features(batch, feature_names)
        if fb is None:
            dev = next(self.parameters()).device
            return torch.tensor([]).to(dev), [], []
        tnames = list(self.embedding_tables.keys())
        ws = [self.embedding_tables[t].weight for t in tnames]
        fb.status = self._id_status(fb.device)
        if torch.is_grad_enabled() and any(w.requires_grad for w in ws):
            x = ops.EmbedPoolFn.apply(fb, out_dim, tnames, *ws)
        else:
            x = ops.embed_pool_fwd(fb, out_dim, status=fb.status)
        return x, dims, names

    def get_feature_embedding(self, feature_name: str, feature_value: torch.Tensor) -> torch.Tensor:
        """base_model.py:262-271, inference-only helper: [B] -> [B, D], [B, L] -> [B, L, D]."""
        if feature_name in self.dense_feature_names:
            return feature_value.float().unsqueeze(1)
        t = self._get_emb_feature_name(feature_name)
        if t not in self.embedding_tables:
            raise ValueError(f"Embedding table not found for {feature_name} (mapped to {t})")
        w = self.embedding_tables[t].weight
        flat = feature_value.reshape(-1)
        spec = ops.FeatSpec("_v", t, 0, w.shape[1], 1, False, 0)
        fb = ops.FeatBinding([spec], {t: w.detach()}, {"_v": flat}, want_inv_den=False)
        out = ops.embed_pool_fwd(fb, w.shape[1], status=self._id_status(w.device))
        return out.view(*feature_value.shape, w.shape[1])

    def array_feature_pooling(self, embedding: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """base_model.py:273-282 on an already gathered [B, L, D] tensor (inference-only helper;
        the training path pools inside K1 and never materialises [B, L, D])."""
        B, Lh, D = embedding.shape
        tbl = embedding.detach().contiguous().view(B * Lh, D)
        ids = torch.arange(B * Lh, device=embedding.device, dtype=torch.int64).view(B, Lh)
        spec = ops.FeatSpec("_a", "_a", 0, D, Lh, True, 0)
        b = {"_a": ids}
        if mask is not None:
            b["_a_mask"] = mask
        fb = ops.FeatBinding([spec], {"_a": tbl}, b, want_inv_den=False)
        return ops.embed_pool_fwd(fb, D)

    # ---- training surface shared by the sort models (e.g. deep/model.py:32-33,45-65) ------
    def bceLoss(self, preds, labels):
        return ops.BceFn.apply(preds.view(-1), labels.view(-1))

    def training_step(self, batch, batch_idx=0):
        """fwd + BCE(label[:,0]).  The reference also calls sklearn's roc_auc_score here every
        step (a D2H sync that raises on single-class batches, deep/model.py:49); metrics are
        out of scope for the accelerated path."""
        scores = self.forward(batch)
        return self.bceLoss(scores, batch["label"][:, 0])

    def configure_optimizers(self):
        from ..model_utils.lr_schedule import CosinDecayLR
        hp = self.train_hparams
        optimizer = torch.optim.AdamW(self.parameters(), lr=hp.lr, betas=(0.9, 0.999))
        sched = CosinDecayLR(optimizer, lrs=[hp.lr, hp.min_lr], milestones=hp.lr_milestones)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": sched, "interval": "step", "frequency": 1}}

    @torch.no_grad()
    def inference(self, batch):
        return self.forward(batch)

    def forward(self, x):
        raise NotImplementedError("Subclasses must implement forward()")

    # ---- validation scoring / metrics on the device (reference :320-478) ---------------------------
    def validation_step(self, batch, batch_idx=0):
        from ...metrics import ValidationMetrics
        if getattr(self, "_val_metrics", None) is None:
            self._val_metrics = ValidationMetrics(k=10, user_in_train_set=getattr(self, "user_in_train_set", None))
        self._val_metrics.update(batch["user_id"], self.inference(batch), batch["label"])

    def on_validation_epoch_end(self):
        """Returns the dict the reference builds and prints (Overall / Warm_Start / Cold_Start)."""
        from ...metrics import ValidationMetrics
        m = getattr(self, "_val_metrics", None) or ValidationMetrics(k=10)
        results = m.compute()
        self._val_metrics = None
        return results

    def load_model(self, path: str):
        """base_model.py:531-536."""
        sd = torch.load(path, map_location="cpu")
        if "state_dict" in sd:
            sd = sd["state_dict"]
        self.load_state_dict(sd, strict=True)
