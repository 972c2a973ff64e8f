"""DSSM — restatement of reference src/model/recall/DSSM/model.py on the B200 kernels.

The reference file is not importable as shipped (stale `BaseModel.*` imports :4,11,13, calls to the
non-existent `get_features_embedding` :151,168), so this mirrors its INTENT line by line:
  towers            :26-44   Linear(in,128)-LeakyReLU(0.2)-Linear(128,128)-LReLU-Linear(128,64)-LReLU-Linear(64,out)
                             (out = 16 in the reference; BASELINE config 4 uses 128 -> `hparams['out_dim']`)
  forward           :51-73   user/item towers, in-batch `randperm` negatives x negative_sample_rate, L2 normalise
  infoNCE_loss      :92-110  cross-entropy over [pos, negs] / temperature, masked by label[:,1]
  get_*_embedding   :148-180 per-feature gather (+ masked mean for array features), concatenated
  on_train_epoch_end:230-254 item tower over the corpus -> IndexFlatIP -> search  => build_item_index / retrieve
Embeddings run on K1/K3, both towers on the fused tcgen05 tower (K4, LeakyReLU), scoring + top-k on K6.
The normalise / InfoNCE tail works on [B, out] tensors and stays in PyTorch (outside the four hot ops of
north_star).  The reference concatenates features in Python-set order (:150,:167); this mirror uses sorted
order, the order every other model of the reference uses (base_model.py:286)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import ops
from ....retrieval import TopkIndex
from ...BaseModel.base_model import BaseModel


def _tower(in_dim, out_dim):
    return nn.Sequential(nn.Linear(in_dim, 128), nn.LeakyReLU(0.2), nn.Linear(128, 128), nn.LeakyReLU(0.2),
                         nn.Linear(128, 64), nn.LeakyReLU(0.2), nn.Linear(64, out_dim))


def _run_tower(seq, x):
    lin = [m for m in seq if isinstance(m, nn.Linear)]
    return ops.TowerFn.apply(x, 0.2, len(lin), *[m.weight for m in lin], *[m.bias for m in lin])


class DSSM(BaseModel):
    def __init__(self, config_path, dataloaders=None, hparams=None):
        super().__init__(config_path)
        self.hparams_ = dict(hparams or {})
        self.hparams_.setdefault("negative_sample_rate", 3)
        out_dim = int(self.hparams_.get("out_dim", 16))
        self.user_fc = _tower(self.user_input_dim, out_dim)
        self.item_fc = _tower(self.item_input_dim, out_dim)
        dataloaders = dataloaders or {}
        self.movies_dataloader = dataloaders.get("movies_dataloader", None)
        self.val_dataloader_ = dataloaders.get("val_dataloader", None)
        self.index = None

    def get_user_embedding(self, batch):
        x, _, _ = self.get_embeddings_from_batch(batch, self.user_feature_names)
        return x

    def get_item_embedding(self, batch):
        x, _, _ = self.get_embeddings_from_batch(batch, self.item_feature_names)
        return x

    def user_tower(self, batch):
        return _run_tower(self.user_fc, self.get_user_embedding(batch))

    def item_tower(self, batch):
        return _run_tower(self.item_fc, self.get_item_embedding(batch))

    def forward(self, x, neg_perms=None):
        user_emb = self.user_tower(x)
        item_emb = self.item_tower(x)
        B = item_emb.size(0)
        if neg_perms is None:  # the reference draws torch.randperm on the CPU generator (:63)
            neg_perms = [torch.randperm(B) for _ in range(self.hparams_["negative_sample_rate"])]
        neg = torch.stack([item_emb[p.to(item_emb.device)] for p in neg_perms], dim=1)
        return F.normalize(user_emb, p=2, dim=1), F.normalize(item_emb, p=2, dim=1), F.normalize(neg, p=2, dim=-1)

    def infoNCE_loss(self, user_emb, pos_item_emb, neg_item_emb, temperature=0.1, mask=None):
        pos = torch.sum(user_emb * pos_item_emb, dim=1) / temperature
        neg = torch.bmm(user_emb.unsqueeze(1), neg_item_emb.permute(0, 2, 1)).squeeze(1) / temperature
        logits = torch.cat([pos.unsqueeze(1), neg], dim=1)
        labels = torch.zeros(user_emb.size(0), dtype=torch.long, device=user_emb.device)
        losses = F.cross_entropy(logits, labels, reduction="none")
        if mask is not None:
            losses = losses * mask
        return losses.mean()

    def triplet_loss(self, user_emb, pos_item_emb, neg_item_emb, margin=1.0, mask=None):
        n = neg_item_emb.size(1)
        pos = torch.sum(user_emb * pos_item_emb, dim=1) * n
        neg = torch.bmm(user_emb.unsqueeze(1), neg_item_emb.permute(0, 2, 1)).squeeze(1).sum(dim=1)
        losses = F.relu(margin - pos + neg)
        if mask is not None:
            losses = losses * mask
        return losses.mean()

    def training_step(self, batch, batch_idx=0, neg_perms=None, fused=True):
        """The reference's step (:112-126): forward (:51-73) + infoNCE_loss (:92-110) with mask = label[:, 1].
        fused=True (default): the normalisations, the negative gather, the loss and their whole backward run as
        ops.InfoNCEFn (nrx_dssm_infonce, two launches) on the raw tower outputs; fused=False keeps the reference's
        operator-by-operator form (forward() / infoNCE_loss() remain available with the reference's signatures)."""
        if not fused:
            u, it, neg = self.forward(batch, neg_perms)
            return self.infoNCE_loss(u, it, neg, mask=batch["label"][:, 1])
        user_raw = self.user_tower(batch)
        item_raw = self.item_tower(batch)
        B = item_raw.size(0)
        if neg_perms is None:  # the reference draws torch.randperm on the CPU generator (:63)
            neg_perms = [torch.randperm(B) for _ in range(self.hparams_["negative_sample_rate"])]
        perms = [p.to(item_raw.device) for p in neg_perms]
        return ops.InfoNCEFn.apply(user_raw, item_raw, batch["label"][:, 1], 0.1, *perms)

    # ---- retrieval (on_train_epoch_end :230-254, hit_rate :209) ---------------------------------------
    @torch.no_grad()
    def build_item_index(self, item_batches, id_base=0, item_key="item_id", group=None, n_total=None):
        """Item tower over the corpus (list of batches with the item features) -> normalised vectors -> index.

        Sharded refresh (BASELINE config 4): with `group` (a torch.distributed group of > 1 ranks on one node) every rank
        passes only ITS contiguous share of the corpus — rows shard_range(n_total, rank, world) in rank order —, runs the
        item tower over that share and keeps it behind parallel.ShardedTopk (one peer-memory search per query batch, the
        result identical on every rank and to a single index over the whole corpus); the item-id columns are all-gathered
        once so that every rank can map global positions to item ids.

        The reference keeps `idx_item_emb_dic`: corpus position -> item id, filled from the id column of every corpus
        batch (`on_train_epoch_end`, recall/DSSM/model.py:236-247) and maps search results through it before the history
        filter / hit test (:212-215).  Same here: the id column (`item_key`) of the corpus batches is collected into
        `self.index_item_ids` (device int64 [N]) and `retrieve_items()` maps ranked positions through it.  Batches
        without that column fall back to position + id_base (the identity mapping of an id-ordered corpus)."""
        embs = [ops.l2_normalize(self.item_tower(b)) for b in item_batches]
        self.all_item_embeddings = torch.cat(embs, dim=0)
        dev = self.all_item_embeddings.device
        n_local = self.all_item_embeddings.shape[0]
        if all(item_key in b for b in item_batches):
            ids = torch.cat([b[item_key].reshape(-1).to(dev, torch.int64) for b in item_batches])
        else:
            ids = None
        if ids is not None and ids.numel() != n_local:
            raise ValueError("build_item_index: the item id column and the corpus disagree on the number of items")
        world = 1
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
        if world > 1:
            import torch.distributed as dist
            from ....parallel import ShardedTopk, shard_range
            counts = torch.zeros(world, dtype=torch.int64, device=dev)
            counts[dist.get_rank(group)] = n_local
            dist.all_reduce(counts, group=group)
            total = int(counts.sum().item())
            if n_total is not None and int(n_total) != total:
                raise ValueError(f"build_item_index: the shards hold {total} items, n_total says {n_total}")
            lo, hi = shard_range(total, dist.get_rank(group), world)
            if hi - lo != n_local:
                raise ValueError(f"build_item_index: this rank must hold rows [{lo}, {hi}) of the corpus ({hi - lo} items), "
                                 f"got {n_local} (shards are the contiguous ranges of parallel.shard_range)")
            if getattr(self.index, "close", None) is not None:
                self.index.close()          # the previous epoch's sharded index: peer buffers are not garbage-collected
            self.index = ShardedTopk(self.all_item_embeddings, total, group=group)
            if ids is None:
                self.index_item_ids = torch.arange(total, device=dev) + int(id_base)
            else:   # shard sizes differ by at most one: pad to the largest, gather, cut the padding out again
                cap = int(counts.max().item())
                padded = torch.full((cap,), -1, dtype=torch.int64, device=dev)
                padded[:n_local] = ids
                allp = torch.empty((world, cap), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(allp, padded, group=group)
                self.index_item_ids = torch.cat([allp[r, : int(counts[r])] for r in range(world)])
            return self.index
        self.index = TopkIndex(self.all_item_embeddings, id_base=0)
        self.index_item_ids = ids if ids is not None else torch.arange(n_local, device=dev) + int(id_base)
        return self.index

    @staticmethod
    def filter_history_hits(ranked_ids, history, targets, k):
        """The filtering tail of the reference's `hit_rate` (:207-223) for a whole batch of users at once.
        ranked_ids [Q, >= k + H] corpus ids best first, history [Q, H] item ids already interacted with (padded with
        -1), targets [Q].  Per user only the first k + |history| candidates count (:207-209), history items are
        dropped, the first k survivors kept, hit = target among them.  Returns a bool [Q].  Tensor ops on
        [Q, k + H] only (device-agnostic)."""
        Q, W = ranked_ids.shape
        hist_len = (history >= 0).sum(dim=1, keepdim=True)
        pos = torch.arange(W, device=ranked_ids.device).unsqueeze(0)
        in_window = pos < (k + hist_len)
        seen = (ranked_ids.unsqueeze(2) == history.unsqueeze(1)).any(dim=2) & (ranked_ids >= 0)
        alive = in_window & ~seen
        rank = torch.cumsum(alive.to(torch.int64), dim=1)          # 1-based rank among the survivors
        kept = alive & (rank <= k)
        return (kept & (ranked_ids == targets.view(-1, 1))).any(dim=1)

    @torch.no_grad()
    def hit_rate(self, batches, k=10, history_key="user_history", target_key="item_id"):
        """Batched `hit_rate` (:183-229): for every user of every batch search k + H candidates (H = padded history
        length), drop the history, keep k, count targets hit.  The reference does this one user per call on the CPU."""
        if self.index is None:
            raise ValueError("Index not initialized. Call build_item_index first.")
        hits, n = 0, 0
        for b in batches:
            hist = b[history_key].clone()
            if history_key + "_mask" in b:
                hist[b[history_key + "_mask"] == 0] = -1
            hist[hist == 0] = -1                                    # id 0 is padding
            _, ids = self.retrieve_items(b, k + hist.shape[1])     # ITEM ids (positions mapped through the corpus id column)
            hits += int(self.filter_history_hits(ids, hist, b[target_key], k).sum())
            n += ids.shape[0]
        return hits / n if n > 0 else 0

    @torch.no_grad()
    def retrieve_items(self, batch, k):
        """(scores [B,k], ITEM ids [B,k]): corpus positions of `retrieve` mapped through the id column collected by
        build_item_index (the reference's idx_item_emb_dic, :212-215); padding (-1) stays -1."""
        s, pos = self.retrieve(batch, k)
        ids = torch.where(pos >= 0, self.index_item_ids[pos.clamp_min(0)], pos)
        return s, ids

    @torch.no_grad()
    def retrieve(self, batch, k):
        """(scores [B,k], corpus positions [B,k]) ordered by (inner product desc, position asc)."""
        if self.index is None:
            raise ValueError("Index not initialized. Call build_item_index first.")
        return self.index.search(ops.l2_normalize(self.user_tower(batch)), k)
