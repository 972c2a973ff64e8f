"""Host-side operator layer: thin wrappers over the C ABI (include/nrx.h) plus the
torch.autograd.Function glue that lets the reference-shaped modules train with
`loss.backward()` exactly like the Lightning modules they replace.

Nothing here computes on the CPU and nothing falls back to PyTorch ops: every
function requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise L.NrxError(f"{what}: expected a CUDA tensor (the hot path has no CPU fallback)")


# --------------------------------------------------------------------------- #
# Feature descriptors                                                          #
# --------------------------------------------------------------------------- #

@dataclass
class FeatSpec:
    """One feature of get_embeddings_from_batch (reference base_model.py:284-308)."""
    name: str
    table: str        # after share_emb_table_features aliasing (base_model.py:119-122)
    table_id: int
    dim: int
    L: int            # 1 = sparse
    is_array: bool
    out_col: int


class FeatBinding:
    """NrxFeat[] for one batch; keeps every tensor it points at alive."""

    def __init__(self, specs: Sequence[FeatSpec], tables: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor],
                 want_inv_den: bool = True):
        self.specs = list(specs)
        n = len(self.specs)
        if n == 0 or n > L.NRX_MAX_FEATS:
            raise L.NrxError(f"{n} features outside [1,{L.NRX_MAX_FEATS}]")
        self.arr = (L.NrxFeat * n)()
        self.keep: List[torch.Tensor] = []
        self.inv_den: Dict[str, torch.Tensor] = {}
        self.B = None
        dev = None
        for i, s in enumerate(self.specs):
            w = tables[s.table]
            _require_cuda(w, f"table {s.table}")
            dev = w.device
            if w.dtype != torch.float32 or w.dim() != 2 or w.stride(1) != 1:
                raise L.NrxError(f"table {s.table}: need fp32 [rows, dim] with unit column stride")
            idx = batch[s.name]
            _require_cuda(idx, f"batch[{s.name}]")
            if idx.dtype not in (torch.int64, torch.int32):
                idx = idx.long()  # the reference casts with .long() (base_model.py:271)
            idx = idx.contiguous()
            B = idx.shape[0]
            if self.B is None:
                self.B = B
            elif self.B != B:
                raise L.NrxError("features disagree on the batch size")
            f = self.arr[i]
            f.table = w.data_ptr()
            f.rows = w.shape[0]
            f.dim = w.shape[1]
            f.row_stride = w.stride(0)
            f.table_id = s.table_id
            f.idx_dtype = L.IDX_I32 if idx.dtype == torch.int32 else L.IDX_I64
            f.idx = idx.data_ptr()
            self.keep += [w, idx]
            if s.is_array:
                if idx.dim() != 2:
                    raise L.NrxError(f"array feature {s.name}: expected [B, L] ids")
                f.L = idx.shape[1]
                mask = batch.get(f"{s.name}_mask", None)
                if mask is not None:
                    _require_cuda(mask, f"batch[{s.name}_mask]")
                    mask = mask.to(torch.float32).contiguous()
                    f.pool = L.POOL_MASKED_MEAN
                    f.mask = mask.data_ptr()
                    self.keep.append(mask)
                    if want_inv_den:
                        inv = torch.empty(B, dtype=torch.float32, device=dev)
                        f.inv_den = inv.data_ptr()
                        self.inv_den[s.name] = inv
                else:
                    f.pool = L.POOL_MEAN  # base_model.py:275-276
            else:
                if idx.dim() != 1:
                    raise L.NrxError(f"sparse feature {s.name}: expected [B] ids")
                f.L = 1
                f.pool = L.POOL_NONE
            f.out_col = s.out_col
        self.device = dev
        self.n = n


def embed_pool_fwd(fb: FeatBinding, out_dim: int, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1 — replaces base_model.py:262-308."""
    out = torch.empty((fb.B, out_dim), dtype=torch.float32, device=fb.device)
    lib = L.load()
    L.check(lib.nrx_embed_pool_fwd(fb.arr, fb.n, fb.B, out.data_ptr(), out.stride(0) if fb.B else out_dim,
                                   L.ptr(status), L.stream_ptr(fb.device)), "nrx_embed_pool_fwd")
    return out


def embed_pool_fwd_img(fb: FeatBinding, out_dim: int, image: torch.Tensor, want_rows: bool = True,
                       status: Optional[torch.Tensor] = None, fm_logit: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """K1 that also writes the tower's bf16 input image (`image`: the a_0 slot of a prepacked tower workspace, see
    tower_input_image); returns the fp32 rows, or None with want_rows=False (nothing else reads them).  `fm_logit` [B]:
    the FM logit over all features, computed in the same epilogue (see fm_epilogue_eligible)."""
    out = torch.empty((fb.B, out_dim), dtype=torch.float32, device=fb.device) if want_rows else None
    lib = L.load()
    L.check(lib.nrx_embed_pool_fwd_img(fb.arr, fb.n, fb.B, L.ptr(out), out.stride(0) if (out is not None and fb.B) else out_dim,
                                       image.data_ptr(), out_dim, L.ptr(fm_logit), L.ptr(status), L.stream_ptr(fb.device)),
            "nrx_embed_pool_fwd_img")
    return out


def fm_epilogue_eligible(fb: FeatBinding, fm_cols, fm_dims) -> bool:
    """The FM field set is every feature of the binding, all single-id, equal width D with D/4 a power of two, <= 32 lanes."""
    if any(s.is_array for s in fb.specs) or len(fm_cols) != len(fb.specs):
        return False
    d = fb.specs[0].dim
    if any(s.dim != d for s in fb.specs) or d % 4 or ((d // 4) & (d // 4 - 1)) or len(fb.specs) * (d // 4) > 32:
        return False
    return sorted(fm_cols) == sorted(s.out_col for s in fb.specs) and all(x == d for x in fm_dims)


def embed_img_eligible(fb: FeatBinding, out_dim: int) -> bool:
    """128-bit K1 path with a concat width the tower can take as is (multiple of 16)."""
    if out_dim % 16:
        return False
    for s, f in zip(fb.specs, fb.arr):
        if f.dim % 4 or f.row_stride % 4 or f.out_col % 4 or f.table % 16:
            return False
    return True


class BwdPlan:
    """Sorted-occurrence plan for one batch (nrx_embed_bwd_plan); reusable for any number of applies."""

    def __init__(self, fb: FeatBinding, stage: int = L.PLAN_ALL):
        """stage=L.PLAN_SORT enqueues the first half only; call merge() (on any stream ordered after this one) before the
        first apply."""
        lib = L.load()
        self.fb = fb
        self.bytes = int(lib.nrx_embed_bwd_workspace_bytes(fb.arr, fb.n, fb.B))
        if self.bytes == 0 and fb.B > 0:
            L.check(-2, "nrx_embed_bwd_workspace_bytes")
        self.ws = torch.empty(max(self.bytes, 16), dtype=torch.uint8, device=fb.device)
        L.check(lib.nrx_embed_bwd_plan_stage(fb.arr, fb.n, fb.B, self.ws.data_ptr(), self.bytes, stage, L.stream_ptr(fb.device)),
                "nrx_embed_bwd_plan" if stage == L.PLAN_ALL else "nrx_embed_bwd_plan(sort)")

    def merge(self):
        fb = self.fb
        L.check(L.load().nrx_embed_bwd_plan_stage(fb.arr, fb.n, fb.B, self.ws.data_ptr(), self.bytes, L.PLAN_MERGE,
                                                  L.stream_ptr(fb.device)), "nrx_embed_bwd_plan(merge)")


def embed_bwd_dense(plan: BwdPlan, grad_out: torch.Tensor, table_by_id: Sequence[Optional[torch.Tensor]]):
    """K3 dense mode: returns one dense gradient per table id (None where unused)."""
    fb = plan.fb
    grad_out = grad_out.contiguous()
    grads = [None if w is None else torch.empty_like(w) for w in table_by_id]
    lib = L.load()
    garr = L.ptr_array(grads, L.NRX_MAX_TABLES)
    L.check(lib.nrx_embed_bwd_apply(fb.arr, fb.n, fb.B, grad_out.data_ptr(), grad_out.stride(0), L.BWD_DENSE,
                                    garr, None, None, plan.ws.data_ptr(), plan.bytes, L.stream_ptr(fb.device)),
            "nrx_embed_bwd_apply(dense)")
    return grads


def embed_bwd_rowopt(plan: BwdPlan, grad_out: torch.Tensor, table_by_id, mode: int, lr: float, step: int = 1,
                     betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, m_by_id=None, v_by_id=None):
    """K3 fused sparse row update (SGD / AdamW on touched rows only), in place."""
    fb = plan.fb
    grad_out = grad_out.contiguous()
    opt = L.NrxRowOpt()
    opt.lr, opt.beta1, opt.beta2, opt.eps, opt.weight_decay, opt.step = lr, betas[0], betas[1], eps, weight_decay, step
    for t in range(len(table_by_id)):
        if m_by_id is not None and m_by_id[t] is not None:
            opt.m[t] = m_by_id[t].data_ptr()
            opt.v[t] = v_by_id[t].data_ptr()
    lib = L.load()
    L.check(lib.nrx_embed_bwd_apply(fb.arr, fb.n, fb.B, grad_out.data_ptr(), grad_out.stride(0), mode, None,
                                    L.ptr_array(table_by_id, L.NRX_MAX_TABLES), C.byref(opt), plan.ws.data_ptr(),
                                    plan.bytes, L.stream_ptr(fb.device)), "nrx_embed_bwd_apply(rowopt)")


class EmbedPoolFn(torch.autograd.Function):
    """features = get_embeddings_from_batch(batch); dense table gradients like nn.Embedding(padding_idx=0)."""

    @staticmethod
    def forward(ctx, fb: FeatBinding, out_dim: int, table_names: List[str], *weights):
        ctx.fb = fb
        ctx.table_names = table_names
        ctx.weights = weights
        return embed_pool_fwd(fb, out_dim, status=getattr(fb, "status", None))

    @staticmethod
    def backward(ctx, grad_out):
        fb = ctx.fb
        plan = BwdPlan(fb)
        by_id: List[Optional[torch.Tensor]] = [None] * L.NRX_MAX_TABLES
        name_to_id = {s.table: s.table_id for s in fb.specs}
        for nme, w in zip(ctx.table_names, ctx.weights):
            if nme in name_to_id:
                by_id[name_to_id[nme]] = w
        grads = embed_bwd_dense(plan, grad_out, by_id)
        out = []
        for nme, w in zip(ctx.table_names, ctx.weights):
            out.append(grads[name_to_id[nme]] if nme in name_to_id else None)
        return (None, None, None, *out)


# --------------------------------------------------------------------------- #
# Field logits (FM / wide / LR)                                                #
# --------------------------------------------------------------------------- #

def field_logit_fwd(x: torch.Tensor, cols, dims, mode: int, logit: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, "x")
    B = x.shape[0]
    acc = 1 if logit is not None else 0
    if logit is None:
        logit = torch.empty(B, dtype=torch.float32, device=x.device)
    lib = L.load()
    L.check(lib.nrx_field_logit_fwd(x.data_ptr(), x.stride(0), B, L.i32_array(cols), L.i32_array(dims), len(cols), mode,
                                    logit.data_ptr(), acc, L.stream_ptr(x.device)), "nrx_field_logit_fwd")
    return logit


def field_logit_bwd(x, cols, dims, mode, dlogit, grad_x, accumulate: bool):
    lib = L.load()
    L.check(lib.nrx_field_logit_bwd(x.data_ptr(), x.stride(0), x.shape[0], L.i32_array(cols), L.i32_array(dims), len(cols),
                                    mode, dlogit.data_ptr(), grad_x.data_ptr(), grad_x.stride(0), 1 if accumulate else 0,
                                    L.stream_ptr(x.device)), "nrx_field_logit_bwd")


class FieldLogitFn(torch.autograd.Function):
    """[B, ΣD] features -> [B] logit term (FM / wide / sum); backward w.r.t. the features."""

    @staticmethod
    def forward(ctx, x, cols, dims, mode):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.meta = (list(cols), list(dims), mode)
        return field_logit_fwd(x, cols, dims, mode)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        cols, dims, mode = ctx.meta
        gx = torch.zeros_like(x)
        field_logit_bwd(x, cols, dims, mode, g.contiguous(), gx, accumulate=False)
        return gx, None, None, None


# --------------------------------------------------------------------------- #
# sigmoid + BCE                                                                #
# --------------------------------------------------------------------------- #

def logit_loss_fwd(terms: Sequence[torch.Tensor], bias: Optional[torch.Tensor], label: Optional[torch.Tensor],
                   want_dlogit: bool = True):
    """prob (+ per-sample loss, dL/dlogit of the MEAN loss when `label` is given)."""
    t0 = terms[0]
    B = t0.shape[0]
    dev = t0.device
    prob = torch.empty(B, dtype=torch.float32, device=dev)
    loss = dl = None
    lptr, lstride = 0, 0
    if label is not None:
        _require_cuda(label, "label")
        if label.dtype != torch.float32:
            label = label.float()
        lptr, lstride = label.data_ptr(), label.stride(0)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        dl = torch.empty(B, dtype=torch.float32, device=dev) if want_dlogit else None
    ts = [t.contiguous().view(-1) for t in terms]
    lib = L.load()
    L.check(lib.nrx_logit_loss_fwd(L.ptr_array(ts), len(ts), L.ptr(bias), B, lptr, lstride, prob.data_ptr(),
                                   L.ptr(loss), L.ptr(dl), L.stream_ptr(dev)), "nrx_logit_loss_fwd")
    return prob, loss, dl


def reduce_sum(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Deterministic (fixed-tree) sum * scale -> 0-dim tensor."""
    x = x.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    lib = L.load()
    L.check(lib.nrx_reduce_f32(x.data_ptr(), x.numel(), scale, out.data_ptr(), L.stream_ptr(x.device)), "nrx_reduce_f32")
    return out[0]


def reduce_mean(x: torch.Tensor) -> torch.Tensor:
    return reduce_sum(x, 1.0 / max(x.numel(), 1))


class SigmoidFn(torch.autograd.Function):
    """prob = sigmoid(sum(terms) + bias) (nrx_logit_loss_fwd); backward = nrx_sigmoid_bwd."""

    @staticmethod
    def forward(ctx, bias, *terms):
        prob, _, _ = logit_loss_fwd(terms, bias, None)
        ctx.save_for_backward(prob)
        ctx.has_bias = bias is not None
        ctx.shapes = [t.shape for t in terms]
        return prob

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        g = g.contiguous().view(-1)
        gz = torch.empty_like(p)
        lib = L.load()
        L.check(lib.nrx_sigmoid_bwd(p.data_ptr(), g.data_ptr(), p.numel(), gz.data_ptr(), L.stream_ptr(p.device)),
                "nrx_sigmoid_bwd")
        gb = reduce_sum(gz).view(1) if ctx.has_bias else None
        return (gb, *[gz.view(s) for s in ctx.shapes])


class BceFn(torch.autograd.Function):
    """F.binary_cross_entropy(prob, label, 'mean') (bceLoss, deep/model.py:32-33)."""

    @staticmethod
    def forward(ctx, prob, label):
        p = prob.contiguous().view(-1)
        y = label.view(-1)
        if y.dtype != torch.float32:
            y = y.float()
        _require_cuda(p, "prob")
        _require_cuda(y, "label")
        loss = torch.empty_like(p)
        lib = L.load()
        L.check(lib.nrx_bce_fwd(p.data_ptr(), y.data_ptr(), y.stride(0), p.numel(), loss.data_ptr(), L.stream_ptr(p.device)),
                "nrx_bce_fwd")
        ctx.save_for_backward(p, y)
        ctx.shape = prob.shape
        return reduce_mean(loss)

    @staticmethod
    def backward(ctx, g):
        p, y = ctx.saved_tensors
        gp = torch.empty_like(p)
        g = g.contiguous().view(1)
        lib = L.load()
        L.check(lib.nrx_bce_bwd(p.data_ptr(), y.data_ptr(), y.stride(0), p.numel(), g.data_ptr(), gp.data_ptr(),
                                L.stream_ptr(p.device)), "nrx_bce_bwd")
        return gp.view(ctx.shape), None


# --------------------------------------------------------------------------- #
# Optimizer                                                                    #
# --------------------------------------------------------------------------- #

def adamw_dense_(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW update on one flat fp32 tensor, in place (deep/model.py:55)."""
    lib = L.load()
    L.check(lib.nrx_adamw_dense(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, betas[0], betas[1],
                                eps, weight_decay, step, L.stream_ptr(p.device)), "nrx_adamw_dense")


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    lib = L.load()
    L.check(lib.nrx_l2_normalize(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], y.data_ptr(), y.stride(0),
                                 L.stream_ptr(x.device)), "nrx_l2_normalize")
    return y


# --------------------------------------------------------------------------- #
# DSSM training tail: normalise + in-batch negatives + InfoNCE, fused           #
# --------------------------------------------------------------------------- #

def dssm_infonce(user: torch.Tensor, item: torch.Tensor, perms: Sequence[torch.Tensor], mask: Optional[torch.Tensor] = None,
                 temperature: float = 0.1, want_grads: bool = True):
    """-> (loss_per_sample [B], grad_user [B,d] | None, grad_item [B,d] | None): gradients of mean(loss_per_sample) with
    respect to the RAW tower outputs `user`, `item` (recall/DSSM/model.py:51-73,92-110 in two launches)."""
    _require_cuda(user, "user"); _require_cuda(item, "item")
    user, item = user.detach().float().contiguous(), item.detach().float().contiguous()
    B, d = user.shape
    if item.shape != user.shape:
        raise L.NrxError(f"dssm_infonce: user {tuple(user.shape)} and item {tuple(item.shape)} must have the same shape")
    dev = user.device
    perms = [p.to(device=dev, dtype=torch.int64).contiguous() for p in perms]
    for p in perms:
        if p.numel() != B:
            raise L.NrxError("dssm_infonce: every permutation must have B elements")
    lib = L.load()
    nbytes = int(lib.nrx_dssm_infonce_workspace_bytes(B, len(perms)))
    if nbytes == 0:
        L.check(-2, "nrx_dssm_infonce_workspace_bytes")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    gu = torch.empty_like(user) if want_grads else None
    gi = torch.empty_like(item) if want_grads else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    mptr, mstride = 0, 0
    if mask is not None:
        _require_cuda(mask, "mask")
        mask = mask.detach().float()
        mptr, mstride = mask.data_ptr(), mask.stride(0) if mask.dim() > 0 else 0
    L.check(lib.nrx_dssm_infonce(user.data_ptr(), user.stride(0), item.data_ptr(), item.stride(0), B, d,
                                 L.ptr_array(perms, max(len(perms), 1)), len(perms), mptr, mstride, float(temperature),
                                 loss.data_ptr(), L.ptr(gu), d, L.ptr(gi), d, status.data_ptr(), ws.data_ptr(), nbytes,
                                 L.stream_ptr(dev)), "nrx_dssm_infonce")
    return loss, gu, gi, status


class InfoNCEFn(torch.autograd.Function):
    """mean InfoNCE loss of (raw user rows, raw item rows) with in-batch negatives item[perm_j]; the backward was computed by
    the forward's two launches and is only scaled here."""

    @staticmethod
    def forward(ctx, user, item, mask, temperature, *perms):
        need = user.requires_grad or item.requires_grad
        loss, gu, gi, status = dssm_infonce(user, item, perms, mask, temperature, want_grads=need)
        ctx.save_for_backward(*(t for t in (gu, gi) if t is not None))
        ctx.need = need
        ctx.status = status
        ctx.n_perms = len(perms)
        return loss.mean()

    @staticmethod
    def backward(ctx, g):
        if not ctx.need:
            return (None, None, None, None) + (None,) * ctx.n_perms
        gu, gi = ctx.saved_tensors
        return (gu * g, gi * g, None, None) + (None,) * ctx.n_perms


# --------------------------------------------------------------------------- #
# Gather-fused FM (sparse-only, equal widths): BASELINE config 2                #
# --------------------------------------------------------------------------- #

def fm_fused_fwd(fb: FeatBinding, bias: Optional[torch.Tensor], label: Optional[torch.Tensor] = None,
                 want_logit: bool = False, status: Optional[torch.Tensor] = None):
    """-> (prob[B], loss_per_sample[B] | None, dlogit[B] | None, logit[B] | None)."""
    dev = fb.device
    B = fb.B
    prob = torch.empty(B, dtype=torch.float32, device=dev)
    logit = torch.empty(B, dtype=torch.float32, device=dev) if want_logit else None
    loss = dl = None
    lptr, lstride = 0, 0
    if label is not None:
        _require_cuda(label, "label")
        if label.dtype != torch.float32:
            label = label.float()
        lptr, lstride = label.data_ptr(), label.stride(0)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        dl = torch.empty(B, dtype=torch.float32, device=dev)
    lib = L.load()
    L.check(lib.nrx_fm_fused_fwd(fb.arr, fb.n, B, L.ptr(bias), lptr, lstride, L.ptr(logit), prob.data_ptr(), L.ptr(loss),
                                 L.ptr(dl), L.ptr(status), L.stream_ptr(dev)), "nrx_fm_fused_fwd")
    return prob, loss, dl, logit


def fm_fused_bwd(fb: FeatBinding, dlogit: torch.Tensor, out_dim: int) -> torch.Tensor:
    gx = torch.empty((fb.B, out_dim), dtype=torch.float32, device=fb.device)
    lib = L.load()
    L.check(lib.nrx_fm_fused_bwd(fb.arr, fb.n, fb.B, dlogit.data_ptr(), gx.data_ptr(), gx.stride(0) if fb.B else out_dim,
                                 L.stream_ptr(fb.device)), "nrx_fm_fused_bwd")
    return gx


def fm_fused_eligible(specs: Sequence[FeatSpec], tables: Dict[str, torch.Tensor]) -> bool:
    if not specs or any(s.is_array for s in specs):
        return False
    d = specs[0].dim
    if any(s.dim != d for s in specs) or d % 4 or d > 128 or ((d // 4) & (d // 4 - 1)):
        return False
    return all(tables[s.table].data_ptr() % 16 == 0 and tables[s.table].stride(0) % 4 == 0 for s in specs)


class FmFusedFn(torch.autograd.Function):
    """prob[B,1] = sigmoid(bias + FM(gathered rows)); the concat is never materialised in forward."""

    @staticmethod
    def forward(ctx, fb: FeatBinding, out_dim: int, table_names: List[str], bias, *weights):
        prob, _, _, _ = fm_fused_fwd(fb, bias, status=getattr(fb, "status", None))
        ctx.fb, ctx.out_dim, ctx.table_names, ctx.weights = fb, out_dim, table_names, weights
        ctx.save_for_backward(prob)
        return prob.view(-1, 1)

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        fb = ctx.fb
        g = g.contiguous().view(-1)
        dl = torch.empty_like(p)
        lib = L.load()
        L.check(lib.nrx_sigmoid_bwd(p.data_ptr(), g.data_ptr(), p.numel(), dl.data_ptr(), L.stream_ptr(p.device)),
                "nrx_sigmoid_bwd")
        gx = fm_fused_bwd(fb, dl, ctx.out_dim)
        plan = BwdPlan(fb)
        by_id: List[Optional[torch.Tensor]] = [None] * L.NRX_MAX_TABLES
        name_to_id = {s.table: s.table_id for s in fb.specs}
        for nme, w in zip(ctx.table_names, ctx.weights):
            if nme in name_to_id:
                by_id[name_to_id[nme]] = w
        grads = embed_bwd_dense(plan, gx, by_id)
        gw = [grads[name_to_id[n]] if n in name_to_id else None for n in ctx.table_names]
        return (None, None, None, reduce_sum(dl).view(1), *gw)


# --------------------------------------------------------------------------- #
# K4: fused bf16 tower (tcgen05)                                               #
# --------------------------------------------------------------------------- #

def _tower_struct(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], negative_slope: Optional[float]):
    n = len(weights)
    if n < 1 or n > L.NRX_MAX_LAYERS:
        raise L.NrxError(f"{n} layers outside [1,{L.NRX_MAX_LAYERS}]")
    t = L.NrxTower()
    t.n_layers = n
    t.dims[0] = weights[0].shape[1]
    keep = []
    for i, (w, b) in enumerate(zip(weights, biases)):
        _require_cuda(w, f"tower weight {i}")
        wc, bc = w.detach().contiguous(), b.detach().contiguous()
        if wc.dtype != torch.float32 or bc.dtype != torch.float32:
            raise L.NrxError("tower parameters must be fp32 (they are packed to bf16 on the device)")
        keep += [wc, bc]
        t.dims[i + 1] = wc.shape[0]
        t.w[i] = wc.data_ptr()
        t.b[i] = bc.data_ptr()
    t.act = L.ACT_RELU if negative_slope is None else L.ACT_LEAKY
    t.negative_slope = 0.0 if negative_slope is None else float(negative_slope)
    return t, keep


def tower_prepack(B: int, weights, biases, negative_slope=None, training=False):
    """Allocate the tower workspace and pack the weights into it on the CURRENT stream (nrx_tower_pack).
    Returns a handle for tower_fwd(..., packed=handle): lets a trainer pack on a forked stream."""
    t, keep = _tower_struct(weights, biases, negative_slope)
    dev = weights[0].device
    lib = L.load()
    nbytes = int(lib.nrx_tower_workspace_bytes(C.byref(t), B, 1 if training else 0))
    if nbytes == 0:
        L.check(-2, "nrx_tower_workspace_bytes")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(lib.nrx_tower_pack(C.byref(t), B, 1 if training else 0, ws.data_ptr(), nbytes, L.stream_ptr(dev)), "nrx_tower_pack")
    return t, keep, ws, nbytes


def tower_head(B: int, device, terms=(), bias=None, label=None, want_loss=True, want_logit=False):
    """NrxTowerHead + its output tensors: (struct, keepalive, prob, loss_per_sample, dlogit, logit)."""
    h = L.NrxTowerHead()
    ts = [t.contiguous().view(-1) for t in terms]
    if len(ts) > 4:
        raise L.NrxError("the fused tower head takes at most 4 extra logit terms")
    h.n_terms = len(ts)
    for i, t in enumerate(ts):
        _require_cuda(t, "head term")
        h.terms[i] = t.data_ptr()
    prob = torch.empty(B, dtype=torch.float32, device=device)
    h.prob = prob.data_ptr()
    keep = list(ts) + [prob]
    loss = dl = logit = None
    if bias is not None:
        h.bias = bias.data_ptr()
        keep.append(bias)
    if label is not None:
        _require_cuda(label, "label")
        if label.dtype != torch.float32:
            label = label.float()
        h.label, h.label_stride = label.data_ptr(), label.stride(0)
        keep.append(label)
        if want_loss:
            loss = torch.empty(B, dtype=torch.float32, device=device)
            dl = torch.empty(B, dtype=torch.float32, device=device)
            h.loss_per_sample, h.dlogit = loss.data_ptr(), dl.data_ptr()
    if want_logit:
        logit = torch.empty(B, dtype=torch.float32, device=device)
        h.logit = logit.data_ptr()
    return h, keep, prob, loss, dl, logit


def tower_fwd(x: Optional[torch.Tensor], weights, biases, negative_slope=None, training=False, packed=None, head=None,
              ximg_rows: Optional[int] = None):
    """y = MLP(x) on tensor cores; returns (y, ctx) where ctx carries the workspace for tower_bwd.

    head       — an ops.tower_head(...) tuple: the last epilogue also emits sigmoid / BCE / dL/dlogit (y is None then);
    ximg_rows  — B when the bf16 input image is already in the (prepacked) workspace (x may be None)."""
    if x is not None:
        _require_cuda(x, "tower input")
        if x.dtype != torch.float32 or x.stride(-1) != 1:
            x = x.float().contiguous()
    B = x.shape[0] if ximg_rows is None else int(ximg_rows)
    lib = L.load()
    flags = L.TOWER_TRAINING if training else 0
    tag = ""
    if packed is not None:
        t, keep, ws, nbytes = packed
        flags |= L.TOWER_PREPACKED
        tag = "(prepacked)"
    else:
        t, keep = _tower_struct(weights, biases, negative_slope)
        nbytes = int(lib.nrx_tower_workspace_bytes(C.byref(t), B, flags & 1))
        if nbytes == 0:
            L.check(-2, "nrx_tower_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=weights[0].device)
    if ximg_rows is not None:
        if packed is None:
            raise L.NrxError("an input image lives in a prepacked workspace: pass packed=")
        flags |= L.TOWER_XIMG
        tag = "(prepacked,ximg)"
    dev = ws.device
    xp, xld = (x.data_ptr(), x.stride(0)) if x is not None else (0, 0)
    if head is not None:
        L.check(lib.nrx_tower_fwd_head(C.byref(t), xp, xld, B, C.byref(head[0]), flags, ws.data_ptr(), nbytes, L.stream_ptr(dev)),
                "nrx_tower_fwd_head" + tag)
        return None, (t, keep, ws, nbytes, x)
    y = torch.empty((B, t.dims[t.n_layers]), dtype=torch.float32, device=dev)
    L.check(lib.nrx_tower_fwd(C.byref(t), xp, xld, B, y.data_ptr(), y.stride(0), flags, ws.data_ptr(), nbytes, L.stream_ptr(dev)),
            "nrx_tower_fwd" + tag)
    return y, (t, keep, ws, nbytes, x)


def tower_input_image(packed, B: int) -> torch.Tensor:
    """The uint8 view of the a_0 image slot inside a prepacked tower workspace (where K1 / K5 write the tower input)."""
    t, keep, ws, nbytes = packed
    n = t.n_layers
    ao, aw = (C.c_int64 * n)(), (C.c_int32 * n)()
    L.check(L.load().nrx_tower_image_layout(C.byref(t), B, ao, aw, None, None), "nrx_tower_image_layout")
    nt = (B + 127) // 128
    return ws[ao[0]: ao[0] + nt * aw[0] * 256]


def tower_image_from_rows(x: torch.Tensor, image: torch.Tensor):
    """fp32 rows [B, width] -> the bf16 tile image (what K1 / K5 write directly on the fused route)."""
    _require_cuda(x, "rows")
    x = x.contiguous()
    L.check(L.load().nrx_tower_image_from_rows(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], image.data_ptr(),
                                               L.stream_ptr(x.device)), "nrx_tower_image_from_rows")


def tower_bwd(ctx, grad_y: torch.Tensor, need_gx: bool = True):
    t, keep, ws, nbytes, x = ctx
    grad_y = grad_y.contiguous()
    B = grad_y.shape[0]
    n = t.n_layers
    dev = ws.device
    gws = [torch.empty((t.dims[i + 1], t.dims[i]), dtype=torch.float32, device=dev) for i in range(n)]
    gbs = [torch.empty((t.dims[i + 1],), dtype=torch.float32, device=dev) for i in range(n)]
    gx = torch.empty((B, t.dims[0]), dtype=torch.float32, device=dev) if need_gx else None
    lib = L.load()
    L.check(lib.nrx_tower_bwd(C.byref(t), 0, 0, B, grad_y.data_ptr(), grad_y.stride(0), L.ptr(gx),
                              gx.stride(0) if gx is not None else 0, 0, L.ptr_array(gws, L.NRX_MAX_LAYERS),
                              L.ptr_array(gbs, L.NRX_MAX_LAYERS), ws.data_ptr(), nbytes, L.stream_ptr(dev)),
            "nrx_tower_bwd")
    return gx, gws, gbs


class TowerFn(torch.autograd.Function):
    """MLP / DSSM tower: forward(x, slope, n, *weights, *biases)."""

    @staticmethod
    def forward(ctx, x, negative_slope, n, *params):
        ws, bs = params[:n], params[n:]
        training = any(ctx.needs_input_grad)  # grad mode is off inside Function.forward
        y, tctx = tower_fwd(x, ws, bs, negative_slope, training=training)
        ctx.tctx = tctx if training else None
        ctx.n = n
        ctx.need_gx = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, gy):
        if ctx.tctx is None:
            raise L.NrxError("tower backward without a training forward")
        gx, gws, gbs = tower_bwd(ctx.tctx, gy, need_gx=ctx.need_gx)
        return (gx, None, None, *gws, *gbs)


def tower_images(ctx, B: int):
    """Decode the saved bf16 tile images of a training forward/backward -> (a[l], dz[l]) as fp32 [B, width]."""
    t, keep, ws, nbytes, x = ctx
    n = t.n_layers
    ao, aw = (C.c_int64 * n)(), (C.c_int32 * n)()
    do, dw = (C.c_int64 * n)(), (C.c_int32 * n)()
    lib = L.load()
    L.check(lib.nrx_tower_image_layout(C.byref(t), B, ao, aw, do, dw), "nrx_tower_image_layout")
    nt = (B + 127) // 128

    def img(off, width):
        raw = ws[off: off + nt * width * 256].view(torch.bfloat16).view(nt, width // 8, 128, 8)
        return raw.permute(0, 2, 1, 3).reshape(nt * 128, width)[:B].float()

    return [img(ao[l], aw[l]) for l in range(n)], [img(do[l], dw[l]) for l in range(n)]


# --------------------------------------------------------------------------- #
# K5: DCN-v1 cross stack                                                       #
# --------------------------------------------------------------------------- #

def dcn_cross_fwd(x: torch.Tensor, ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor]) -> torch.Tensor:
    """cat[x, cross_L(x)] ([B, 2d]) — dcn_arch.py:14-30,53-70 + dcn/model.py:29."""
    _require_cuda(x, "cross input")
    x = x.contiguous()
    B, d = x.shape
    out = torch.empty((B, 2 * d), dtype=torch.float32, device=x.device)
    wl = [w.detach().contiguous().view(-1) for w in ws]
    bl = [b.detach().contiguous().view(-1) for b in bs]
    lib = L.load()
    L.check(lib.nrx_dcn_cross_fwd(x.data_ptr(), x.stride(0), B, d, len(wl), L.ptr_array(wl), L.ptr_array(bl), out.data_ptr(),
                                  out.stride(0), None, L.stream_ptr(x.device)), "nrx_dcn_cross_fwd")
    return out


def dcn_cross_bwd(x, ws, bs, grad_out, need_gx=True):
    x = x.contiguous()
    grad_out = grad_out.contiguous()
    B, d = x.shape
    n = len(ws)
    wl = [w.detach().contiguous().view(-1) for w in ws]
    bl = [b.detach().contiguous().view(-1) for b in bs]
    gws = [torch.empty(d, dtype=torch.float32, device=x.device) for _ in range(n)]
    gbs = [torch.empty(d, dtype=torch.float32, device=x.device) for _ in range(n)]
    gx = torch.empty_like(x) if need_gx else None
    lib = L.load()
    nbytes = int(lib.nrx_dcn_cross_workspace_bytes(B, d, n))
    wsb = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    L.check(lib.nrx_dcn_cross_bwd(x.data_ptr(), x.stride(0), B, d, n, L.ptr_array(wl), L.ptr_array(bl), grad_out.data_ptr(),
                                  grad_out.stride(0), None, L.ptr(gx), gx.stride(0) if gx is not None else 0,
                                  L.ptr_array(gws), L.ptr_array(gbs), wsb.data_ptr(), nbytes, L.stream_ptr(x.device)),
            "nrx_dcn_cross_bwd")
    return gx, gws, gbs


def dcn_cross_fwd_img(x: torch.Tensor, ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], image: torch.Tensor):
    """Cross stack writing cat[x, x_L] straight into the tower's bf16 input image (no fp32 concat)."""
    _require_cuda(x, "cross input")
    x = x.contiguous()
    B, d = x.shape
    wl = [w.detach().contiguous().view(-1) for w in ws]
    bl = [b.detach().contiguous().view(-1) for b in bs]
    lib = L.load()
    L.check(lib.nrx_dcn_cross_fwd_img(x.data_ptr(), x.stride(0), B, d, len(wl), L.ptr_array(wl), L.ptr_array(bl), None, 0,
                                      image.data_ptr(), L.stream_ptr(x.device)), "nrx_dcn_cross_fwd_img")


def dcn_img_eligible(x: torch.Tensor, ws, bs) -> bool:
    d = x.shape[1]
    return (d % 8 == 0 and d <= 256 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0
            and all(w.data_ptr() % 16 == 0 for w in ws) and all(b.data_ptr() % 16 == 0 for b in bs))


class CrossFn(torch.autograd.Function):
    """forward(x, n, *w[d,1], *b[d,1]) -> cat[x, x_L]."""

    @staticmethod
    def forward(ctx, x, n, *params):
        ws, bs = params[:n], params[n:]
        ctx.save_for_backward(x, *params)
        ctx.n = n
        return dcn_cross_fwd(x, ws, bs)

    @staticmethod
    def backward(ctx, g):
        x, *params = ctx.saved_tensors
        n = ctx.n
        gx, gws, gbs = dcn_cross_bwd(x, params[:n], params[n:], g, need_gx=ctx.needs_input_grad[0])
        return (gx, None, *[a.view_as(p) for a, p in zip(gws, params[:n])], *[a.view_as(p) for a, p in zip(gbs, params[n:])])
