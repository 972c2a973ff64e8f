"""FusedTrainer — one whole training step (fwd + BCE + bwd + optimizer) of a sort model as a single
captured CUDA graph over libnrx kernels.

What it replaces: the per-step work of Lightning's fit loop around the reference modules
(`training_step` e.g. sort/deep/model.py:45-52, `loss.backward()`, `AdamW.step`, `CosinDecayLR.step`
:54-65) minus logging and the per-step sklearn AUC.

Differences from the autograd route (model(batch); loss.backward(); torch.optim.AdamW):
  * embedding tables: `table_update="dense"` (the default) reproduces the reference's dense AdamW over whole tables
    (every row decays and moves every step); `table_update="sparse"` (explicit opt-in) is the fused sparse row AdamW
    inside K3, rows the batch touched only — see DESIGN.md "optimizer semantics";
  * all dense parameters live in one flat fp32 buffer and take one nrx_adamw_dense_dev launch;
  * lr / bias-correction scalars are produced on the device (nrx_hparams_step), so the graph replays
    with no host-written arguments; the sort plan runs on a forked stream (its chunk sort beside K1, its merge
    after the dX chain — DESIGN.md section 4, "Fused training step").
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops


def _align(x, a=256):
    return (x + a - 1) // a * a


class BatchLayout:
    """Byte layout of one batch as a single blob (ids | masks | label) so a step needs ONE copy
    (pinned host -> device, or device pool slot -> static slot)."""

    def __init__(self, model, B: int, id_dtype=torch.int64, n_labels: int = 2):
        """n_labels: columns of the batch's label tensor (MIND feature files carry 2; the reference accepts any count and
        its loss reads column 0 only, `deep/model.py:69`)."""
        if n_labels < 1:
            raise L.NrxError("n_labels must be >= 1")
        self.B, self.n_labels = B, int(n_labels)
        names = sorted(model.user_feature_names | model.item_feature_names)
        self.fields = []  # (key, dtype, shape, offset)
        off = 0
        isz = 8 if id_dtype == torch.int64 else 4
        for n in names:
            if n in model.array_feature_names:
                Lh = int(model.array_max_length[n])
                self.fields.append((n, id_dtype, (B, Lh), off)); off = _align(off + B * Lh * isz)
                self.fields.append((n + "_mask", torch.float32, (B, Lh), off)); off = _align(off + B * Lh * 4)
            else:
                self.fields.append((n, id_dtype, (B,), off)); off = _align(off + B * isz)
        self.fields.append(("label", torch.float32, (B, self.n_labels), off)); off = _align(off + B * self.n_labels * 4)
        self.nbytes = off

    def views(self, blob: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for key, dt, shape, off in self.fields:
            n = 1
            for s in shape:
                n *= s
            esz = torch.empty((), dtype=dt).element_size()
            out[key] = blob[off: off + n * esz].view(dt).view(*shape)
        return out

    def pack(self, batch: Dict[str, torch.Tensor], blob: torch.Tensor):
        v = self.views(blob)
        for key, dt, shape, off in self.fields:
            src = batch[key]
            if key == "label" and src.numel() != self.B * self.n_labels:
                raise L.NrxError(f"label has {src.numel() // max(self.B, 1)} columns per sample, the layout was built for "
                                 f"n_labels={self.n_labels} (pass n_labels= to BatchLayout / FusedTrainer)")
            v[key].copy_(src.to(dt).view(*shape))
        return blob


class FusedTrainer:
    KINDS = ("lr", "fm", "deep", "widedeep", "dcn", "deepfm")
    _TOWER_PREFIX = {"deep": "score_fc.network.network", "deepfm": "score_fc.deep_network.network",
                     "widedeep": "score_fc.deep_network.network", "dcn": "score_fc.score_fc.network"}

    def __init__(self, model, B: int, kind: Optional[str] = None, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                 use_graph: bool = True, id_dtype=torch.int64, table_update: str = "dense", dense_impl: str = "flat",
                 n_labels: int = 2):
        """n_labels: label columns per sample in the batches (the loss reads column 0, as the reference does).
           table_update:
             "dense"  — (default) the reference's semantics (sort/deep/model.py:55): AdamW with weight decay 0.01 moves
                        every row of every table every step, touched or not;
             "sparse" — explicit opt-in: fused sparse-row AdamW inside K3, only the rows the batch touched move
                        (lazy rows) — a different trajectory than the reference's, needed once a dense pass over
                        the tables per step is unaffordable (BASELINE config 5).
           dense_impl (dense only; bitwise-identical results, tests/test_gpu_trainer.py):
             "flat"   — K3 writes dense table gradients into the same flat buffer as the tower gradients and ONE dense
                        AdamW sweeps every parameter (also what the data-parallel trainer exchanges);
             "split"  — touched rows: fused row AdamW inside K3; untouched rows: the g = 0 AdamW step swept on a few
                        SMs beside forward/backward (nrx_adamw_untouched_rows).  Measured equal to "flat" at
                        MIND-small (122-124 us vs 124 us per step): the sweep's full-SM CTAs collide with the
                        persistent tower CTAs; kept as an option."""
        if table_update not in ("sparse", "dense"):
            raise L.NrxError(f"table_update must be 'sparse' or 'dense', got {table_update!r}")
        if dense_impl not in ("split", "flat"):
            raise L.NrxError(f"dense_impl must be 'split' or 'flat', got {dense_impl!r}")
        self.table_update = table_update
        # dense semantics, two implementations with identical results:
        #   "split" — touched rows: fused row AdamW inside K3; untouched rows: the g = 0 AdamW step, swept on the forked
        #             stream right after the sort plan, concurrently with forward/backward (nrx_adamw_untouched_rows);
        #   "flat"  — K3 writes dense table gradients into the flat gradient buffer, one AdamW sweeps everything
        #             (what the data-parallel trainer all-reduces).
        self._flat_tables = table_update == "dense" and dense_impl == "flat"
        self._split_dense = table_update == "dense" and dense_impl == "split"
        self.model = model
        self.kind = kind or type(model).__name__.lower()
        if self.kind not in self.KINDS:
            raise L.NrxError(f"unsupported model kind {self.kind}")
        self.B = B
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise L.NrxError("FusedTrainer needs the model on a CUDA device (no CPU fallback)")
        hp = model.train_hparams
        self.lr, self.min_lr = float(hp.lr), float(hp.min_lr)
        self.milestones = [int(x) for x in hp.lr_milestones]
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        self.lib = L.load()
        self.layout = BatchLayout(model, B, id_dtype, n_labels)
        self.blob = torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=self.dev)
        self.batch = self.layout.views(self.blob)
        # loss and the id-status word share one 8-byte buffer so that feed() reads both back with one copy:
        # K1 sets the status when an id is outside its table (nn.Embedding would raise; base_model.py:271)
        self._loss_status = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self.id_status = self._loss_status[1:2].view(torch.int32)
        self.fused_input = os.environ.get("NRX_FUSED_INPUT", "1") == "1"
        # asynchronous status read-back (step()): stream, pinned word and events exist before the first step
        self._stat = dict(n=0, posted=False, stream=torch.cuda.Stream(device=self.dev),
                          host=torch.zeros(2, dtype=torch.float32).pin_memory(), ev=torch.cuda.Event(), done=torch.cuda.Event())
        self.d_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.d_hp = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self._dense_applied = False
        self._flatten_dense()
        self._table_state()
        self.fb, self.dims, self.names, self.out_dim = model.bind_features(self.batch, model.user_feature_names | model.item_feature_names)
        self.fm_fused = self.kind == "fm" and ops.fm_fused_eligible(self.fb.specs, model._weights())
        self._wd_idx = None
        if self.kind == "widedeep":  # built outside graph capture (host -> device copy)
            _, deep_cols = model._split_cols(self.dims, self.names)
            self._wd_idx = torch.as_tensor(deep_cols, device=self.dev)
        # the plan's stream shares the chain's priority: its merge must be placed BEFORE the dW GEMMs (lower priority, below)
        # when both become ready at the end of the dX chain — the apply and the optimizer wait for the merge, nobody for dW
        hi = -1 if os.environ.get("NRX_MAIN_PRIO", "1") == "1" else 0
        self.side = torch.cuda.Stream(device=self.dev, priority=hi)
        self.s_dw = torch.cuda.Stream(device=self.dev, priority=0)
        self.side2 = torch.cuda.Stream(device=self.dev)
        self.side3 = torch.cuda.Stream(device=self.dev, priority=hi)   # FM-logit backward + table-gradient apply: also ahead of dW
        self.side4 = torch.cuda.Stream(device=self.dev)   # untouched-row sweep: from the end of the plan to the end of the step
        self.loss = self._loss_status[0:1]
        self.prob = None
        self.graph = None
        self.launches_per_step = None
        self.use_graph = use_graph
        if use_graph:
            self._capture()

    # ---- parameter plumbing ------------------------------------------------------------------
    def _alloc_flat(self, n: int) -> torch.Tensor:
        return torch.zeros(n, dtype=torch.float32, device=self.dev)

    def _flatten_dense(self):
        dense = [(n, p) for n, p in self.model.named_parameters()
                 if self._flat_tables or not n.startswith("embedding_tables.")]
        total = sum(_align(p.numel(), 4) for _, p in dense)
        self.flat_p = self._alloc_flat(max(total, 4))   # hook: the multi-GPU trainer puts these in peer-mapped memory
        self.flat_g = self._alloc_flat(max(total, 4))
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self.dense_views, self.grad_views = {}, {}
        off = 0
        for n, p in dense:
            k = p.numel()
            self.flat_p[off: off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off: off + k].view(p.shape)  # parameters become views of the flat buffer
            self.dense_views[n] = p.data
            self.grad_views[n] = self.flat_g[off: off + k].view(p.shape)
            off += _align(k, 4)
        self.n_dense = off

    def _table_state(self):
        self.tables_by_id: List[Optional[torch.Tensor]] = [None] * L.NRX_MAX_TABLES
        self.m_by_id: List[Optional[torch.Tensor]] = [None] * L.NRX_MAX_TABLES
        self.v_by_id: List[Optional[torch.Tensor]] = [None] * L.NRX_MAX_TABLES
        self.table_grads_by_id: List[Optional[torch.Tensor]] = [None] * L.NRX_MAX_TABLES
        span = (1 << 62, 0)
        for name, tid in self.model._table_ids.items():
            w = self.model.embedding_tables[name].weight.data
            self.tables_by_id[tid] = w
            if self._flat_tables:   # moments live in the flat buffers
                self.table_grads_by_id[tid] = self.grad_views[f"embedding_tables.{name}.weight"]
                e0 = (self.table_grads_by_id[tid].data_ptr() - self.flat_g.data_ptr()) // 4
                span = (min(span[0], e0), max(span[1], e0 + w.numel()))
            else:
                self.m_by_id[tid] = torch.zeros_like(w)
                self.v_by_id[tid] = torch.zeros_like(w)
        if self._flat_tables:
            # the tables' gradients as ONE span of the flat buffer (named_parameters lists the tables together); a dense
            # parameter that happened to sit inside it would merely have its gradient zeroed before it is rewritten
            self._table_grad_span = self.flat_g[span[0]:span[1]] if span[1] > span[0] else self.flat_g[:0]

    # ---- raw op helpers writing into preallocated buffers ------------------------------------------
    def _sp(self):
        return L.stream_ptr(self.dev)

    def _reduce(self, x, scale, out):
        L.check(self.lib.nrx_reduce_f32(x.data_ptr(), x.numel(), scale, out.data_ptr(), self._sp()), "nrx_reduce_f32")

    def _tower(self, x, lin_names, ws_override=None, packed=None, head=None, ximg=False):
        ws = ws_override or [self.dense_views[n + ".weight"] for n in lin_names]
        bs = [self.dense_views[n + ".bias"] for n in lin_names]
        y, tctx = ops.tower_fwd(None if ximg else x, ws, bs, None, training=True, packed=packed, head=head,
                                ximg_rows=self.B if ximg else None)
        return y, tctx

    def _prepack(self, lin_names, main):
        """Pack the tower weights on a second forked stream while K1 / the field logits run (the weights were
        final when the previous step's optimizer finished)."""
        ws = [self.dense_views[n + ".weight"] for n in lin_names]
        bs = [self.dense_views[n + ".bias"] for n in lin_names]
        self.side2.wait_stream(main)
        with torch.cuda.stream(self.side2):
            return ops.tower_prepack(self.B, ws, bs, None, training=True)

    def _tower_bwd_dx(self, tctx, dl):
        t, keep, ws, nbytes, x = tctx
        gx = torch.empty((self.B, t.dims[0]), dtype=torch.float32, device=self.dev)
        L.check(self.lib.nrx_tower_bwd_dx(C.byref(t), self.B, dl.data_ptr(), 1, gx.data_ptr(), gx.stride(0), 0,
                                          ws.data_ptr(), nbytes, self._sp()), "nrx_tower_bwd_dx")
        return gx

    def _tower_bwd_dw(self, tctx, lin_names, gw_override=None):
        t, keep, ws, nbytes, x = tctx
        gws = gw_override or [self.grad_views[n + ".weight"] for n in lin_names]
        gbs = [self.grad_views[n + ".bias"] for n in lin_names]
        L.check(self.lib.nrx_tower_bwd_dw(C.byref(t), self.B, L.ptr_array(gws, L.NRX_MAX_LAYERS),
                                          L.ptr_array(gbs, L.NRX_MAX_LAYERS), ws.data_ptr(), nbytes, self._sp()),
                "nrx_tower_bwd_dw")

    def _loss_and_bias_grad(self, loss_ps, dl, bias):
        if bias is not None:  # mean loss and d/dbias = sum(dlogit) in one launch
            L.check(self.lib.nrx_reduce2_f32(loss_ps.data_ptr(), loss_ps.numel(), 1.0 / self.B, self.loss.data_ptr(),
                                             dl.data_ptr(), dl.numel(), 1.0, self.grad_views["score_fc.bias"].data_ptr(),
                                             self._sp()), "nrx_reduce2_f32")
        else:
            self._reduce(loss_ps, 1.0 / self.B, self.loss)

    def _lin_names(self, prefix):
        names, i = [], 0
        while f"{prefix}.{i}.weight" in self.dense_views:
            names.append(f"{prefix}.{i}")
            i += 2
        return names

    # ---- the step -------------------------------------------------------------------------------------
    def _step(self):
        self._fwd_bwd()
        self._update(self._plan_fb(), self._plan, self._gx)

    def _plan_fb(self):
        """Features whose occurrences the backward plan covers (the data-parallel trainer returns the global batch)."""
        return self.fb

    def _embed_fwd(self):
        """K1: features [B, ΣD] of the local batch (the row-sharded trainer overrides this with the exchange)."""
        return ops.embed_pool_fwd(self.fb, self.out_dim, status=self.id_status)

    # single-GPU trainers apply the sparse-row update inside _fwd_bwd (overlapped with dW); the distributed
    # trainers need the gradient exchange first and keep it in _update
    _inline_update = os.environ.get("NRX_INLINE_APPLY", "0") == "1"

    def _fwd_bwd(self):
        m, fb, lib = self.model, self.fb, self.lib
        self._rows_applied = False
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):  # off the critical path: only the optimizer consumes these
            # optimizer clock (lr schedule + bias corrections) and the sort plan (depends on the ids only)
            L.check(lib.nrx_hparams_step(self.d_step.data_ptr(), self.d_hp.data_ptr(), self.lr, self.min_lr, self.milestones[0],
                                         self.milestones[1], self.betas[0], self.betas[1], self._sp()), "nrx_hparams_step")
            if self._split_dense:   # needs the ids and the optimizer clock only; nothing in this step reads the rows it writes
                self.side4.wait_stream(self.side)
                with torch.cuda.stream(self.side4):
                    self._sweep_untouched(self._plan_fb())
            # The plan in two halves when a tower follows: the chunk sort runs beside K1 (neither needs much shared memory);
            # the merge (64 KB CTAs on every SM) waits until the dX chain has been launched and runs beside the dW GEMMs —
            # beside the forward it held SMs the persistent tower CTAs were waiting for (8 us of the step's critical path).
            pfb = self._plan_fb()
            staged = (self.kind in ("deep", "deepfm", "widedeep", "dcn") and not self.fm_fused
                      and lib.nrx_embed_bwd_plan_is_staged(pfb.arr, pfb.n, pfb.B) == 1)
            plan = ops.BwdPlan(self._plan_fb(), L.PLAN_SORT if staged else L.PLAN_ALL)
            if self._flat_tables:
                # dense table gradients: K3 writes the touched rows only, the rest must be zero.  ONE fill over the tables'
                # span of the flat gradient buffer, early in the step and off its critical path (five per-table memsets
                # inside the apply call used to sit between the backward and the optimizer: ~20 us of the step)
                self._table_grad_span.zero_()
        self._dense_applied = False
        packed = None
        if self.kind in ("deep", "deepfm", "dcn"):
            packed = self._prepack(self._lin_names(self._TOWER_PREFIX[self.kind]), main)
        label = self.batch["label"][:, 0]
        bias = self.dense_views.get("score_fc.bias")
        kind = self.kind
        if self.fm_fused:
            prob, loss_ps, dl, _ = ops.fm_fused_fwd(fb, bias, label, status=self.id_status)
            gx = ops.fm_fused_bwd(fb, dl, self.out_dim)
        else:
            has_tower = kind in ("deep", "deepfm", "widedeep", "dcn")
            # the tower streams its input as a bf16 tile image: K1 (Deep / DeepFM) or the cross stack (DCN) write that
            # image straight into the tower workspace, so the fp32 concat is only materialised where something else
            # reads it (FM logit, cross stack)
            img = None
            if packed is not None and self.fused_input:
                main.wait_stream(self.side2)   # the workspace (and its packed weights) exists from here on
                img = ops.tower_input_image(packed, self.B)
            x_img = (img is not None and kind in ("deep", "deepfm") and ops.embed_img_eligible(fb, self.out_dim)
                     and type(self)._embed_fwd is FusedTrainer._embed_fwd)
            cols, c = [], 0
            for d in self.dims:
                cols.append(c)
                c += d
            terms, tctx, lin, field = [], None, None, None
            gw_override = None
            fm_term = None
            if kind in ("fm", "deepfm"):
                if kind == "deepfm":
                    fcols, fdims = m.fm_fields(self.dims, self.names)
                else:
                    fcols, fdims = cols, list(self.dims)
                field = (fcols, fdims, L.FIELD_FM)
                if x_img and ops.fm_epilogue_eligible(fb, fcols, fdims):   # FM logit in K1's epilogue: no separate launch
                    fm_term = torch.empty(self.B, dtype=torch.float32, device=self.dev)
            if x_img:
                x = ops.embed_pool_fwd_img(fb, self.out_dim, img, want_rows=(kind != "deep"), status=self.id_status, fm_logit=fm_term)
            else:
                x = self._embed_fwd()
            if kind == "widedeep":
                wide_cols, deep_cols = m._split_cols(self.dims, self.names)
                field = (wide_cols, [1] * len(wide_cols), L.FIELD_WIDE)
            elif kind == "lr":
                field = (cols, list(self.dims), L.FIELD_SUM)
            s3 = self.side3
            if fm_term is not None:
                terms.append(fm_term)
            elif field is not None:   # with a tower: the term feeds the fused head of the tower's last epilogue
                terms.append(ops.field_logit_fwd(x, field[0], field[1], field[2]))
            if has_tower:
                lin = self._lin_names(self._TOWER_PREFIX[kind])
                tin, ws_override = x, None
                ximg = x_img
                if kind == "widedeep":  # column selection moved to the weight side (see widedeep/model.py)
                    w0 = self.dense_views[lin[0] + ".weight"]
                    idx = self._wd_idx
                    w0_full = w0.new_zeros(w0.shape[0], self.out_dim).index_copy(1, idx, w0)
                    ws_override = [w0_full] + [self.dense_views[n + ".weight"] for n in lin[1:]]
                    gw0_full = torch.empty_like(w0_full)
                    gw_override = [gw0_full] + [self.grad_views[n + ".weight"] for n in lin[1:]]
                if kind == "dcn":
                    cw = [self.dense_views[f"score_fc.cross_net.cross_net.{i}.w"] for i in range(len(m.score_fc.cross_net.cross_net))]
                    cb = [self.dense_views[f"score_fc.cross_net.cross_net.{i}.b"] for i in range(len(cw))]
                    if img is not None and ops.dcn_img_eligible(x, cw, cb):
                        ops.dcn_cross_fwd_img(x, cw, cb, img)
                        tin, ximg = None, True
                    else:
                        tin = ops.dcn_cross_fwd(x, cw, cb)
                if packed is not None and img is None:
                    main.wait_stream(self.side2)
                head = ops.tower_head(self.B, self.dev, terms, bias, label)
                _, tctx = self._tower(tin, lin, ws_override, packed=packed, head=head, ximg=ximg)
                prob, loss_ps, dl = head[2], head[3], head[4]
            else:
                prob, loss_ps, dl = ops.logit_loss_fwd(terms, bias, label)
            # ---- backward ----
            gx = None
            if tctx is not None:
                g_tin = self._tower_bwd_dx(tctx, dl)
                merged = None
                if staged:
                    # second half of the plan, as soon as the dX chain is out of the way: it shares the machine with the dW
                    # GEMMs (measured: making dW wait for the merge instead lengthens the step by 2 us — the apply then
                    # collides with dW, NRX_DW_AFTER_MERGE=1 keeps that order for experiments)
                    self.side.wait_stream(main)
                    with torch.cuda.stream(self.side):
                        plan.merge()
                        merged = torch.cuda.Event()
                        merged.record(self.side)
                        self._loss_and_bias_grad(loss_ps, dl, bias)   # scalar reductions: only the optimizer reads them
                # the field-logit backward, the table-gradient apply and the scalar reductions only need dl / grad_x: they
                # run on the third stream while the dW GEMMs (and, for DCN, the cross backward) proceed on the main one
                s3.wait_stream(main)
                inline = self._inline_update and kind != "dcn" and not self._flat_tables
                dense_early = self._flat_tables and kind != "dcn" and type(self)._dense_table_grads is FusedTrainer._dense_table_grads
                with torch.cuda.stream(s3):
                    if field is not None:
                        ops.field_logit_bwd(x, field[0], field[1], field[2], dl, g_tin, accumulate=True)
                    if inline or dense_early:
                        if merged is not None:
                            s3.wait_event(merged)
                        else:
                            s3.wait_stream(self.side)
                    if inline:  # grad_x is final here: update the embedding rows while dW is still being computed
                        self._apply_rows(self._plan_fb(), plan, g_tin)
                        self._rows_applied = True
                    elif dense_early:  # likewise for the dense table gradients (the optimizer then only waits for the join)
                        self._dense_table_grads(self._plan_fb(), plan, g_tin)
                        self._dense_applied = True
                    if not staged:
                        self._loss_and_bias_grad(loss_ps, dl, bias)
                if merged is not None and os.environ.get("NRX_DW_AFTER_MERGE", "0") == "1":
                    main.wait_event(merged)
                dw_low = os.environ.get("NRX_DW_LOW_PRIO", "0") == "1"   # measured: 0.0985 vs 0.0947 ms — the apply then collides with dW
                dws = self.s_dw if dw_low else main
                if dw_low:
                    self.s_dw.wait_stream(main)
                with torch.cuda.stream(dws):
                    self._tower_bwd_dw(tctx, lin, gw_override)
                    if kind == "widedeep":
                        self.grad_views[lin[0] + ".weight"].copy_(gw_override[0].index_select(1, idx))
                if kind == "dcn":
                    gx, gcw, gcb = ops.dcn_cross_bwd(x, cw, cb, g_tin)
                    for i in range(len(cw)):
                        self.grad_views[f"score_fc.cross_net.cross_net.{i}.w"].copy_(gcw[i].view(-1, 1))
                        self.grad_views[f"score_fc.cross_net.cross_net.{i}.b"].copy_(gcb[i].view(-1, 1))
                else:
                    gx = g_tin
                if dw_low:
                    main.wait_stream(self.s_dw)
                main.wait_stream(s3)
            else:
                gx = torch.zeros_like(x)
                ops.field_logit_bwd(x, field[0], field[1], field[2], dl, gx, accumulate=True)
                self._loss_and_bias_grad(loss_ps, dl, bias)
        if self.fm_fused:
            self._loss_and_bias_grad(loss_ps, dl, bias)
        main.wait_stream(self.side)
        self.prob = prob
        self._plan = plan
        self._gx = gx.contiguous()

    def _apply_rows(self, fb, plan, gx):
        """K3 apply: fused sparse-row AdamW on the embedding tables, on the current stream."""
        lib = self.lib
        opt = L.NrxRowOpt()
        opt.lr, opt.beta1, opt.beta2, opt.eps, opt.weight_decay, opt.step = self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 1
        for t in range(L.NRX_MAX_TABLES):
            if self.m_by_id[t] is not None:
                opt.m[t] = self.m_by_id[t].data_ptr()
                opt.v[t] = self.v_by_id[t].data_ptr()
        opt.d_hparams = self.d_hp.data_ptr()
        L.check(lib.nrx_embed_bwd_apply(fb.arr, fb.n, fb.B, gx.data_ptr(), gx.stride(0), L.BWD_ADAMW, None,
                                        L.ptr_array(self.tables_by_id, L.NRX_MAX_TABLES), C.byref(opt), plan.ws.data_ptr(),
                                        plan.bytes, self._sp()), "nrx_embed_bwd_apply")

    def _row_opt(self):
        opt = L.NrxRowOpt()
        opt.lr, opt.beta1, opt.beta2, opt.eps, opt.weight_decay, opt.step = self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 1
        for t in range(L.NRX_MAX_TABLES):
            if self.m_by_id[t] is not None:
                opt.m[t] = self.m_by_id[t].data_ptr()
                opt.v[t] = self.v_by_id[t].data_ptr()
        opt.d_hparams = self.d_hp.data_ptr()
        return opt

    def _sweep_untouched(self, fb):
        """Dense semantics, split implementation: g = 0 AdamW step on every row the plan does not touch."""
        if getattr(self, "_row_map", None) is None:
            n = self.lib.nrx_adamw_untouched_rows_scratch_bytes(fb.arr, fb.n)
            self._row_map = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        opt = self._row_opt()
        L.check(self.lib.nrx_adamw_untouched_rows(fb.arr, fb.n, fb.B, L.ptr_array(self.tables_by_id, L.NRX_MAX_TABLES),
                                                  C.byref(opt), self._row_map.data_ptr(), self._row_map.numel(), self._sp()),
                "nrx_adamw_untouched_rows")

    def _dense_table_grads(self, fb, plan, gx):
        """K3 dense mode: per-table [rows, D] gradients written into the flat grad buffer (zero-filled at the start of the
        step); a no-op when _fwd_bwd already ran it beside the dW GEMMs."""
        if self._dense_applied:
            return
        L.check(self.lib.nrx_embed_bwd_apply(fb.arr, fb.n, fb.B, gx.data_ptr(), gx.stride(0), L.BWD_DENSE | L.BWD_NO_ZERO,
                                             L.ptr_array(self.table_grads_by_id, L.NRX_MAX_TABLES), None, None,
                                             plan.ws.data_ptr(), plan.bytes, self._sp()), "nrx_embed_bwd_apply(dense)")

    def _adamw_flat(self):
        L.check(self.lib.nrx_adamw_dense_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                             self.flat_v.data_ptr(), self.n_dense, self.d_hp.data_ptr(), self.betas[0],
                                             self.betas[1], self.eps, self.wd, self._sp()), "nrx_adamw_dense_dev")

    def _update(self, fb, plan, gx):
        self._update_impl(fb, plan, gx)
        if self._split_dense:
            torch.cuda.current_stream(self.dev).wait_stream(self.side4)

    def _update_impl(self, fb, plan, gx):
        """Optimizer: fused sparse-row AdamW on the tables (K3 apply) + dense AdamW on the flat buffer."""
        lib = self.lib
        if self._flat_tables:
            self._dense_table_grads(fb, plan, gx)
            self._adamw_flat()
            return
        if self._rows_applied:  # the rows were updated inside _fwd_bwd; only the dense parameters are left
            if self.n_dense > 0:
                L.check(lib.nrx_adamw_dense_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                                self.flat_v.data_ptr(), self.n_dense, self.d_hp.data_ptr(), self.betas[0],
                                                self.betas[1], self.eps, self.wd, self._sp()), "nrx_adamw_dense_dev")
            return
        opt = L.NrxRowOpt()
        opt.lr, opt.beta1, opt.beta2, opt.eps, opt.weight_decay, opt.step = self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 1
        for t in range(L.NRX_MAX_TABLES):
            if self.m_by_id[t] is not None:
                opt.m[t] = self.m_by_id[t].data_ptr()
                opt.v[t] = self.v_by_id[t].data_ptr()
        opt.d_hparams = self.d_hp.data_ptr()
        main = torch.cuda.current_stream(self.dev)
        if self.n_dense > 0:  # dense AdamW || sparse-row apply: disjoint parameters
            self.side3.wait_stream(main)
            with torch.cuda.stream(self.side3):
                L.check(lib.nrx_adamw_dense_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                                self.flat_v.data_ptr(), self.n_dense, self.d_hp.data_ptr(), self.betas[0],
                                                self.betas[1], self.eps, self.wd, self._sp()), "nrx_adamw_dense_dev")
        L.check(lib.nrx_embed_bwd_apply(fb.arr, fb.n, fb.B, gx.data_ptr(), gx.stride(0), L.BWD_ADAMW, None,
                                        L.ptr_array(self.tables_by_id, L.NRX_MAX_TABLES), C.byref(opt), plan.ws.data_ptr(),
                                        plan.bytes, self._sp()), "nrx_embed_bwd_apply")
        if self.n_dense > 0:
            main.wait_stream(self.side3)

    def _capture(self):
        # warm up on a side stream (lazy module/attribute initialisation must not happen inside capture)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        snap = self._snapshot()
        with torch.cuda.stream(s):
            self._step()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._restore(snap)
        self.graph = torch.cuda.CUDAGraph()
        # the forward/backward chain is captured on a high-priority stream: when its CTAs and those of the forked
        # streams (sort plan, weight packing, field logits) are pending together, the chain's are placed first
        cap = torch.cuda.Stream(device=self.dev, priority=-1) if os.environ.get("NRX_MAIN_PRIO", "1") == "1" else None
        with torch.cuda.graph(self.graph, stream=cap):
            self._step()
        torch.cuda.synchronize(self.dev)
        self._restore(snap)  # capture does not execute, but keep state identical to "no step taken" regardless

    def _snapshot(self):
        return dict(p=self.flat_p.clone(), m=self.flat_m.clone(), v=self.flat_v.clone(), step=self.d_step.clone(),
                    t=[None if w is None else w.clone() for w in self.tables_by_id],
                    tm=[None if w is None else w.clone() for w in self.m_by_id],
                    tv=[None if w is None else w.clone() for w in self.v_by_id])

    def _restore(self, s):
        self.flat_p.copy_(s["p"]); self.flat_m.copy_(s["m"]); self.flat_v.copy_(s["v"]); self.d_step.copy_(s["step"])
        for dst, src in ((self.tables_by_id, s["t"]), (self.m_by_id, s["tm"]), (self.v_by_id, s["tv"])):
            for a, b in zip(dst, src):
                if a is not None:
                    a.copy_(b)

    # ---- public API -------------------------------------------------------------------------------------
    def load_blob(self, src: torch.Tensor):
        """One copy: pinned host blob or device pool slot -> the static batch the graph reads."""
        self.blob.copy_(src, non_blocking=True)

    def load_batch(self, batch: Dict[str, torch.Tensor]):
        for key, dt, shape, off in self.layout.fields:
            src = batch[key]
            if key == "label" and src.numel() != self.B * self.layout.n_labels:
                raise L.NrxError(f"label has {src.numel() // max(self.B, 1)} column(s) per sample, this trainer was built for "
                                 f"n_labels={self.layout.n_labels} (FusedTrainer(..., n_labels=))")
            self.batch[key].copy_(src.to(device=self.dev, dtype=dt).view(*shape), non_blocking=True)

    def load_rows(self, device_file, rows: Optional[torch.Tensor] = None, start: int = 0):
        """Assemble the batch on the GPU from a device-resident feature file (ingest.DeviceFeatureFile): rows `rows`
        (device int64[B], e.g. a slice of a device-side permutation) or the contiguous rows [start, start + B)."""
        device_file.assemble(self.layout, self.blob, rows=rows, start=start, status=self.id_status)

    _STATUS_EVERY = 16   # steps between two asynchronous read-backs of the status word

    def step(self) -> torch.Tensor:
        """Run one training step on the currently loaded batch; returns the (device) mean BCE loss.
        Raises NrxError if an EARLIER step flagged an out-of-table id or a dead peer exchange: the status word is read
        back asynchronously every few steps on a side stream, so the check never stalls the step it follows."""
        self._poll_status()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        self._post_status()
        return self.loss

    def _post_status(self):
        st = self._stat
        st["n"] += 1
        if st["n"] % self._STATUS_EVERY:
            return
        if st["posted"] and not st["done"].query():
            return   # the previous read-back is still in flight
        st["posted"] = True
        st["ev"].record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(st["stream"]):
            st["stream"].wait_event(st["ev"])
            st["host"].copy_(self._loss_status, non_blocking=True)
            st["done"].record(st["stream"])

    def _poll_status(self):
        st = self._stat
        if st["posted"] and st["done"].query():
            self._raise_for_status(int(st["host"].view(torch.int32)[1]))

    @staticmethod
    def _raise_for_status(bits: int):
        if bits & 2:
            raise L.NrxError("the peer-memory gradient exchange (K7) timed out waiting for a rank: the exchange is dead on "
                             "every rank and no parameter was updated since; restart from the last checkpoint")
        if bits & 4:
            raise L.NrxError("load_rows(): a row index outside the device-resident feature file was requested (the sample was "
                             "assembled as padding with label 0)")
        if bits & 1:
            raise L.NrxError("a feature id outside its embedding table reached the GPU (vocabulary / config mismatch): "
                             "the reference's nn.Embedding raises here (base_model.py:271)")

    def train_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        self.load_batch(batch)
        return self.step()

    # ---- pipelined feeding from pinned host memory ---------------------------------------------------
    def feed(self, host_blob: torch.Tensor) -> Optional[float]:
        """One host-fed training step, software-pipelined one deep.

        Enqueues (i) the H2D copy of this step's pinned blob on a copy stream — it overlaps the kernels of the
        step launched by the previous call —, (ii) the step, (iii) the D2H read of its loss into pinned memory;
        then returns the loss of the PREVIOUS fed step (None on the first call), which is already on the host, so
        the CPU never stalls on the step it just launched.  `drain()` returns the last loss."""
        f = self._feed_state()
        k = f["i"] & 1
        f["i"] += 1
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(f["copy"]):
            f["copy"].wait_event(f["free"][k])           # the step that read stage[k] two calls ago has consumed it
            f["stage"][k].copy_(host_blob, non_blocking=True)
            f["ready"][k].record(f["copy"])
        main.wait_event(f["ready"][k])
        self.blob.copy_(f["stage"][k], non_blocking=True)
        f["free"][k].record(main)
        self.step()
        f["loss_host"][k].copy_(self._loss_status, non_blocking=True)
        f["done"][k].record(main)
        prev, f["pending"] = f["pending"], k
        if prev is None:
            return None
        f["done"][prev].synchronize()
        return self._host_loss(f["loss_host"][prev])

    def drain(self) -> Optional[float]:
        """Loss of the last fed step (waits for it)."""
        f = self._feed_state()
        prev, f["pending"] = f["pending"], None
        if prev is None:
            return None
        f["done"][prev].synchronize()
        return self._host_loss(f["loss_host"][prev])

    @classmethod
    def _host_loss(cls, buf: torch.Tensor) -> float:
        cls._raise_for_status(int(buf.view(torch.int32)[1]))
        return float(buf[0])

    def check_status(self):
        """Synchronises and raises if any step so far saw an id outside its table (K1's status word) or a dead
        peer exchange (K7)."""
        self._raise_for_status(int(self.id_status.item()))

    check_ids = check_status

    # ---- optimizer state (resume) -----------------------------------------------------------------
    def optimizer_state_dict(self) -> Dict[str, torch.Tensor]:
        """AdamW moments + step counter (what a Lightning checkpoint of the reference keeps under 'optimizer_states')."""
        sd = {"step": self.d_step.clone(), "flat_m": self.flat_m.clone(), "flat_v": self.flat_v.clone()}
        for t in range(L.NRX_MAX_TABLES):
            if self.m_by_id[t] is not None:
                sd[f"table_m.{t}"] = self.m_by_id[t].clone()
                sd[f"table_v.{t}"] = self.v_by_id[t].clone()
        return sd

    def load_optimizer_state_dict(self, sd: Dict[str, torch.Tensor]):
        self.d_step.copy_(sd["step"])
        self.flat_m.copy_(sd["flat_m"])
        self.flat_v.copy_(sd["flat_v"])
        for t in range(L.NRX_MAX_TABLES):
            if self.m_by_id[t] is not None:
                self.m_by_id[t].copy_(sd[f"table_m.{t}"])
                self.v_by_id[t].copy_(sd[f"table_v.{t}"])

    def _feed_state(self):
        f = getattr(self, "_feed", None)
        if f is None:
            ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
            f = dict(copy=torch.cuda.Stream(device=self.dev), stage=[torch.empty_like(self.blob) for _ in range(2)],
                     loss_host=[torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)],
                     ready=ev(), free=ev(), done=ev(), i=0, pending=None)
            for e in f["free"]:
                e.record(torch.cuda.current_stream(self.dev))
            self._feed = f
        return f
