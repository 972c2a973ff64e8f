"""Batch ingestion: a columnar binary feature file and a batch assembler that writes straight into the
trainer's pinned batch blob.

What it replaces: `DataReader.__getitem__` (src/dataset/DataReader/data_reader.py:54-114) — per SAMPLE: split on tab /
space / ':' / ',', build Python lists, pad, one `torch.tensor` per feature — followed by torch's default collate in
`MINDDataModule.train_dataloader` (pl_dataloader.py:77-95).  At a 0.12 ms training step that loader is the end-to-end
bottleneck by three orders of magnitude.

  compile_feature_file(config, text, out)   one pass over the reference's text format
                                             ("name:value name:v1,v2,... \\t label label"), same parsing rules and the
                                             same errors, into `out` (.nrxf):
                                               sparse feature  -> int32 column [n_rows]
                                               dense feature   -> float64 column [n_rows]
                                               array feature   -> CSR: int64 offsets [n_rows + 1] + int32 values, already
                                                                  truncated to the first max_len ids (data_reader.py:103-105)
                                               labels          -> float32 [n_rows, n_labels]
  FeatureFile(path)                          memory-mapped reader
      .batch(rows | start, B)                the dict the reference's DataLoader yields (int64 ids, [B, L] + `<name>_mask`)
      .pack(layout, blob, rows | start)      the same batch written into a (pinned) blob with the trainer's BatchLayout:
                                             three host copies per feature through libnrx (`nrx_ingest_*`), no temporaries

File layout: b"NRXF0001" | u64 header bytes | JSON header | 64-byte aligned sections (offsets in the header).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import yaml

from . import _lib as L

MAGIC = b"NRXF0001"


def _al(x, a=64):
    return (x + a - 1) // a * a


def parse_line(raw_line: str, idx: int, sparse, dense, arrays, array_max_length) -> dict:
    """One line -> {name: int | float | list[int] (truncated), 'label': list[float]} with the reference's rules and
    error messages (data_reader.py:58-112)."""
    try:
        feature_part, label_part = raw_line.split("\t")
    except ValueError:
        raise ValueError(f"Line {idx} format error: missing tab separator between features and labels.")
    out = {}
    for item in feature_part.split(" "):
        if ":" not in item:
            raise ValueError(f"Feature item format error: '{item}' does not contain ':' separator.")
        name, val = item.split(":", 1)
        if name in sparse:
            out[name] = int(val)
        elif name in dense:
            out[name] = float(val)
        elif name in arrays:
            max_len = array_max_length.get(name)
            if max_len is None:
                raise ValueError(f"Max length for array feature '{name}' missing in config.")
            ids = [int(x) for x in val.split(",")] if val else []
            out[name] = ids[:max_len]
    out["label"] = [float(l) for l in label_part.strip().split(" ")]
    return out


def compile_feature_file(config_path: str, text_path: str, out_path: str) -> dict:
    cfg = yaml.safe_load(open(config_path))
    feats = cfg["features"]
    sparse = set(feats.get("sparse_feature_names") or [])
    dense = set(feats.get("dense_feature_names") or [])
    arrays = set(feats.get("array_feature_names") or [])
    amax = dict(feats.get("array_max_length") or {})
    if not os.path.exists(text_path):
        raise FileNotFoundError(f"Data file not found: {text_path}")
    with open(text_path, "r", encoding="utf-8") as f:
        lines = [line.strip() for line in f if line.strip()]
    rows = [parse_line(l, i, sparse, dense, arrays, amax) for i, l in enumerate(lines)]
    n = len(rows)
    names = [k for k in rows[0] if k != "label"] if n else []
    n_labels = len(rows[0]["label"]) if n else 0
    for i, r in enumerate(rows):
        if set(r) - {"label"} != set(names):
            raise ValueError(f"Line {i}: features {sorted(set(r) - {'label'})} differ from line 0 {sorted(names)} "
                             "(the reference's default collate would fail on this batch)")
        if len(r["label"]) != n_labels:
            raise ValueError(f"Line {i}: {len(r['label'])} labels, line 0 has {n_labels}")
    sections, cols = [], []

    def add(arr):
        sections.append(np.ascontiguousarray(arr))
        return len(sections) - 1

    for name in sorted(names):
        if name in sparse:
            col = np.array([r[name] for r in rows], dtype=np.int64)
            if n and (col.min() < -2**31 or col.max() >= 2**31):
                raise ValueError(f"feature {name}: id outside int32")
            cols.append({"name": name, "kind": "sparse", "data": add(col.astype(np.int32))})
        elif name in dense:
            cols.append({"name": name, "kind": "dense", "data": add(np.array([r[name] for r in rows], dtype=np.float64))})
        else:
            lens = np.array([len(r[name]) for r in rows], dtype=np.int64)
            off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            vals = np.fromiter((x for r in rows for x in r[name]), dtype=np.int64, count=int(off[-1]))
            if vals.size and (vals.min() < -2**31 or vals.max() >= 2**31):
                raise ValueError(f"feature {name}: id outside int32")
            cols.append({"name": name, "kind": "array", "max_len": int(amax[name]), "offsets": add(off),
                         "data": add(vals.astype(np.int32))})
    lab = add(np.array([r["label"] for r in rows], dtype=np.float32).reshape(n, max(n_labels, 0)))
    header = {"n_rows": n, "n_labels": n_labels, "columns": cols, "labels": lab, "sections": []}
    # two passes: section offsets depend on the header length
    def layout(hlen):
        pos = _al(len(MAGIC) + 8 + hlen)
        out = []
        for s in sections:
            out.append({"offset": pos, "nbytes": int(s.nbytes), "dtype": str(s.dtype), "shape": list(s.shape)})
            pos = _al(pos + s.nbytes)
        return out, pos
    header["sections"], _ = layout(0)
    hbytes = json.dumps(header).encode()
    header["sections"], total = layout(len(hbytes) + 64)   # room for the offsets growing by a few digits
    hbytes = json.dumps(header).encode().ljust(len(hbytes) + 64)
    with open(out_path, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(hbytes)).tobytes())
        f.write(hbytes)
        for s, meta in zip(sections, header["sections"]):
            f.seek(meta["offset"])
            f.write(s.tobytes())
        f.truncate(max(total, f.tell()))
    return {"n_rows": n, "n_labels": n_labels, "features": sorted(names), "bytes": total}


class FeatureFile:
    def __init__(self, path: str):
        self.path = path
        self.mm = np.memmap(path, dtype=np.uint8, mode="r")
        if bytes(self.mm[:8]) != MAGIC:
            raise ValueError(f"{path}: not an NRXF feature file")
        hlen = int(np.frombuffer(self.mm[8:16], dtype=np.uint64)[0])
        h = json.loads(bytes(self.mm[16:16 + hlen]).decode())
        self.n_rows, self.n_labels = int(h["n_rows"]), int(h["n_labels"])
        self._sec = [self._view(m) for m in h["sections"]]
        self.columns = {c["name"]: c for c in h["columns"]}
        self.labels = self._sec[h["labels"]]
        self.lib = L.load()

    def _view(self, m):
        a = np.frombuffer(self.mm, dtype=np.dtype(m["dtype"]), count=int(np.prod(m["shape"])) if m["shape"] else 1,
                          offset=m["offset"])
        return a.reshape(m["shape"])

    def __len__(self):
        return self.n_rows

    # ---- raw copies through the C ABI -----------------------------------------------------------------------
    @staticmethod
    def _rows(rows, start, B):
        if rows is None:
            return None, int(start), int(B)
        r = np.ascontiguousarray(np.asarray(rows, dtype=np.int64))
        return r, 0, int(r.shape[0])

    def _ids(self, name, rows, start, B, out_ptr, idt):
        c = self.columns[name]
        L.check(self.lib.nrx_ingest_gather_ids(self._sec[c["data"]].ctypes.data, self.n_rows,
                                               None if rows is None else rows.ctypes.data, start, B, out_ptr, idt),
                "nrx_ingest_gather_ids")

    def _array(self, name, rows, start, B, ids_ptr, idt, mask_ptr):
        c = self.columns[name]
        L.check(self.lib.nrx_ingest_csr_expand(self._sec[c["offsets"]].ctypes.data, self._sec[c["data"]].ctypes.data, self.n_rows,
                                               None if rows is None else rows.ctypes.data, start, B, c["max_len"], ids_ptr, idt,
                                               mask_ptr), "nrx_ingest_csr_expand")

    def _labels(self, rows, start, B, out_ptr, out_ld):
        L.check(self.lib.nrx_ingest_gather_labels(self.labels.ctypes.data, self.n_rows, self.n_labels,
                                                  None if rows is None else rows.ctypes.data, start, B, out_ptr, out_ld),
                "nrx_ingest_gather_labels")

    # ---- the reference DataLoader's batch ------------------------------------------------------------------------
    def batch(self, rows: Optional[Sequence[int]] = None, start: int = 0, B: Optional[int] = None,
              id_dtype=torch.int64) -> Dict[str, torch.Tensor]:
        """default_collate([DataReader[i] for i in rows]) — same keys, dtypes (int64 ids unless asked), shapes."""
        r, start, B = self._rows(rows, start, self.n_rows - start if B is None and rows is None else B)
        idt = L.IDX_I64 if id_dtype == torch.int64 else L.IDX_I32
        out = {}
        for name, c in self.columns.items():
            if c["kind"] == "sparse":
                t = torch.empty(B, dtype=id_dtype)
                self._ids(name, r, start, B, t.data_ptr(), idt)
                out[name] = t
            elif c["kind"] == "dense":
                col = self._sec[c["data"]]
                out[name] = torch.from_numpy(np.array(col[r] if r is not None else col[start:start + B]))
            else:
                t = torch.empty((B, c["max_len"]), dtype=id_dtype)
                m = torch.empty((B, c["max_len"]), dtype=torch.float32)
                self._array(name, r, start, B, t.data_ptr(), idt, m.data_ptr())
                out[name], out[name + "_mask"] = t, m
        lab = torch.empty((B, self.n_labels), dtype=torch.float32)
        self._labels(r, start, B, lab.data_ptr(), self.n_labels)
        out["label"] = lab
        return out

    # ---- straight into the trainer's blob -------------------------------------------------------------------------
    def pack(self, layout, blob: torch.Tensor, rows: Optional[Sequence[int]] = None, start: int = 0) -> torch.Tensor:
        """Write the batch (layout.B rows) into `blob` (host uint8 tensor, ideally pinned) in `trainer.BatchLayout`."""
        if blob.device.type != "cpu" or blob.dtype != torch.uint8 or blob.numel() < layout.nbytes:
            raise L.NrxError("pack() needs a host uint8 blob of at least layout.nbytes bytes")
        r, start, B = self._rows(rows, start, layout.B)
        if B != layout.B:
            raise L.NrxError(f"{B} rows for a layout of batch {layout.B}")
        base = blob.data_ptr()
        offs = {key: (dt, shape, off) for key, dt, shape, off in layout.fields}
        for key, (dt, shape, off) in offs.items():
            if key == "label":
                self._labels(r, start, B, base + off, shape[1])
            elif key.endswith("_mask") and key[:-5] in self.columns:
                continue   # written together with its ids
            elif key not in self.columns:
                raise L.NrxError(f"feature '{key}' of the batch layout is not in {self.path}")
            else:
                idt = L.IDX_I64 if dt == torch.int64 else L.IDX_I32
                c = self.columns[key]
                if c["kind"] == "array":
                    if shape[1] != c["max_len"]:
                        raise L.NrxError(f"feature '{key}': layout length {shape[1]} != file max_len {c['max_len']}")
                    self._array(key, r, start, B, base + off, idt, base + offs[key + "_mask"][2])
                else:
                    self._ids(key, r, start, B, base + off, idt)
        return blob


class BlobPrefetcher:
    """Blobs of successive batches, packed ahead by worker threads (one whole batch per task; the C calls release the
    GIL) into a ring of `depth` host buffers (pinned when CUDA is available) — the producer side of
    `FusedTrainer.feed()`.  `batches`: iterable of row-index arrays (shuffled epochs) or ints (start row of a contiguous
    batch).  A slot is refilled only after the consumer has taken `depth - 2` later blobs, i.e. after the `feed()` that
    copied it has been followed by two more — `feed()` waits for the previous step on every call, so that copy is done."""

    def __init__(self, ff: FeatureFile, layout, batches, depth: int = 4, workers: int = 2, pin: Optional[bool] = None):
        import threading
        if depth < 3:
            raise L.NrxError("BlobPrefetcher needs depth >= 3")
        self.ff, self.layout, self.depth = ff, layout, depth
        self.batches = list(batches)
        pin = torch.cuda.is_available() if pin is None else pin
        self.ring = [torch.zeros(layout.nbytes, dtype=torch.uint8) for _ in range(depth)]
        if pin:
            self.ring = [b.pin_memory() for b in self.ring]
        self._cv = threading.Condition()
        self._done = [False] * len(self.batches)
        self._taken = 0
        self._next = 0
        self._err = None
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, workers))]
        for t in self._threads:
            t.start()

    def _work(self):
        while True:
            with self._cv:
                while True:
                    i = self._next
                    if i >= len(self.batches) or self._err is not None:
                        return
                    if i - self._taken <= self.depth - 2:   # slot i % depth is free again
                        self._next += 1
                        break
                    self._cv.wait()
            try:
                b = self.batches[i]
                if isinstance(b, (int, np.integer)):
                    self.ff.pack(self.layout, self.ring[i % self.depth], start=int(b))
                else:
                    self.ff.pack(self.layout, self.ring[i % self.depth], rows=b)
            except Exception as e:   # surfaced to the consumer
                with self._cv:
                    self._err = e
                    self._cv.notify_all()
                return
            with self._cv:
                self._done[i] = True
                self._cv.notify_all()

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        for i in range(len(self.batches)):
            with self._cv:
                while not self._done[i] and self._err is None:
                    self._cv.wait()
                if self._err is not None:
                    raise self._err
                self._taken = i   # blobs < i - ... may be refilled (see class docstring)
                self._cv.notify_all()
            yield self.ring[i % self.depth]


class DeviceFeatureFile:
    """The columnar file resident in HBM (MIND-small: 12 MB; a billion-click log: tens of GB of 180) and batches
    assembled ON the GPU straight into the trainer's static blob — no host work and no H2D copy per step beyond the
    optional row permutation.  Same output as FeatureFile.pack()."""

    def __init__(self, ff: FeatureFile, device):
        if torch.device(device).type != "cuda":
            raise L.NrxError("DeviceFeatureFile needs a CUDA device (the host path is FeatureFile.pack)")
        self.ff, self.dev = ff, torch.empty(0, device=device).device   # normalised: "cuda" -> cuda:<current>
        up = lambda a: torch.from_numpy(np.array(a)).to(self.dev)
        self.n_rows, self.n_labels = ff.n_rows, ff.n_labels
        self.cols = {}
        for name, c in ff.columns.items():
            if c["kind"] == "dense":
                continue
            self.cols[name] = dict(kind=c["kind"], data=up(ff._sec[c["data"]]), max_len=c.get("max_len", 1),
                                   offsets=up(ff._sec[c["offsets"]]) if c["kind"] == "array" else None)
        self.labels = up(ff.labels)
        self.lib = L.load()

    def assemble(self, layout, blob: torch.Tensor, rows: Optional[torch.Tensor] = None, start: int = 0,
                 status: Optional[torch.Tensor] = None):
        """Fill `blob` (device uint8, trainer.BatchLayout) with rows `rows` (device int64[B]) or [start, start + B).
        `status` (device int32[1], optional): bit 2 is set when an element of `rows` lies outside the file (such a sample
        is assembled as all-padding with label 0; FusedTrainer.load_rows passes its status word and raises)."""
        if blob.device != self.dev or blob.dtype != torch.uint8 or blob.numel() < layout.nbytes:
            raise L.NrxError("assemble() needs a device uint8 blob of at least layout.nbytes bytes on the file's device")
        B = layout.B
        if rows is not None:
            if rows.device != self.dev or rows.dtype != torch.int64 or rows.numel() != B:
                raise L.NrxError(f"rows must be a device int64 tensor of {B} elements")
        elif not (0 <= start and start + B <= self.n_rows):
            raise L.NrxError(f"rows [{start}, {start + B}) outside the file ({self.n_rows} rows)")
        base = blob.data_ptr()
        offs = {key: (dt, shape, off) for key, dt, shape, off in layout.fields}
        names = [k for k in offs if k != "label" and not (k.endswith("_mask") and k[:-5] in self.cols)]
        arr = (L.NrxIngestCol * len(names))()
        for i, key in enumerate(names):
            if key not in self.cols:
                raise L.NrxError(f"feature '{key}' of the batch layout is not in {self.ff.path}")
            dt, shape, off = offs[key]
            c = self.cols[key]
            arr[i].data = c["data"].data_ptr()
            arr[i].idx_dtype = L.IDX_I64 if dt == torch.int64 else L.IDX_I32
            arr[i].out_ids = base + off
            if c["kind"] == "array":
                if shape[1] != c["max_len"]:
                    raise L.NrxError(f"feature '{key}': layout length {shape[1]} != file max_len {c['max_len']}")
                arr[i].offsets, arr[i].L = c["offsets"].data_ptr(), shape[1]
                arr[i].out_mask = base + offs[key + "_mask"][2]
            else:
                arr[i].offsets, arr[i].L, arr[i].out_mask = None, 1, None
        ldt, lshape, loff = offs["label"]
        L.check(self.lib.nrx_ingest_assemble_device(arr, len(names), self.labels.data_ptr(), self.n_labels, base + loff, lshape[1],
                                                    self.n_rows, None if rows is None else rows.data_ptr(), int(start), B,
                                                    L.ptr(status), L.stream_ptr(self.dev)), "nrx_ingest_assemble_device")
        return blob
