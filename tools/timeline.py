"""Kernel timeline of ONE graph-replayed training step (start / duration / stream per kernel) through torch.profiler
(CUPTI via kineto; no nsys in the image).  Answers what the per-API event timings cannot: which kernels of the forked
streams actually overlap the forward/backward chain, and where the gaps are.  Dev tool, not a bench line.

    python tools/timeline.py [--workload deepfm] [--table-update dense] > gpurun_out/timeline.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from news_recsys_b200.synthetic import synth_batch  # noqa: E402
from news_recsys_b200.trainer import FusedTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="deepfm")
    ap.add_argument("--table-update", default="dense")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    kind, cfg, B, _ = bench.workload_cfg(a.workload)
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    tr = FusedTrainer(bench.model_class(kind)(cfg).to(dev), B, kind=kind, table_update=a.table_update)
    blobs = []
    for i in range(4):
        hb = torch.empty(tr.layout.nbytes, dtype=torch.uint8)
        tr.layout.pack(synth_batch(cfg, B, seed=42 + i), hb)
        blobs.append(hb.to(dev))
    for i in range(20):
        tr.load_blob(blobs[i % 4]); tr.step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            tr.load_blob(blobs[i % 4]); tr.step()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append({"name": e.name[:80], "start_us": e.time_range.start, "dur_us": e.time_range.elapsed_us(),
                       "stream": getattr(e, "device_resource_id", None) if hasattr(e, "device_resource_id") else None})
    ev.sort(key=lambda x: x["start_us"])
    if ev:
        t0 = ev[0]["start_us"]
        for x in ev:
            x["start_us"] -= t0
    print(json.dumps(ev, indent=0))


if __name__ == "__main__":
    main()
