"""SURVEY §8(d), last line: "the reference source on the same B200 (eager torch) as the GPU-side comparison".
The reference itself cannot travel to the GPU box, so this times its CPU restatement (oracle/ref_path.py — plain eager
PyTorch ops, device-agnostic) with the tensors on the GPU: forward + BCE + backward + torch.optim.AdamW per step, the
same workload and step definition as bench.py.  A measurement tool (like bench.py's CPU arm), never a product path.

    python tests/tools/eager_gpu_baseline.py [--device cuda] [--steps 50] [--workload deepfm]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from news_recsys_b200.synthetic import synth_batch  # noqa: E402
from oracle import ref_path as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="deepfm")
    a = ap.parse_args()
    kind, cfg, B, desc = bench.workload_cfg(a.workload)
    dev = torch.device(a.device)
    torch.manual_seed(42)
    model = bench.model_class(kind)(cfg)
    leaf = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(leaf.values()), lr=cfg["train_hparams"]["lr"], betas=(0.9, 0.999))
    batches = [{k: v.to(dev) for k, v in synth_batch(cfg, B, seed=1000 + i).items()} for i in range(4)]

    def step(b):
        opt.zero_grad(set_to_none=True)
        loss = R.bce(R.model_forward(kind, leaf, cfg, b, dcn_materialise=False), b["label"][:, 0])
        loss.backward()
        opt.step()
        return loss

    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    for i in range(a.warmup):
        step(batches[i % 4])
    sync()
    t0 = time.perf_counter()
    for i in range(a.steps):
        loss = step(batches[i % 4])
    sync()
    dt = (time.perf_counter() - t0) / a.steps
    print(json.dumps({"what": "eager PyTorch restatement of the reference step on " + str(dev), "workload": desc, "batch": B,
                      "ms_per_step": dt * 1e3, "samples_per_s": B / dt, "steps": a.steps, "final_loss": float(loss)}))


if __name__ == "__main__":
    main()
