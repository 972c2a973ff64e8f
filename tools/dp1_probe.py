"""Isolate K7 / peer-memory overhead: DataParallelTrainer with world_size 1 vs FusedTrainer (dense, flat) on one GPU."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from news_recsys_b200.trainer import FusedTrainer
from news_recsys_b200.parallel import DataParallelTrainer
from news_recsys_b200.synthetic import synth_batch

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
kind, cfg, B, _ = bench.workload_cfg("deepfm")
dev = torch.device("cuda", 0)

def run(tr, label):
    blobs = []
    for i in range(8):
        hb = torch.empty(tr.layout.nbytes, dtype=torch.uint8)
        tr.layout.pack(synth_batch(cfg, B, seed=42 + i), hb)
        blobs.append(hb.to(dev))
    for i in range(30):
        tr.load_blob(blobs[i % 8]); tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(300):
        tr.load_blob(blobs[i % 8]); tr.step()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:40s} {e0.elapsed_time(e1) / 300 * 1e3:8.1f} us/step", flush=True)

torch.manual_seed(42)
run(FusedTrainer(bench.model_class(kind)(cfg).to(dev), B, kind=kind, table_update="dense", dense_impl="flat"), "FusedTrainer dense flat")
torch.manual_seed(42)
run(DataParallelTrainer(bench.model_class(kind)(cfg).to(dev), B, kind=kind, exchange="peer"), "DP world=1 peer (K7, peer memory)")
torch.manual_seed(42)
run(DataParallelTrainer(bench.model_class(kind)(cfg).to(dev), B, kind=kind, exchange="nccl"), "DP world=1 nccl (2 graphs)")
dist.destroy_process_group()
