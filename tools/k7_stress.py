"""K7 stress: N ranks train for many steps while random ranks are delayed by random amounts before
random steps (device-side spin, so the skew hits the flag barriers inside the captured step, not the host), then every
rank's parameters are compared bitwise with rank 0's and with a run WITHOUT delays (same batches).

    torchrun --nproc-per-node 2 tools/k7_stress.py [steps] [max_delay_us]

Prints one JSON line on rank 0.  Any time-out, divergence or mismatch is a non-zero exit."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from news_recsys_b200.model.sort.deepfm.model import DeepFM  # noqa: E402
from news_recsys_b200.parallel import DataParallelTrainer  # noqa: E402
from news_recsys_b200.synthetic import mind_config, synth_batch  # noqa: E402


def run(rank, world, dev, steps, max_delay_us, delays):
    rows = {"user_id": 3000, "item_id": 2000, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("deepfm", rows)
    B = 256
    torch.manual_seed(1)
    model = DeepFM(cfg).to(dev)
    tr = DataParallelTrainer(model, B, kind="deepfm", table_update="dense", exchange="peer")
    blobs = []
    for i in range(8):
        full = synth_batch(cfg, B * world, seed=100 + i, label_p=0.5)
        hb = torch.empty(tr.layout.nbytes, dtype=torch.uint8)
        tr.layout.pack({k: v[rank * B:(rank + 1) * B] for k, v in full.items()}, hb)
        blobs.append(hb.to(dev))
    g = torch.Generator().manual_seed(1234 + rank)
    cycles_per_us = 1900
    t0 = time.perf_counter()
    for s in range(steps):
        if delays and float(torch.rand((), generator=g)) < 0.05:      # this rank falls behind before 5 % of its steps
            torch.cuda._sleep(int(float(torch.rand((), generator=g)) * max_delay_us * cycles_per_us))
        tr.load_blob(blobs[s % 8])
        tr.step()
    torch.cuda.synchronize(dev)
    tr.check_status()                                                 # raises on a dead exchange
    assert not tr.peer_timed_out()
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone(), time.perf_counter() - t0


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    max_delay_us = float(sys.argv[2]) if len(sys.argv) > 2 else 3000.0
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    p_delay, secs = run(rank, world, dev, steps, max_delay_us, True)
    p_plain, _ = run(rank, world, dev, steps, max_delay_us, False)
    ref = p_delay.clone()
    dist.broadcast(ref, src=0)
    ok = torch.tensor([int(torch.equal(ref, p_delay)), int(torch.equal(p_delay, p_plain)), int(torch.isfinite(p_delay).all())],
                      device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"tool": "k7_stress", "world": world, "steps": steps, "max_delay_us": max_delay_us,
                          "delayed_step_fraction_per_rank": 0.05, "replicas_bitwise_equal": bool(ok[0]),
                          "equal_to_run_without_delays": bool(ok[1]), "finite": bool(ok[2]), "seconds": round(secs, 2)}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if bool(ok.min()) else 1)


if __name__ == "__main__":
    main()
