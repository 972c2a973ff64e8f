"""Per-kernel device times of the sharded peer search (torchrun, one rank per GPU), from torch.profiler (CUPTI).
usage: torchrun --nproc-per-node G tools/topk_peer_profile.py [N] [Q] [k]
Also works as a single process (world 1) to look at ONE shard of a larger job: `--emulate G` shrinks the corpus to N / G rows
and sets the per-shard threshold rank a G-rank job would use."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from news_recsys_b200.parallel import ShardedTopk, shard_range  # noqa: E402


def main():
    argv = sys.argv[1:]
    emulate = 0
    if "--emulate" in argv:
        j = argv.index("--emulate")
        emulate = int(argv[j + 1])
        del argv[j:j + 2]
    args = argv
    N = int(args[0]) if len(args) > 0 else 1_000_000
    Q = int(args[1]) if len(args) > 1 else 1024
    k = int(args[2]) if len(args) > 2 else 100
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    kp = 0
    if emulate:
        N = N // emulate
        kp = (2 * k + 64 + emulate - 1) // emulate + 24
    g = torch.Generator(device=dev).manual_seed(5)
    lo, hi = shard_range(N, rank, world)
    full = torch.nn.functional.normalize(torch.randn((N, 128), generator=g, device=dev), dim=1)
    qs = [torch.nn.functional.normalize(torch.randn((Q, 128), generator=g, device=dev), dim=1) for _ in range(4)]
    st = ShardedTopk(full[lo:hi].contiguous(), N, exchange="peer", kprime=kp)
    del full
    for i in range(5):
        st.search_peer_(qs[i % 4], k)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        st.search_peer_(qs[i % 4], k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(10):
            st.search_peer_(qs[i % 4], k)
        torch.cuda.synchronize()
    if rank == 0:
        print(f"world {world} emulate {emulate} N_local {hi - lo} Q {Q} k {k} kprime {kp}: {ms * 1e3:.1f} us per search, "
              f"{Q / ms / 1e3:.2f} M queries/s, exact fallbacks {st.exact_fallbacks(Q, k)}")
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        for e in rows[:14]:
            print(f"  {e.device_time_total / 10:9.1f} us/search  x{e.count / 10:.0f}  {e.key[:110]}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
