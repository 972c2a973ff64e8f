#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "peer_memory or dp2" 2>&1 | tail -n 3
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/topk_peer_profile.py 2>&1 | grep -v "Warn\|^\*\*\|OMP" | tail -n 13
