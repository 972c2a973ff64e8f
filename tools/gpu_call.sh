#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_dssm.py -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 300 python tools/topk_peer_profile.py --emulate 8 2>&1 | grep -v Warn | tail -13
timeout -s KILL 300 python tools/topk_peer_profile.py 2>&1 | grep -v Warn | tail -13
