#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_dssm.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -n 2
timeout -s KILL 300 python tools/profile_kernels.py --only topk 2>&1 | grep K6
