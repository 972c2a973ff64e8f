#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_models.py tests/test_gpu_trainer.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -n 3
timeout -s KILL 300 python tools/profile_kernels.py --only fm --sizes 16384,65536,262144,1048576 2>&1 | grep "K2"
