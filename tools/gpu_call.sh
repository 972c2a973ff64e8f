#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py -q -m gpu 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_gpu_tests_p.txt
grep -n "Error\|error\|assert\|passed\|failed" gpurun_out/r2_gpu_tests_p.txt | head -30
