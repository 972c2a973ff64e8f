#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_gpu_dssm.py tests/test_gpu_valmetrics.py -x -q -m gpu 2>&1 | tail -n 25
