#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "dssm or peer_memory" 2>&1 | tail -n 15
