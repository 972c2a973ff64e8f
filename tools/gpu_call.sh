#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_trainer.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -n 3
NRX_BENCH_LEGS=cfg1 timeout -s KILL 600 python bench.py --no-retrieval --steps 50 --warmup 5 --cpu-steps 1 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); v=j['legs']['cfg1_deep_hist']; print(v['value'], v['ms_per_step']); [print('  ',k,x) for k,x in v['kernels'].items()]"
