#!/bin/bash
timeout -s KILL 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -n 2
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-legs --no-retrieval --steps 200 --warmup 20 --cpu-steps 1 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['kernels']['nrx_adamw_dense_dev'])"
done
