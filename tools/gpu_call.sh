#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-legs --no-retrieval --steps 200 --warmup 20 --cpu-steps 1 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['launches_per_step'], j.get('variants'))"
done
NRX_DW_AFTER_MERGE=1 timeout -s KILL 600 python bench.py --no-legs --no-retrieval --steps 200 --warmup 20 --cpu-steps 1 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('dw after merge', j['value'], j['ms_per_step'], j['e2e']['value'], j['launches_per_step'], j.get('variants'))"
