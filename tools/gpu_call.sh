#!/bin/bash
for v in 0 1 0 1; do
if [ $v = 1 ]; then export NRX_K3_CW32=1; else unset NRX_K3_CW32; fi
NRX_BENCH_LEGS=cfg5 timeout -s KILL 600 python bench.py --no-retrieval --steps 100 --warmup 10 --cpu-steps 1 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('cw32=$v', {k:(round(v['ms_per_step'],4), v['kernels'].get('nrx_embed_bwd_apply',{}).get('us_per_step')) for k,v in j['legs'].items()})"
done
