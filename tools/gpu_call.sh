#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_n8.json')); r=j['retrieval']
print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['parity']['parity_ok'])
print('  ', {k:(round(v['value']/1e6,1), round(v['ms_per_step'],4)) for k,v in j['legs'].items()}, r['value'], r['ms_per_search'], r['e2e']['value'])
PY
