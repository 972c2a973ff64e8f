#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n1_final.json','gpurun_out/r2_bench_n2_final.json'):
    j=json.load(open(f)); r=j['retrieval']
    print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'])
    print('  ', {k:(round(v['value']/1e6,1), round(v['ms_per_step'],4)) for k,v in j['legs'].items()}, r['value'], r['ms_per_search'], r['e2e']['value'])
PY
