#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
tail -n 3 gpurun_out/r2_bench_n1_final.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_n1_final.json')); r=j['retrieval']
print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['kernel'], round(j['roofline']['frac'],4))
print('  ', {k:(round(v['value']/1e6,1), round(v['ms_per_step'],4)) for k,v in j['legs'].items()}, r['value'], r['ms_per_search'], r['e2e']['value'], r['q1_latency_ms'])
print(j['variants'])
PY
