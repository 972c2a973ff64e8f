#!/bin/bash
bash tools/capture_profiles.sh
python tools/summarize_profiles.py gpurun_out gpurun_out > gpurun_out/summarize.log 2>&1
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_n1_final.json')); r=j['retrieval']
print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['kernel'], j['roofline']['bound'], round(j['roofline']['frac'],4), j['roofline']['traffic'])
print('  ', {k:(round(v['value']/1e6,1), round(v['ms_per_step'],4)) for k,v in j['legs'].items()})
print('  ', r['value'], r['ms_per_search'], r['e2e']['value'], r['q1_latency_ms'], r['roofline']['frac'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
