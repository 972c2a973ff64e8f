#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -n 3
for v in 1 0; do
NRX_DW_LOW_PRIO=$v timeout -s KILL 600 python bench.py --no-legs --no-retrieval --steps 200 --warmup 20 --cpu-steps 1 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('dw_low=$v', j['value'], j['ms_per_step'], j['e2e']['value'], j['variants']['table_update=sparse']['ms_per_step'])"
done
NRX_DW_LOW_PRIO=1 timeout -s KILL 300 python tools/timeline.py --steps 3 > gpurun_out/r2_timeline.json 2> gpurun_out/timeline.err
python - <<'PY'
import json
ev=json.load(open('gpurun_out/r2_timeline.json'))
idx=[i for i,e in enumerate(ev) if 'hparams' in e['name']]
a=idx[-1]
t0=ev[a]['start_us']
for e in ev[max(0,a-2):]:
    print(f"{e['start_us']-t0:8.1f} +{e['dur_us']:6.1f}  s{e['stream']}  {e['name'][:70]}")
PY
