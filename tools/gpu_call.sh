#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:tower_bwd_dx3_kernel -s 1 -c 1 -f -o gpurun_out/r2_full_tower_dx3_B262144 python tools/profile_kernels.py --only tower --sizes 262144 --once > gpurun_out/ncu_full_dx3_big.log 2>&1
ncu -i gpurun_out/r2_full_tower_dx3_B262144.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_tower_dx3_B262144.raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n4.json',):
    j=json.load(open(f)); r=j['retrieval']
    print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j.get('parity',{}).get('parity_ok') if j.get('parity') else None)
    print('  ', {k:(round(v['value']/1e6,1), round(v['ms_per_step'],4)) for k,v in j['legs'].items()}, r['value'], r['ms_per_search'], r['e2e']['value'])
PY
tail -n 2 gpurun_out/ncu_full_dx3_big.log
