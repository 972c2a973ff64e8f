#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python -c "
import json; j=json.load(open('gpurun_out/r2_bench_n8.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j.get('parity',{}).get('parity_ok'))
for k,v in j['legs'].items(): print(k, round(v['value']/1e6,1), v['ms_per_step'])
r=j['retrieval']; print(r['value'], r['ms_per_search'], r['sharded'])"
tail -n 3 gpurun_out/r2_bench_n8.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/topk_peer_profile.py 2>&1 | grep -v "Warn\|^\*\*\|OMP" | tail -14 > gpurun_out/r2_topk_peer_profile_n8.txt
cat gpurun_out/r2_topk_peer_profile_n8.txt
