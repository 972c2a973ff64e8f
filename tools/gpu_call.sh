set -x
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -q -x -k "sharded" 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests_multi3.txt
NRX_BENCH_LEGS=cfg5 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-retrieval > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; tail -3 gpurun_out/r2_bench_n2c.err; python -c "
import json; j=json.load(open('gpurun_out/r2_bench_n2c.json')); print(j['value'], j['ms_per_step']);
[print(k, round(v['value']/1e6,1), v['ms_per_step'], v.get('exchange')) for k,v in j['legs'].items()]"
