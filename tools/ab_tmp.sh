python -m pytest tests/test_gpu_embed.py -x -q -m gpu 2>&1 | tail -3
for prio in 0 1; do
NRX_MAIN_PRIO=$prio python bench.py --steps 300 --warmup 20 --no-retrieval --cpu-steps 2 > gpurun_out/bench_p$prio.json 2> gpurun_out/bench_p$prio.err; tail -c 200 gpurun_out/bench_p$prio.err
python -c "
import json; d=json.load(open('gpurun_out/bench_p$prio.json')); print('prio$prio', d['value'], d['ms_per_step'], d['e2e']['value'], d['variants']['table_update=sparse']['ms_per_step']); print(d['kernels']['nrx_embed_bwd_plan'])"
done
