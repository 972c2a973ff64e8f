set -x
timeout -s KILL 600 python -m pytest tests/test_gpu_tower.py tests/test_gpu_models.py tests/test_gpu_trainer.py tests/test_gpu_embed.py tests/test_gpu_dssm.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -6 | tee gpurun_out/r2_gpu_tests_j.txt
timeout -s KILL 300 python tools/profile_kernels.py --only tower --sizes 16384,65536,262144 2>&1 | grep -E "bwd|training" | tee gpurun_out/r2_sweep_j.txt
NRX_TOWER_DX3=0 timeout -s KILL 300 python tools/profile_kernels.py --only tower --sizes 65536 2>&1 | grep -E "bwd" | tee -a gpurun_out/r2_sweep_j.txt
timeout -s KILL 600 python bench.py --steps 50 --warmup 10 --no-legs > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; python -c "
import json; j=json.load(open('gpurun_out/r2_bench_d.json')); print(j['value'], j['ms_per_step'], j['e2e']['value']); [print(k,v) for k,v in j['kernels'].items()]"
