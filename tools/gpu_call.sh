set -x
timeout -s KILL 600 python -m pytest tests/test_gpu_tower.py tests/test_gpu_models.py -q -x 2>&1 | tail -3 | tee gpurun_out/r2_gpu_tests_f.txt
timeout -s KILL 300 python tools/profile_kernels.py --only tower --sizes 65536,262144 2>&1 | grep fwd3 | tee gpurun_out/r2_sweep_f.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tower_fwd3 -c 1 -o gpurun_out/r2_fwd3_v3 python tools/profile_kernels.py --only tower --sizes 262144 --once 2>&1 | tail -2
