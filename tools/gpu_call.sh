set -x
timeout -s KILL 90 ./tools/_bin/umma_probe2 2>&1 | tee gpurun_out/r2_probe2b.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tail -15 | tee gpurun_out/r2_gpu_tests_a.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tower_fwd3 -c 1 -o gpurun_out/r2_fwd3_deep112 python tools/profile_kernels.py --only tower --sizes 65536 --once 2>&1 | tail -5
