#!/bin/bash
# Round-2 evidence capture (run under gpurun on ONE GPU): complete test log, full bench line, launch lists with DRAM bytes,
# full captures of the dominant kernels, kernel sweep.  Everything lands in gpurun_out/; copy what is quoted into profiles/.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_final_1gpu.txt 2>&1
tail -n 3 gpurun_out/r2_gpu_tests_final_1gpu.txt
$T 1200 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
$T 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
$T 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --launch-skip 60 -c 300 --csv --log-file gpurun_out/r2_ncu_launches_deepfm.csv \
    python bench.py --quick --steps 3 --warmup 3 --no-retrieval --cpu-steps 1 > gpurun_out/ncu_bench.log 2>&1
$T 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --launch-skip 8 -c 60 --csv --log-file gpurun_out/r2_ncu_launches_topk.csv python tools/topk_once.py > gpurun_out/ncu_topk.log 2>&1
for k in tower_bwd_dw_kernel tower_fwd3_kernel tower_bwd_dx3_kernel; do
  $T 600 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 -c 1 -f -o gpurun_out/r2_full_$k \
      python bench.py --quick --steps 3 --warmup 3 --no-retrieval --cpu-steps 1 > gpurun_out/ncu_full_$k.log 2>&1
  ncu -i gpurun_out/r2_full_$k.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$k.raw.csv 2>/dev/null
done
$T 600 ncu --set full --import-source on --clock-control none -k regex:topk_scan_kernel -s 3 -c 1 -f -o gpurun_out/r2_full_topk_scan \
    python tools/topk_once.py > gpurun_out/ncu_full_topk.log 2>&1
ncu -i gpurun_out/r2_full_topk_scan.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_topk_scan.raw.csv 2>/dev/null
$T 600 ncu --set full --import-source on --clock-control none -k regex:tower_fwd3_kernel -s 2 -c 1 -f -o gpurun_out/r2_full_tower_fwd3_B262144 \
    python tools/profile_kernels.py --only tower --sizes 262144 --once > gpurun_out/ncu_full_fwd3_big.log 2>&1
ncu -i gpurun_out/r2_full_tower_fwd3_B262144.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_tower_fwd3_B262144.raw.csv 2>/dev/null
$T 600 ncu --set full --import-source on --clock-control none -k regex:tower_bwd_dx3_kernel -s 1 -c 1 -f -o gpurun_out/r2_full_tower_dx3_B262144 \
    python tools/profile_kernels.py --only tower --sizes 262144 --once > gpurun_out/ncu_full_dx3_big.log 2>&1
ncu -i gpurun_out/r2_full_tower_dx3_B262144.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_tower_dx3_B262144.raw.csv 2>/dev/null
$T 900 python tools/profile_kernels.py > gpurun_out/r2_kernel_sweep_final.log 2>&1
tail -n 60 gpurun_out/r2_kernel_sweep_final.log
