#!/bin/bash
# Round-end evidence capture (run under gpurun on ONE GPU): launch list with DRAM bytes, two full captures, kernel sweep.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --launch-skip 60 -c 300 --csv --log-file gpurun_out/r1_ncu_launches_deepfm_final.csv \
    python bench.py --quick --steps 3 --warmup 3 --no-retrieval --cpu-steps 1 > gpurun_out/ncu_bench.log 2>&1
for k in tower_bwd_dw_kernel segment_reduce_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 -c 1 -f -o gpurun_out/r1_full_$k \
      python bench.py --quick --steps 3 --warmup 3 --no-retrieval --cpu-steps 1 > gpurun_out/ncu_full_$k.log 2>&1
  ncu -i gpurun_out/r1_full_$k.ncu-rep --page raw --csv > gpurun_out/r1_ncu_full_$k.raw.csv 2>/dev/null
done
python tools/profile_kernels.py > gpurun_out/sweep_final.log 2>&1
tail -40 gpurun_out/sweep_final.log
