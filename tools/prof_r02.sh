#!/bin/bash
# Round-2 profiling pass (run on the GPU box under gpurun): launch list of the bench command + `--set full` captures of the
# kernels that changed this round (centroid chain) and of the kernels whose DRAM traffic feeds roofline.rows[].traffic.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
for spec in "accum:centroid_" "proto:proto_umma_kernel" "pl:pseudo_label_kernel" "cm:classmix_blend_kernel" "select:consensus_select_kernel"; do
  which=${spec%%:*}; kern=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 3 -o gpurun_out/r02_prof_$which -f \
      python tools/prof_one.py $which > gpurun_out/r02_ncu_$which.log 2>&1
  tail -1 gpurun_out/r02_ncu_$which.log
done
ncu --set full --clock-control none --import-source on -k regex:kd_kernel -s 4 -c 2 -o gpurun_out/r02_prof_kd -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-stages --no-e2e > gpurun_out/r02_ncu_kd.log 2>&1
tail -1 gpurun_out/r02_ncu_kd.log
