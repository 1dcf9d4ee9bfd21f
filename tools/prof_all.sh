#!/bin/bash
# Run on the GPU box (gpurun): ncu launch list of one bench pass + one `--set full` capture per hot kernel.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
for spec in "proto:proto_umma_kernel" "accum:centroid_accum" "select:consensus_select_kernel" "plup:pseudo_label_upsampled_kernel" \
            "pl:pseudo_label_kernel" "cm:classmix_blend_kernel"; do
  which=${spec%%:*}; kern=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 2 -o gpurun_out/prof_$which -f \
      python tools/prof_one.py $which > gpurun_out/ncu_$which.log 2>&1
  tail -1 gpurun_out/ncu_$which.log
done
