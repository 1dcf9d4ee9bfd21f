#!/bin/bash
# Round 2, late pass: the loss-path tests after the scalar-launch removal, the config-3 step timed, its launch list.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses_up.py tests/test_gpu_dropin.py tests/test_gpu_lazy_upsample.py tests/test_gpu_ohem.py -m gpu -q -x 2>&1 | tail -5
for i in 1 2 3; do python tools/step_config3.py 200 fused; done 2>&1 | tee gpurun_out/r02_step3_e.log
python tools/step_config3.py 100 dropin 2>&1 | tee -a gpurun_out/r02_step3_e.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_step3_launches_e.csv \
    python tools/step_config3.py 10 fused > /dev/null 2>&1
echo "ncu rc=$?"
