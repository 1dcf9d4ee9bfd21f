#!/bin/bash
# Round 2, late pass: the loss-path / ClassMix tests, the config-3 step timed, its launch list.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses_up.py tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q -x -k "classmix or losses or total or dropin or presence or golden" 2>&1 | tail -3
for i in 1 2 3; do python tools/step_config3.py 200 fused; done 2>&1 | tee gpurun_out/r02_step3_e.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_step3_launches_e.csv \
    python tools/step_config3.py 10 fused > /dev/null 2>&1
echo "ncu rc=$?"
