#!/bin/bash
# End-of-round measurement pass (run on the GPU box under gpurun): the bench line, the launch lists of the bench command
# and of the config-3 step, the loss timings, and compute-sanitizer over the rewritten loss / label kernels.
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
tail -c 400 gpurun_out/r02_bench_n1_d.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_n1_d_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_step3_launches.csv \
    python tools/step_config3.py 10 fused > /dev/null 2>&1
RYS=32 python tools/time_losses_up.py > gpurun_out/r02_losses_up.log 2>&1
tail -3 gpurun_out/r02_losses_up.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_losses_up.py tests/test_gpu_ohem.py tests/test_gpu_labels.py -m gpu -q -x > gpurun_out/r02_memcheck_losses.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck_losses.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_losses_up.py -m gpu -q -x -k "golden or vs_oracle" > gpurun_out/r02_racecheck_losses.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck_losses.log
