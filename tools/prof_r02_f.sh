#!/bin/bash
# Round 2, closing pass on one GPU: the whole GPU suite, smoke(), the bench line of both arms, the launch list of the
# config-3 step.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r02_gputests_g.log
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/r02_smoke_g.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r02_bench_n1_g.err
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_g_reference_arm.json 2>/dev/null ) 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_step3_launches_e.csv \
    python tools/step_config3.py 10 fused > /dev/null 2>&1
echo "ncu rc=$?"
head -c 300 gpurun_out/r02_bench_n1_g.json
