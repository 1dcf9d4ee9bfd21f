#!/bin/bash
# Round 2, closing pass: the whole GPU suite, the bench line of both arms (timed), smoke().
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r02_gputests_f.log
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/r02_smoke_f.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r02_bench_n1_f.err
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_f_reference_arm.json 2>/dev/null ) 2>&1 | tail -4
head -c 600 gpurun_out/r02_bench_n1_f.json
