#!/bin/bash
# round-2 second GPU session: parity suite, NS=8 effect, priority-graph e2e sweep, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
timeout 300 python tools/kernels_at.py 416 1250 2500 > gpurun_out/r2b_kernels.log 2>&1
SWEEP_PARTS=1,2,3,4 timeout 600 python tools/sweep_parts.py 1250 2500 10000 > gpurun_out/r2b_sweep.log 2>&1
SWEEP_PARTS=6,8 timeout 300 python tools/sweep_parts.py 10000 >> gpurun_out/r2b_sweep.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2b_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2b_racecheck.log
tail -4 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_kernels.log | tail -4; tail -20 gpurun_out/r2b_sweep.log; tail -4 gpurun_out/r2b_memcheck.log; tail -4 gpurun_out/r2b_racecheck.log
