#!/bin/bash
# round-2 first GPU session: parity suite, per-kernel times at shard sizes, parts sweep, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 300 python tools/kernels_at.py 416 1250 2500 10000 > gpurun_out/r2a_kernels.log 2>&1
timeout 600 python tools/sweep_parts.py 1250 2500 5000 10000 > gpurun_out/r2a_sweep.log 2>&1
BODYFIT_GRAPH=0 SWEEP_PARTS=1,4 timeout 300 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2a_sweep_nograph.log 2>&1
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_kernels.log | tail -5; tail -20 gpurun_out/r2a_sweep.log; tail -3 gpurun_out/r2a_bench.err; head -c 600 gpurun_out/r2a_bench.json
