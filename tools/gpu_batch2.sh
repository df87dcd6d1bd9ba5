#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/b2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b2_tests.log
timeout 300 python tools/time_kernels.py > gpurun_out/b2_time_a.log 2>&1
timeout 300 python tools/time_kernels.py > gpurun_out/b2_time_b.log 2>&1
tail -n 4 gpurun_out/b2_tests.log; tail -n 1 gpurun_out/b2_time_a.log; tail -n 1 gpurun_out/b2_time_b.log
