#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/b4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b4_tests.log
for w in 2 4; do BODYFIT_POSE_WPB=$w timeout 300 python tools/time_kernels.py 2>&1 | tail -n 1 > gpurun_out/b4_time_wpb$w.log; done
tail -n 3 gpurun_out/b4_tests.log; for w in 2 4; do echo wpb $w; cat gpurun_out/b4_time_wpb$w.log; done
