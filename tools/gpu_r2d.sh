#!/bin/bash
# round-2 session d (1 GPU): parity suite with block masks + yaw sort, A/B of the masks, BK=32 library, e2e timeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2d_kernels_mask.log 2>&1
BODYFIT_BLKMASK=0 timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2d_kernels_nomask.log 2>&1
BODYFIT_LIB=$PWD/bodyfitting_b200/libbodyfit_b200_bk32.so timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2d_kernels_bk32.log 2>&1
SWEEP_PARTS=1,2,4 SWEEP_E2E=0 timeout 600 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2d_sweep.log 2>&1
BODYFIT_PARTS=4 timeout 300 python tools/e2e_timeline.py 10000 > gpurun_out/r2d_timeline4.log 2>&1
BODYFIT_PARTS=2 timeout 300 python tools/e2e_timeline.py 10000 > gpurun_out/r2d_timeline2.log 2>&1
tail -6 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_kernels_mask.log gpurun_out/r2d_kernels_nomask.log gpurun_out/r2d_kernels_bk32.log | tail -8; tail -8 gpurun_out/r2d_sweep.log; tail -3 gpurun_out/r2d_timeline4.log; tail -3 gpurun_out/r2d_timeline2.log
