#!/bin/bash
# session n (1 GPU): packed fp32x2 loss loop; A/B of register budget, joint-slot order, persistence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.log
timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2n_kernels.log 2>&1
BODYFIT_FRAME_MINB=4 timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2n_kernels_minb4.log 2>&1
BODYFIT_LJ_SORT=0 timeout 300 python tools/kernels_at.py 10000 > gpurun_out/r2n_kernels_nosort.log 2>&1
BODYFIT_FRAME_PERSIST=0 timeout 300 python tools/kernels_at.py 10000 > gpurun_out/r2n_kernels_nopersist.log 2>&1
BODYFIT_FRAME_PERSIST=0 BODYFIT_FRAME_MINB=4 timeout 300 python tools/kernels_at.py 10000 > gpurun_out/r2n_kernels_nopersist_minb4.log 2>&1
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 10000 > gpurun_out/r2n_sweep.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2n_tests.log | tail -5; grep "^{" gpurun_out/r2n_kernels*.log gpurun_out/r2n_sweep.log
